import ctypes as C, torch, sys
sys.path.insert(0,'/root/repo')
from efa_xray_b200 import _lib
torch.cuda.set_device(0)
for name in ('exb_measure_fp64_peak','exb_measure_dmma_peak'):
    tf=C.c_double(0)
    _lib.call(name, C.byref(tf), _lib.stream_ptr()); print(name, tf.value)

import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device('cuda', 0)
n = 312_000_000 // 8
src = torch.randn(8, n, dtype=torch.float64, device=dev)
dst = torch.empty(8, n, dtype=torch.float64).pin_memory()
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); 
def kernel_work():
    for _ in range(2): torch.mm(a, a)
copy = torch.cuda.Stream(device=dev)
def run(with_copy, main_stream):
    torch.cuda.synchronize()
    with torch.cuda.stream(main_stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(8):
            kernel_work()
            if with_copy:
                ev = torch.cuda.Event(); ev.record(main_stream); copy.wait_event(ev)
                with torch.cuda.stream(copy):
                    dst[i].copy_(src[i], non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for ms in (torch.cuda.current_stream(), torch.cuda.Stream(device=dev)):
    run(False, ms)
    print('main', ms, 'kernels only', run(False, ms), 'with D2H on copy stream', run(True, ms))

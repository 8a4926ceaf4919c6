"""efa_xray_b200 -- B200-native serial EnSRF analysis step behind efa_xray's Python API.

Sub-packages mirror the reference (lmadaus/efa_xray):
    efa_xray_b200.state.ensemble           EnsembleState
    efa_xray_b200.observation.observation  Observation, gaspari_cohn, haversine
    efa_xray_b200.assimilation.ensrf       EnSRF
    efa_xray_b200.assimilation.assimilation  Assimilation, update
The arithmetic runs in hand-written sm_100a CUDA kernels (efa_xray_b200/csrc) reached through the C ABI
declared in include/efa_xray_b200.h.  Importing this package does not need a GPU; running an analysis
does, and there is no CPU fallback.
"""
__version__ = '1.0'

# How EnsembleState.interpolate treats an ob within 1 km of one of its 4 selected grid points:
#   'raise'   -> IndexError, exactly what the reference does (state/ensemble.py:195-196)
#   'nearest' -> weight 1 on the nearest point (what that branch was written to do)
EXACT_MATCH_POLICY = 'raise'

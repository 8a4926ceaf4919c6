#!/usr/bin/env python
"""efa_demo.ipynb cells 8-16 without the notebook widgets: a single-point ensemble forecast trajectory is adjusted
with dummy observations of its first valid times (Ensemble Forecast Adjustment, Madaus and Hakim 2015), through the
B200 library (efa_xray_b200.demo.enkf); tests/test_gpu_parity.py checks it against the notebook's numpy arithmetic.

    python examples/efa_demo.py [--obs-range 1 5] [--ob-error 1.0] [--inflation 1.0] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from efa_xray_b200.demo import enkf, synthetic_point_ensemble  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--obs-range', type=int, nargs=2, default=(1, 5))
ap.add_argument('--ob-error', type=float, default=1.0)
ap.add_argument('--inflation', type=float, default=1.0)
ap.add_argument('--seed', type=int, default=0)
a = ap.parse_args()

times, prior = synthetic_point_ensemble(seed=a.seed)
obs = [275.0, 275.0, 275.0, 275.0, 276.0]                     # cell 8
order = np.random.default_rng(a.seed).permutation(a.obs_range[1] - a.obs_range[0] + 1)
post = enkf(obs, prior, obs_range=tuple(a.obs_range), ob_error=a.ob_error, inflation=a.inflation, order=order)
print('valid time            prior mean  post mean   prior var  post var')
for t, pm, qm, pv, qv in zip(times, prior.mean(1), post.mean(1), prior.var(1), post.var(1)):
    print('%s  %9.3f  %9.3f  %9.4f  %9.4f' % (str(t)[:16], pm, qm, pv, qv))

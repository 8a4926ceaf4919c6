"""Multi-GPU tests (skipped on boxes with fewer than 2 GPUs): the obs-space solve distributed over the ranks (obs
dealt round-robin or in blocks of consecutive obs) with records published over NVLink peer memory must reproduce the
replicated solve bit for bit."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('nobs,nens,cutoff,block', [(3000, 50, 2500.0, 1), (2501, 100, 1500.0, 7), (3001, 50, 2500.0, 256)])
def test_distributed_obs_solve_is_bit_identical(nobs, nens, cutoff, block):
    n = _ngpus()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29571',
           os.path.join(ROOT, 'tests', 'helpers', 'dist_obs_solve_worker.py'), str(nobs), str(nens), str(cutoff), str(block)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    assert r.returncode == 0 and lines, r.stderr[-2000:]
    out = json.loads(lines[-1][7:])
    assert out['block'] == block
    assert out['ok'] and out['pairs'][0] == out['pairs'][1] and out['nan_pattern_equal']
    assert out['maxdiff_yp'] == 0.0 and out['maxdiff_ym'] == 0.0 and out['maxdiff_rec'] == 0.0

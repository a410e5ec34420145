"""Multi-GPU slab decomposition vs the single-GPU engine (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# halo paths: push (zone blocks pushed into the neighbour's inbox, one launch per exchange, default), direct (scatter kernels add
# into the neighbours' grids over NVLink, opt-in), chain (the per-substep launch chain of round 1), host (NCCL send/recv from the host)
MODES = {'direct': (1, {'PLB_SLAB_DIRECT': '1'}), 'push': (1, {}), 'chain': (1, {'PLB_SLAB_FUSED': '0'}), 'host': (0, {})}


@pytest.mark.parametrize('mode', sorted(MODES))
@pytest.mark.parametrize('dtype', ['float64', 'float32'])
def test_slab_parity_two_ranks(dtype, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    peer, extra = MODES[mode]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29611', os.path.join(ROOT, 'tests', 'multi', 'slab_parity.py'), '--dtype', dtype, '--peer', str(peer)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, **extra))
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_slab_parity_two_ranks_multi_material():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29617', os.path.join(ROOT, 'tests', 'multi', 'slab_parity.py'), '--dtype', 'float64', '--materials', '1']
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0

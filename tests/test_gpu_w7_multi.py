"""GPU test of the z-sharded path with the solver variants of SURVEY 8f: 2 ranks against the same problem on
one GPU.  Needs >= 2 GPUs on the box; skipped otherwise."""
import os
import subprocess
import sys
import pytest

from conftest import ROOT
from fbpic_b200 import _lib

pytestmark = pytest.mark.gpu


def test_two_gpu_pml_antenna_matches_one_gpu():
    """Open z + radial PML + laser antenna + moving window with injection on 2 slabs vs one domain."""
    if _lib.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MGPU_EXTRA='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29643',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and 'MGPU_EXTRA_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_gpu_diagnostics_match_one_gpu():
    """Field / particle diagnostics gathered over 2 slabs equal those of the single-domain run."""
    if _lib.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MGPU_EXTRA='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29645',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and 'MGPU_DIAG_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_gpu_checkpoint_restart():
    """A 2-slab run restarted from its per-rank checkpoints continues like the uninterrupted one."""
    if _lib.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MGPU_EXTRA='3')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29651',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and 'MGPU_RESTART_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_two_gpu_ionization_levels_migrate():
    """Ionization levels travel with the ions across the slab boundaries; one electron per event over all ranks."""
    if _lib.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, MGPU_EXTRA='4')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29649',
           os.path.join(ROOT, 'tests', 'workers', 'mgpu_parity_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and 'MGPU_IONIZATION_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

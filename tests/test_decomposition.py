"""CPU tests of the multi-GPU host logic: the z-slab decomposition rule and the guard-cell
exchange plan, including a world_size-2 run over gloo."""
import os
import subprocess
import sys
import pytest

from conftest import ROOT
from fbpic_b200.boundaries import decompose_z, halo_plan


def test_decompose_covers_domain():
    # fbpic/boundaries/boundary_communicator.py:437-457: int(Nz/size) each, last takes the rest
    for Nz, size, ng, nd, ni in ((4096, 8, 63, 0, 0), (1000, 3, 10, 64, 5), (32768, 8, 78, 64, 39)):
        phys = [decompose_z(Nz, size, r, ng, nd, ni, with_damp=False, with_guard=False) for r in range(size)]
        assert phys[0][1] == 0 and sum(n for n, _ in phys) == Nz
        for (n0, i0), (n1, i1) in zip(phys[:-1], phys[1:]):
            assert i0 + n0 == i1
        full = [decompose_z(Nz, size, r, ng, nd, ni) for r in range(size)]
        assert full[0] == (phys[0][0] + nd + ni + 2 * ng, -(nd + ni) - ng)
        assert full[-1][0] == phys[-1][0] + nd + ni + 2 * ng
        for r in range(1, size - 1):
            assert full[r] == (phys[r][0] + 2 * ng, phys[r][1] - ng)


def test_halo_plan_ranges():
    p = halo_plan(100, 7, 'replace')
    assert p == dict(send_l=(7, 14), send_r=(86, 93), recv_l=(0, 7), recv_r=(93, 100))
    p = halo_plan(100, 7, 'add')
    assert p == dict(send_l=(0, 14), send_r=(86, 100), recv_l=(0, 14), recv_r=(86, 100))
    with pytest.raises(ValueError):
        halo_plan(10, 1, 'max')


def test_halo_exchange_world_size_2_gloo():
    env = dict(os.environ, OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29631',
           os.path.join(ROOT, 'tests', 'workers', 'gloo_halo_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert 'GLOO_HALO_OK' in out.stdout


def test_gather_scatter_world_size_2_gloo():
    """gather_grid_array / scatter_grid_array / allreduce_sum of the set-up routines, 2 ranks over gloo."""
    env = dict(os.environ, OMP_NUM_THREADS='1')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29633',
           os.path.join(ROOT, 'tests', 'workers', 'gloo_gather_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert 'GLOO_GATHER_OK' in out.stdout

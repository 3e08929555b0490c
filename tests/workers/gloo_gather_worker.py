"""world_size-2 (gloo, CPU) check of the host-side global <-> local grid helpers of
fbpic_b200.boundaries.BoundaryCommunicator used by the one-off set-up routines (laser injection, bunch space
charge): gather_grid_array / scatter_grid_array / allreduce_sum
(fbpic/boundaries/boundary_communicator.py:1011-1130) on an open-z, 2-slab decomposition."""
import os
import sys
import numpy as np
import torch.distributed as dist
from scipy.constants import c

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpic_b200.boundaries import BoundaryCommunicator   # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, size = dist.get_rank(), dist.get_world_size()
    Nz, Nr, zmin, zmax, rmax = 70, 5, 0., 35.e-6, 10.e-6
    dt = (zmax - zmin) / Nz / c
    comm = BoundaryCommunicator(Nz, zmin, zmax, Nr, rmax, 2, dt, None, False, {'z': 'open', 'r': 'open'},
                                16, 12, {'z': 6, 'r': 3}, c * dt / (rmax / Nr))
    assert comm.size == size and comm.rank == rank
    for with_damp in (False, True):
        Nz_g, iz_g = comm.get_Nz_and_iz(local=False, with_damp=with_damp, with_guard=False)
        Nr_g = comm.get_Nr(with_damp=with_damp)
        glob = np.arange(Nz_g * Nr_g).reshape(Nz_g, Nr_g) * (1. + 0.5j)
        loc = comm.scatter_grid_array(glob, with_damp=with_damp)
        n_loc, iz_loc = comm.get_Nz_and_iz(local=True, with_damp=with_damp, with_guard=False, rank=rank)
        assert loc.shape == (n_loc, Nr_g) and np.array_equal(loc, glob[iz_loc - iz_g:iz_loc - iz_g + n_loc])
        # local array with guard (and damp) cells around the local part, junk in the guards
        n_arr, iz_arr = comm.get_Nz_and_iz(local=True, with_damp=True, with_guard=True, rank=rank)
        arr = np.full((n_arr, comm.get_Nr(with_damp=True)), -7. + 0.j)
        arr[iz_loc - iz_arr:iz_loc - iz_arr + n_loc, :Nr_g] = loc
        back = comm.gather_grid_array(arr, with_damp=with_damp)
        assert back.shape == glob.shape and np.array_equal(back, glob), 'gather(scatter(x)) != x'
    s0, s1 = comm.allreduce_sum([rank + 1., 10. * (rank + 1)])
    assert s0 == size * (size + 1) / 2 and s1 == 10. * s0
    dist.barrier()
    if rank == 0:
        print('GLOO_GATHER_OK')


if __name__ == '__main__':
    main()

"""world_size-2 (gloo, CPU) check of the z-slab host logic of fbpic_b200.boundaries:
`decompose_z` + `halo_plan` reproduce, with NumPy slabs exchanged over torch.distributed,
what the reference's replace/add guard-cell exchanges do to a global array
(fbpic/boundaries/boundary_communicator.py:556-707, field_buffer_handling.py:270-347)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpic_b200.boundaries import decompose_z, halo_plan   # noqa: E402


def exchange(local, plan, left, right, method):
    """Neighbour exchange of row slabs with gloo: non-blocking isend/irecv on both sides, then
    wait -- the same pattern as exchange_domains (boundary_communicator.py:674-707); with 2 ranks
    both neighbours are the same peer, so messages are told apart by direction tags (1 = sent to
    the left, 2 = sent to the right)."""
    reqs, bufs = [], {}
    for side, peer, s_key, r_key in (('l', left, 'send_l', 'recv_l'), ('r', right, 'send_r', 'recv_r')):
        if peer is None:
            continue
        s0, s1 = plan[s_key]
        send = torch.from_numpy(np.ascontiguousarray(local[s0:s1]).view(np.float64).copy())
        recv = torch.empty_like(send)
        tag_s = 1 if side == 'l' else 2
        tag_r = 2 if side == 'l' else 1
        reqs.append(dist.isend(send, peer, tag=tag_s))
        reqs.append(dist.irecv(recv, peer, tag=tag_r))
        bufs[r_key] = (recv, s1 - s0, send)
    for r in reqs:
        r.wait()
    for r_key, (recv, nrow, _) in bufs.items():
        data = recv.numpy().view(np.complex128).reshape(nrow, -1)
        r0, r1 = plan[r_key]
        if method == 'replace':
            local[r0:r1] = data
        else:
            local[r0:r1] += data


def main():
    dist.init_process_group('gloo')
    rank, size = dist.get_rank(), dist.get_world_size()
    Nz_g, Nr, ng = 64, 6, 5
    rng = np.random.default_rng(7)
    glob = rng.normal(size=(Nz_g, Nr)) + 1.j * rng.normal(size=(Nz_g, Nr))
    left, right = (rank - 1) % size, (rank + 1) % size          # periodic ring
    Nz, iz0 = decompose_z(Nz_g, size, rank, ng, 0, 0)
    rows = (np.arange(iz0, iz0 + Nz)) % Nz_g                    # global row of every local row
    # ---- replace: every rank knows only its physical rows; guards are filled by the exchange
    local = np.zeros((Nz, Nr), dtype=np.complex128)
    local[ng:Nz - ng] = glob[rows[ng:Nz - ng]]
    exchange(local, halo_plan(Nz, ng, 'replace'), left, right, 'replace')
    assert np.array_equal(local, glob[rows]), 'replace exchange'
    # ---- add: a deposited quantity, every rank holds partial sums in guard + inner rows
    part = rng.normal(size=(size, Nz_g + 2 * ng, Nr)) + 0.j     # contribution of each rank's particles
    mine = np.zeros((Nz, Nr), dtype=np.complex128)
    contrib = np.zeros((size, Nz_g, Nr), dtype=np.complex128)    # what each rank deposited, in global rows
    for r in range(size):
        nz_r, iz_r = decompose_z(Nz_g, size, r, ng, 0, 0)
        g_rows = np.arange(iz_r, iz_r + nz_r) % Nz_g
        dep = part[r, :nz_r]
        np.add.at(contrib[r], g_rows, dep)
        if r == rank:
            mine[:] = dep
    exchange(mine, halo_plan(Nz, ng, 'add'), left, right, 'add')
    total = contrib.sum(axis=0)
    # after the exchange the 2*ng overlap rows on both sides and the interior hold the full sums
    ok_rows = np.arange(0, Nz)
    assert np.allclose(mine[ok_rows], total[rows[ok_rows]], rtol=0, atol=1e-12), 'add exchange'
    dist.barrier()
    if rank == 0:
        print('GLOO_HALO_OK')


if __name__ == '__main__':
    main()

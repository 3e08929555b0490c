"""N-GPU vs 1-GPU parity of the z-sharded PIC loop (run under torchrun, one rank per GPU).
Every rank advances its slab with NCCL guard-cell exchange + particle migration; rank 0 also
advances the same global problem alone on its GPU; the physical regions must agree.

Two passes:
  * correct_currents=False: every operation of the cycle is local in z up to the stencil reach
    (finite-order PSATD), so the sharded run must reproduce the single-domain one to rounding
    (1e-9 of the field maximum).  Exercises E/B 'replace' and J 'add' guard exchanges and the
    particle migration.
  * correct_currents=True: the curl-free current correction is a global operation in z
    ("`curl-free` is faster but less local", fbpic/main.py:179-182); each slab applies it on its own
    periodic box, exactly as the reference does per MPI rank, so sharded and single-domain runs
    differ by the truncated tail of its Green's function (~1e-5 here).  Checked to 5e-4; exercises
    the spect2partial_interp / exchange / partial_interp2spect path of main.py:536-538.
Prints MGPU_PARITY_OK on success."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
from scipy.constants import c, e, m_e

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbpic_b200 import Simulation                     # noqa: E402
from fbpic_b200.particles import generate_evenly_spaced   # noqa: E402


def global_particles(Nz, Nr, zmax, rmax, n_e, seed=11):
    np.random.seed(seed)
    Ntot, x, y, z, ux, uy, uz, ig, w = generate_evenly_spaced(
        2 * Nz, 0., zmax, 2 * (Nr - 2), 0., rmax * (Nr - 2) / Nr, 8, n_e, None, 0., 0., 0., 0., 0., 0.)
    k0 = 2 * np.pi / zmax * 3
    uz = 0.2 * np.sin(k0 * z) * np.exp(-(x**2 + y**2) / (6.e-6)**2)
    ux = 0.05 * x / 6.e-6 * np.cos(k0 * z) * np.exp(-(x**2 + y**2) / (6.e-6)**2)
    uy = 0.05 * y / 6.e-6 * np.cos(k0 * z) * np.exp(-(x**2 + y**2) / (6.e-6)**2)
    ig = 1. / np.sqrt(1 + ux**2 + uy**2 + uz**2)
    return dict(x=x, y=y, z=z, ux=ux, uy=uy, uz=uz, inv_gamma=ig, w=w)


def set_species(sim, P, zlo, zhi):
    sp = sim.add_new_species(q=-e, m=m_e)
    sel = (P['z'] >= zlo) & (P['z'] < zhi)
    for k, v in P.items():
        setattr(sp, k, v[sel].copy())
    sp.Ntot = int(sel.sum())
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(sp.Ntot))
    return sp


def run_case(correct, tol):
    rank, size = dist.get_rank(), dist.get_world_size()
    nsteps = int(os.environ.get('MGPU_STEPS', '24'))
    shape = os.environ.get('MGPU_SHAPE', 'linear')
    # 96 physical cells per rank + 2*32 guard cells: every local box (160 cells) is smaller than the
    # global one, so the decomposition is non-trivial already with 2 ranks
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nz, Nr, Nm, zmax, rmax, n_e, n_order = nzr * size, 24, 2, 0.2e-6 * nzr * size, 12.e-6, 2.e24, 8
    dt = zmax / Nz / c
    P = global_particles(Nz, Nr, zmax, rmax, n_e)
    kw = dict(n_order=n_order, particle_shape=shape, boundaries={'z': 'periodic', 'r': 'reflective'})
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    assert sim.comm.size == size and sim.comm.n_guard > 0
    zlo, zhi = sim.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=rank)
    set_species(sim, P, zlo, zhi)
    sim.step(nsteps, correct_currents=correct)
    ng = sim.comm.n_guard
    names = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')
    loc = np.stack([getattr(sim.fld.interp[m], k)[ng:sim.fld.interp[m].Nz - ng] for m in range(Nm) for k in names])
    n_local = sim.ptcl[0].Ntot
    gathered = [None] * size
    dist.all_gather_object(gathered, (loc, n_local))
    ok = True
    if rank == 0:
        glob = np.concatenate([g[0] for g in gathered], axis=1)
        ref = Simulation(Nz, zmax, Nr, rmax, Nm, dt, use_all_mpi_ranks=False, n_guard=ng, **kw)
        assert ref.comm.size == 1
        set_species(ref, P, -1., 1.e9)
        ref.step(nsteps, correct_currents=correct)
        full = np.stack([getattr(ref.fld.interp[m], k) for m in range(Nm) for k in names])
        assert sum(g[1] for g in gathered) == ref.ptcl[0].Ntot, 'particle count not conserved'
        groups = {'E': (0, 3), 'B': (3, 6), 'J': (6, 9), 'rho': (9, 10)}
        for gname, (g0, g1) in groups.items():
            idx = [m * 10 + j for m in range(Nm) for j in range(g0, g1)]
            scale = max(np.abs(full[i]).max() for i in idx)
            for i in idx:
                err = np.abs(glob[i] - full[i]).max()
                if not err <= tol * scale:
                    ok = False
                    d = np.abs(glob[i] - full[i])
                    rows = np.argsort(d.max(axis=1))[::-1][:6]
                    print('MISMATCH %s m%d: err %.3e scale %.3e  worst z-rows %s (row err %s)  mean-row err %.2e'
                          % (names[i % 10], i // 10, err, scale, rows.tolist(),
                             ['%.1e' % v for v in d.max(axis=1)[rows]], d.max(axis=1).mean()))
        print('max particles/rank', [g[1] for g in gathered])
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    return bool(int(flag[0]))


def main():
    dist.init_process_group('gloo')
    ok = run_case(False, 1e-9)
    ok = run_case(True, 5e-4) and ok
    if dist.get_rank() == 0 and ok:
        print('MGPU_PARITY_OK size=%d' % dist.get_world_size())
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()

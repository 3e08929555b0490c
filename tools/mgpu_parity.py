"""N-rank vs 1-rank parity of the z-sharded PIC loop on a small periodic box (one rank per GPU, torch.distributed
already initialised).  Every rank advances its slab with the NCCL guard-cell exchange + particle migration; rank 0
also advances the same global problem alone on its GPU; the physical regions must agree:
  * correct_currents=False: every operation of the cycle is local in z up to the stencil reach (finite-order
    PSATD), so the sharded run reproduces the single-domain one to rounding (tolerance 1e-9 of the field maximum);
  * correct_currents=True: the curl-free correction is global in z and each slab applies it on its own periodic
    box, exactly as the reference does per MPI rank (fbpic/main.py:179-182, 536-538), so the two runs differ by the
    truncated tail of its Green's function (~1e-5 here; tolerance 5e-4).
Used by tests/workers/mgpu_parity_worker.py and, untimed, by `bench.py --gpus N` (the `mgpu_parity` key)."""
import numpy as np
from scipy.constants import c, e, m_e


def global_particles(Nz, Nr, zmax, rmax, n_e, seed=11):
    from fbpic_b200.particles import generate_evenly_spaced
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


def periodic_case(dist, correct, tol, nsteps=24, nzr=96, shape='linear', verbose=True):
    """Returns (ok, max relative error, particle counts per rank) on every rank (broadcast from rank 0)."""
    import torch
    from fbpic_b200 import Simulation
    rank, size = dist.get_rank(), dist.get_world_size()
    # nzr physical cells per rank + 2*32 guard cells: every local box is smaller than the global one, so the
    # decomposition is non-trivial already with 2 ranks
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
    ok, worst = True, 0.
    if rank == 0:
        glob = np.concatenate([g[0] for g in gathered], axis=1)
        ref = Simulation(Nz, zmax, Nr, rmax, Nm, dt, use_all_mpi_ranks=False, n_guard=ng, **kw)
        assert ref.comm.size == 1
        set_species(ref, P, -1., 1.e9)
        ref.step(nsteps, correct_currents=correct)
        full = np.stack([getattr(ref.fld.interp[m], k) for m in range(Nm) for k in names])
        if sum(g[1] for g in gathered) != ref.ptcl[0].Ntot:
            ok = False
            if verbose:
                print('MISMATCH particle count not conserved')
        groups = {'E': (0, 3), 'B': (3, 6), 'J': (6, 9), 'rho': (9, 10)}
        for gname, (g0, g1) in groups.items():
            idx = [m * 10 + j for m in range(Nm) for j in range(g0, g1)]
            scale = max(np.abs(full[i]).max() for i in idx)
            for i in idx:
                err = np.abs(glob[i] - full[i]).max()
                worst = max(worst, err / scale)
                if not err <= tol * scale:
                    ok = False
                    if verbose:
                        d = np.abs(glob[i] - full[i])
                        rows = np.argsort(d.max(axis=1))[::-1][:6]
                        print('MISMATCH %s m%d: err %.3e scale %.3e  worst z-rows %s (row err %s)  mean-row err %.2e'
                              % (names[i % 10], i // 10, err, scale, rows.tolist(),
                                 ['%.1e' % v for v in d.max(axis=1)[rows]], d.max(axis=1).mean()))
        if verbose:
            print('max particles/rank', [g[1] for g in gathered])
    res = torch.tensor([1. if ok else 0., worst], dtype=torch.float64)
    dist.broadcast(res, src=0)
    dist.barrier()
    return bool(res[0] > 0.5), float(res[1]), [g[1] for g in gathered]

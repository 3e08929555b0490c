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
if '--fake-device' in sys.argv:
    # CPU run of the multi-rank HOST logic (tests/test_host_flow.py): the fake library of tests/fake_device.py
    # stands in for libfbpic_b200.so and gloo for NCCL.  Never set in the GPU tests.
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import fake_device
    fake_device.install_global()
from fbpic_b200 import Simulation                     # noqa: E402
from fbpic_b200.particles import generate_evenly_spaced   # noqa: E402


sys.path.insert(0, os.path.join(ROOT, 'tools'))
from mgpu_parity import global_particles, set_species, periodic_case   # noqa: E402,F401


def run_case(correct, tol):
    ok, _, _ = periodic_case(dist, correct, tol, nsteps=int(os.environ.get('MGPU_STEPS', '24')),
                             nzr=int(os.environ.get('MGPU_NZ_PER_RANK', '96')),
                             shape=os.environ.get('MGPU_SHAPE', 'linear'))
    return ok


def window_dens(z, r):
    return np.clip((z - 4.e-6) / 3.e-6, 0., 1.)


def run_window_case(tol):
    """Open z, moving window at c, plasma injected at the right edge by the last rank, particles
    dropped at the left edge by rank 0, migration in between (correct_currents=False: local in z)."""
    rank, size = dist.get_rank(), dist.get_world_size()
    nsteps = int(os.environ.get('MGPU_WINDOW_STEPS', '40'))
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nz, Nr, Nm, rmax, n_order = nzr * size, 16, 2, 8.e-6, 8
    zmax = 0.25e-6 * Nz
    dt = zmax / Nz / c
    kw = dict(p_zmin=4.e-6, p_zmax=1., p_rmin=0, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4, n_e=1.e24,
              dens_func=window_dens, n_order=n_order, n_damp={'z': 32, 'r': 32},
              boundaries={'z': 'open', 'r': 'reflective'})

    def launch(sim):
        sim.set_moving_window(v=c)
        g1 = sim.fld.interp[1]
        zz, rr = np.meshgrid(g1.z, g1.r, indexing='ij')
        z0 = 0.6 * zmax
        prof = 3.e11 * np.exp(-(zz - z0)**2 / (2.e-6)**2) * np.exp(-rr**2 / (3.e-6)**2) * np.cos(2 * np.pi * (zz - z0) / 1.e-6)
        g1.Er[:, :], g1.Et[:, :] = 0.5 * prof, -0.5j * prof
        g1.Br[:, :], g1.Bt[:, :] = 0.5j * prof / c, 0.5 * prof / c
        np.random.seed(3)
        sim.step(nsteps, correct_currents=False)

    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    assert sim.comm.size == size
    ng = sim.comm.n_guard
    # the loader draws its azimuthal offsets from np.random: build the global plasma once (host
    # only, same seed on every rank) and give every slab the particles of its physical range
    np.random.seed(21)
    ref = Simulation(Nz, zmax, Nr, rmax, Nm, dt, use_all_mpi_ranks=False, n_guard=ng, **kw)
    zlo, zhi = sim.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=rank)
    sp, rp = sim.ptcl[0], ref.ptcl[0]
    sel = (rp.z >= zlo) & (rp.z < zhi)
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        setattr(sp, k, getattr(rp, k)[sel].copy())
    sp.Ntot = int(sel.sum())
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(sp.Ntot))
    launch(sim)
    names = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'rho')
    loc = np.stack([getattr(sim.fld.interp[m], k)[ng:sim.fld.interp[m].Nz - ng] for m in range(Nm) for k in names])
    sp = sim.ptcl[0]
    part = np.stack([getattr(sp, k) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w')])
    gathered = [None] * size
    dist.all_gather_object(gathered, (loc, part))
    ok = True
    if rank == 0:
        glob = np.concatenate([g[0] for g in gathered], axis=1)
        launch(ref)
        full = np.stack([getattr(ref.fld.interp[m], k)[ng:ref.fld.interp[m].Nz - ng] for m in range(Nm) for k in names])
        assert abs(ref.fld.interp[0].zmin - sim.fld.interp[0].zmin) < 1e-12
        for gname, (g0, g1) in {'E': (0, 3), 'B': (3, 6), 'rho': (6, 7)}.items():
            idx = [m * 7 + j for m in range(Nm) for j in range(g0, g1)]
            scale = max(np.abs(full[i]).max() for i in idx)
            for i in idx:
                err = np.abs(glob[i] - full[i]).max()
                if not err <= tol * scale:
                    ok = False
                    print('WINDOW MISMATCH %s m%d: err %.3e scale %.3e' % (names[i % 7], i // 7, err, scale))
        allp = np.concatenate([g[1] for g in gathered], axis=1)
        refp = np.stack([getattr(ref.ptcl[0], k) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w')])
        print('window: particles/rank', [g[1].shape[1] for g in gathered], 'single', refp.shape[1])
        if allp.shape != refp.shape:
            ok = False
            print('WINDOW MISMATCH particle count %s vs %s' % (allp.shape, refp.shape))
        else:
            # particle sets compared by nearest neighbour in normalised phase space (sorting on
            # the coordinates is not stable against rounding noise between equal keys)
            from scipy.spatial import cKDTree
            sc = np.abs(refp).max(axis=1, keepdims=True) + 1e-300
            d, j = cKDTree((refp / sc).T).query((allp / sc).T)
            if len(np.unique(j)) != len(j) or d.max() > 1e-7:
                ok = False
                print('WINDOW MISMATCH particles: max phase-space distance %.3e, %d unmatched'
                      % (d.max(), len(j) - len(np.unique(j))))
            else:
                print('window: particle sets agree, max normalised distance %.2e' % d.max())
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    return bool(int(flag[0]))


def run_pml_antenna_case(tol):
    """Open z + radial PML + a laser antenna sitting in the first slab + moving window with plasma
    injection (correct_currents=False: local in z).  The PML split components travel with E, B in the guard
    exchange (boundary_communicator.py:621-627); the antenna deposits on the rank that holds it
    (antenna_injection.py:171-194).  Nm = MGPU_NM (4: more slabs than one staging launch carries, the
    exchange is chunked)."""
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser
    rank, size = dist.get_rank(), dist.get_world_size()
    nsteps = int(os.environ.get('MGPU_WINDOW_STEPS', '40'))
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nm = int(os.environ.get('MGPU_NM', '2'))
    Nz, Nr, rmax, n_order = nzr * size, 12, 6.e-6, 8
    zmax = 0.25e-6 * Nz
    dt = zmax / Nz / c
    kw = dict(p_zmin=0.45 * zmax, p_zmax=1., p_rmin=0, p_rmax=5.e-6, p_nz=2, p_nr=2, p_nt=4, n_e=1.e24,
              n_order=n_order, n_damp={'z': 32, 'r': 6}, boundaries={'z': 'open', 'r': 'open'})

    def launch(sim):
        prof = GaussianLaser(a0=1., waist=2.5e-6, tau=4.e-15, z0=0.1 * zmax - 3.e-6, zf=0.5 * zmax,
                             theta_pol=0.3, lambda0=1.e-6)
        add_laser_pulse(sim, prof, method='antenna', z0_antenna=0.1 * zmax)
        sim.set_moving_window(v=0.5 * c)
        np.random.seed(3)
        sim.step(nsteps, correct_currents=False)

    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    assert sim.comm.size == size and sim.use_pml
    ng = sim.comm.n_guard
    np.random.seed(21)
    ref = Simulation(Nz, zmax, Nr, rmax, Nm, dt, use_all_mpi_ranks=False, n_guard=ng, **kw)
    zlo, zhi = sim.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=rank)
    sp, rp = sim.ptcl[0], ref.ptcl[0]
    sel = (rp.z >= zlo) & (rp.z < zhi)
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        setattr(sp, k, getattr(rp, k)[sel].copy())
    sp.Ntot = int(sel.sum())
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(sp.Ntot))
    launch(sim)
    names = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml', 'rho')
    nn = len(names)
    loc = np.stack([getattr(sim.fld.interp[m], k)[ng:sim.fld.interp[m].Nz - ng] for m in range(Nm) for k in names])
    gathered = [None] * size
    dist.all_gather_object(gathered, (loc, sim.ptcl[0].Ntot))
    ok = True
    if rank == 0:
        glob = np.concatenate([g[0] for g in gathered], axis=1)
        launch(ref)
        full = np.stack([getattr(ref.fld.interp[m], k)[ng:ref.fld.interp[m].Nz - ng] for m in range(Nm) for k in names])
        assert abs(ref.fld.interp[0].zmin - sim.fld.interp[0].zmin) < 1e-12
        # PML components are compared on the scale of their parent field
        for gname, cols in {'E': (0, 1, 2, 6, 7), 'B': (3, 4, 5, 8, 9), 'rho': (10,)}.items():
            idx = [m * nn + j for m in range(Nm) for j in cols]
            scale = max(np.abs(full[i]).max() for i in idx)
            assert scale > 0, gname
            for i in idx:
                err = np.abs(glob[i] - full[i]).max()
                if not err <= tol * scale:
                    ok = False
                    print('PML/ANTENNA MISMATCH %s m%d: err %.3e scale %.3e' % (names[i % nn], i // nn, err, scale))
        if sum(g[1] for g in gathered) != ref.ptcl[0].Ntot:
            ok = False
            print('PML/ANTENNA MISMATCH particle count', [g[1] for g in gathered], ref.ptcl[0].Ntot)
        print('pml+antenna: particles/rank', [g[1] for g in gathered], 'single', ref.ptcl[0].Ntot)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    return bool(int(flag[0]))


def run_diag_case(tol):
    """Field / particle diagnostics of a sharded run (gathered over the ranks, written by rank 0) against those of
    the single-domain run: same files (fbpic_b200/diags.py; gathering rules of field_diag.py:192-212)."""
    import tempfile
    from fbpic_b200.diags import (FieldDiagnostic, ParticleDiagnostic, BackTransformedFieldDiagnostic,
                                  BackTransformedParticleDiagnostic)
    rank, size = dist.get_rank(), dist.get_world_size()
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nz, Nr, Nm, zmax, rmax, n_e, n_order = nzr * size, 16, 2, 0.2e-6 * nzr * size, 8.e-6, 2.e24, 8
    dt = zmax / Nz / c
    P = global_particles(Nz, Nr, zmax, rmax, n_e)
    P['uz'] = P['uz'] + 0.3          # a drift, so that particles cross the slab boundaries (and the ring closure)
    P['inv_gamma'] = 1. / np.sqrt(1 + P['ux']**2 + P['uy']**2 + P['uz']**2)
    kw = dict(n_order=n_order, boundaries={'z': 'periodic', 'r': 'reflective'})
    dirs = [tempfile.mkdtemp() if rank == 0 else None, tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(dirs, src=0)

    def run(sim, zlo, zhi, d):
        sp = set_species(sim, P, zlo, zhi)
        sp.track(sim.comm)
        sim.id_w_before = (np.array(sp.tracker.id), np.array(sp.w), np.array(sp.x))
        sim.diags = [FieldDiagnostic(period=3, fldobject=sim.fld, comm=sim.comm, fieldtypes=['E', 'B', 'rho'], write_dir=d),
                     ParticleDiagnostic(period=3, species={'e': sp}, comm=sim.comm, select={'uz': [0.05, None]},
                                        write_dir=d)]
        # lab-frame slicing (gamma = 2) of the same data: the plane of snapshot 1 starts 2 cells right of the middle
        # of the box and moves left by 1.15 cells per cycle, i.e. across the boundary between the slabs of a 2-rank run
        gamma, beta = 2., np.sqrt(0.75)
        t_lab = (0.5 * zmax + 2.2 * zmax / Nz) * gamma * beta / c
        sim.diags.append(BackTransformedFieldDiagnostic(0., 2 * gamma * zmax, 0., t_lab, 2, gamma, 3, sim.fld,
                                                        comm=sim.comm, fieldtypes=['E', 'B', 'rho'],
                                                        write_dir=os.path.join(d, 'lab')))
        sim.diags.append(BackTransformedParticleDiagnostic(0., 2 * gamma * zmax, 0., t_lab, 2, gamma, 3, sim.fld,
                                                           species={'e': sp}, comm=sim.comm,
                                                           write_dir=os.path.join(d, 'lab')))
        sim.step(5, correct_currents=False)
        sim.step(1, correct_currents=False)      # a new call starts with a particle exchange: migration happens

    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    zlo, zhi = sim.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=rank)
    run(sim, zlo, zhi, dirs[0])
    ok = True
    # tracked ids migrate with their particles: every id is still unique over the ranks and labels a particle
    # with the weight (never modified by the cycle) it had at the start
    sp = sim.ptcl[0]
    both = [None] * size
    dist.all_gather_object(both, (sim.id_w_before[0], sim.id_w_before[1], np.array(sp.tracker.id), np.array(sp.w)))
    if rank == 0:
        id0, w0 = np.concatenate([b[0] for b in both]), np.concatenate([b[1] for b in both])
        id1, w1 = np.concatenate([b[2] for b in both]), np.concatenate([b[3] for b in both])
        moved = sum(len(np.setdiff1d(b[2], b[0])) for b in both)
        if len(np.unique(id1)) != len(id1) or not np.array_equal(np.sort(id0), np.sort(id1)) \
                or not np.array_equal(w0[np.argsort(id0)], w1[np.argsort(id1)]):
            ok = False
            print('DIAG MISMATCH tracked ids')
        print('diag: %d tracked particles, %d changed rank' % (len(id1), moved))
        if moved == 0:
            ok = False
            print('DIAG MISMATCH: no particle migrated, the id exchange was not exercised')
    if rank == 0:
        ref = Simulation(Nz, zmax, Nr, rmax, Nm, dt, use_all_mpi_ranks=False, n_guard=sim.comm.n_guard, **kw)
        run(ref, -1., 1.e9, dirs[1])
        for it in (0, 3):
            from fbpic_b200.diags import read_diag
            a = pa = read_diag(dirs[0], it)
            b = pb = read_diag(dirs[1], it)
            for grp in ('E', 'B'):
                scale = max(np.abs(b['fields/%s/%s' % (grp, k)]).max() for k in 'rtz') + 1e-300
                for k in 'rtz':
                    key = 'fields/%s/%s' % (grp, k)
                    if a[key].shape != b[key].shape or not np.abs(a[key] - b[key]).max() <= tol * scale:
                        ok = False
                        print('DIAG MISMATCH', it, key, a[key].shape, b[key].shape)
            if not np.abs(a['fields/rho'] - b['fields/rho']).max() <= tol * np.abs(b['fields/rho']).max():
                ok = False
                print('DIAG MISMATCH rho', it)
            # (the two runs wrap / migrate their particles at different iterations: compare modulo the box length)
            za, zb = np.sort(pa['particles/e/position/z'] % zmax), np.sort(pb['particles/e/position/z'] % zmax)
            if za.shape != zb.shape or (len(za) and np.abs(za - zb).max() > 1e-9 * zmax):
                ok = False
                print('DIAG MISMATCH particles', it, za.shape, zb.shape, np.abs(za - zb).max() if za.shape == zb.shape else '',
                      za.min(), zb.min(), za.max(), zb.max())
        print('diag: selected particles at iteration 3:', len(za))
        a, b = read_diag(os.path.join(dirs[0], 'lab'), 1), read_diag(os.path.join(dirs[1], 'lab'), 1)
        filled = np.flatnonzero(np.abs(b['fields/E/z']).sum(axis=(0, 1)))
        print('diag: lab-frame snapshot, filled columns', filled)
        if len(filled) < 3:
            ok = False
            print('DIAG MISMATCH: lab-frame snapshot holds %d slices' % len(filled))
        # (the ids differ between the runs: rank r hands out r, r + size, ...; x and y do not change in this drift)
        order = lambda t: np.lexsort((np.round(t['particles/e/position/y'] / 1.e-12),       # noqa: E731
                                      np.round(t['particles/e/position/x'] / 1.e-12)))
        ia, ib = order(a), order(b)
        print('diag: lab-frame snapshot, particles caught', len(ia), len(ib))
        if len(ib) < 50 or len(ia) != len(ib) or len(np.unique(a['particles/e/id'])) != len(ia):
            ok = False
            print('DIAG MISMATCH lab particle number / ids')
        else:
            for key, ref_scale in (('position/x', rmax), ('position/z', zmax), ('momentum/z', 9.1e-31 * c), ('weighting', 0.)):
                va, vb = a['particles/e/' + key][ia], b['particles/e/' + key][ib]
                if not np.abs(va - vb).max() <= 1e-9 * (ref_scale or np.abs(vb).max()):
                    ok = False
                    print('DIAG MISMATCH lab particles', key, np.abs(va - vb).max())
        for key in [k for k in b if k.startswith('fields/') and '@' not in k]:
            record = key.rsplit('/', 1)[0] if key[-2:] in ('/r', '/t', '/z') else key
            scale = max(np.abs(b[k]).max() for k in b if '@' not in k and (k == record or k.startswith(record + '/')))
            if a[key].shape != b[key].shape or not np.abs(a[key] - b[key]).max() <= tol * scale or scale == 0:
                ok = False
                print('DIAG MISMATCH lab', key, np.abs(a[key] - b[key]).max(), scale)
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    return bool(int(flag[0]))


def run_restart_case(tol):
    """Checkpoint / restart of a sharded run (checkpoint_restart.py:22-189): every rank writes and reads its own
    `proc<rank>` directory; the restarted run continues like the uninterrupted one, rank by rank."""
    import tempfile
    from fbpic_b200.diags import set_periodic_checkpoint, restart_from_checkpoint
    rank, size = dist.get_rank(), dist.get_world_size()
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nz, Nr, Nm, zmax, rmax, n_e, n_order = nzr * size, 16, 2, 0.2e-6 * nzr * size, 8.e-6, 2.e24, 8
    dt = zmax / Nz / c
    P = global_particles(Nz, Nr, zmax, rmax, n_e)
    P['uz'] = P['uz'] + 0.3          # particles cross the slab boundaries between the checkpoint and the end
    P['inv_gamma'] = 1. / np.sqrt(1 + P['ux']**2 + P['uy']**2 + P['uz']**2)
    kw = dict(n_order=n_order, boundaries={'z': 'periodic', 'r': 'reflective'})
    d = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    a = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    zlo, zhi = a.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=rank)
    sp = set_species(a, P, zlo, zhi)
    sp.track(a.comm)
    set_periodic_checkpoint(a, 4, checkpoint_dir=d[0])
    a.step(4, correct_currents=False)
    a.step(3, correct_currents=False)
    b = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    set_species(b, P, zlo, zhi).track(b.comm)
    restart_from_checkpoint(b, checkpoint_dir=d[0])
    ok = b.iteration == 4
    b.step(3, correct_currents=False)
    sa, sb = a.ptcl[0], b.ptcl[0]
    if sa.Ntot != sb.Ntot or not np.array_equal(np.sort(sa.tracker.id), np.sort(sb.tracker.id)):
        ok = False
        print('RESTART MISMATCH rank %d: particle number / ids %d %d' % (rank, sa.Ntot, sb.Ntot))
    else:
        ia, ib = np.argsort(sa.tracker.id), np.argsort(sb.tracker.id)
        for k, scale in (('x', rmax), ('z', zmax), ('ux', 1.), ('uz', 1.), ('w', 0.)):
            va, vb = np.asarray(getattr(sa, k))[ia], np.asarray(getattr(sb, k))[ib]
            if not np.abs(va - vb).max() <= tol * (scale or np.abs(va).max()):
                ok = False
                print('RESTART MISMATCH rank %d particles %s: %.3e' % (rank, k, np.abs(va - vb).max()))
    for grp in ('E', 'B'):
        scale = max(np.abs(getattr(a.fld.interp[m], grp + k)).max() for m in range(Nm) for k in 'rtz')
        for m in range(Nm):
            for k in 'rtz':
                err = np.abs(getattr(a.fld.interp[m], grp + k) - getattr(b.fld.interp[m], grp + k)).max()
                if not err <= tol * scale:
                    ok = False
                    print('RESTART MISMATCH rank %d %s%s m%d: %.3e vs %.3e' % (rank, grp, k, m, err, scale))
    if rank == 0:
        print('restart: rank 0 holds %d particles, %d of them came from another rank'
              % (sb.Ntot, int(np.sum(np.asarray(sb.tracker.id) % size != rank))))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    return bool(int(flag[0]))


def run_ionization_case():
    """Ionizable species on 2 slabs (particles/elementary_process/ionization; particle_buffer_handling.py:120-172,
    413-417): drifting nitrogen ions in a static field cross the slab boundaries and the ring closure; their
    ionization levels travel with them (an ion never loses charge, identified by its tracked id), the deposition weight
    stays w * level on every rank, and over all ranks one electron appears per ionization event."""
    from scipy.constants import m_p
    rank, size = dist.get_rank(), dist.get_world_size()
    nzr = int(os.environ.get('MGPU_NZ_PER_RANK', '96'))
    Nz, Nr, Nm, zmax, rmax, n_order = nzr * size, 8, 2, 0.25e-6 * nzr * size, 6.e-6, 8
    dt = zmax / Nz / c
    np.random.seed(5)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, n_order=n_order, boundaries={'z': 'periodic', 'r': 'reflective'})
    kw = dict(p_zmin=0., p_zmax=zmax, p_rmin=0., p_rmax=5.e-6, p_nz=1, p_nr=2, p_nt=4, continuous_injection=False)
    elec = sim.add_new_species(q=-e, m=m_e)
    ions = sim.add_new_species(q=0, m=14. * m_p, n=1.e18, uz_m=0.5, **kw)
    ions.make_ionizable('N', target_species=elec, level_start=1, level_max=5)
    ions.track(sim.comm)
    for m in range(Nm):
        if m == 0:
            sim.fld.interp[0].Ez[:, :] = 1.8e11
    ok = True
    seen = {}
    start = [None] * size
    dist.all_gather_object(start, (np.array(ions.tracker.id), ions.ionizer.ionization_level.copy()))
    for ids, lv in start:
        seen.update(zip(ids.tolist(), lv.tolist()))
    n_e0 = torch.tensor([elec.Ntot]); dist.all_reduce(n_e0)
    lev0 = sum(seen.values())
    for _ in range(3):
        sim.step(8, correct_currents=False, move_momenta=False)      # rigid drift: nobody leaves radially
        lv = ions.ionizer.ionization_level
        if len(lv) != ions.Ntot or not np.array_equal(ions.ionizer.w_times_level, np.array(ions.w) * lv):
            ok = False
            print('IONIZATION MISMATCH rank %d: weights / lengths' % rank)
        now = [None] * size
        dist.all_gather_object(now, (np.array(ions.tracker.id), lv.copy()))
        allids = np.concatenate([a for a, _ in now])
        if len(np.unique(allids)) != len(allids) or len(allids) != len(seen):
            ok = False
            print('IONIZATION MISMATCH: ids', len(allids), len(seen))
        for ids, levels in now:
            for pid, level in zip(ids.tolist(), levels.tolist()):
                if level < seen[pid]:
                    ok = False
                seen[pid] = level
    n_e = torch.tensor([elec.Ntot]); dist.all_reduce(n_e)
    moved = int(np.sum(np.array(ions.tracker.id) % size != rank))
    if int(n_e[0] - n_e0[0]) != sum(seen.values()) - lev0 or sum(seen.values()) == lev0:
        ok = False
        print('IONIZATION MISMATCH: electrons %d vs events %d' % (int(n_e[0] - n_e0[0]), sum(seen.values()) - lev0))
    if rank == 0:
        print('ionization: %d events, %d electrons, %d ions on rank 0 came from another rank'
              % (sum(seen.values()) - lev0, int(n_e[0] - n_e0[0]), moved))
        if moved == 0:
            ok = False
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    return bool(int(flag[0]))


def main():
    dist.init_process_group('gloo')
    if os.environ.get('MGPU_EXTRA') == '4':
        ok = run_ionization_case()
        if dist.get_rank() == 0 and ok:
            print('MGPU_IONIZATION_OK size=%d' % dist.get_world_size())
        sys.exit(0 if ok else 1)
    if os.environ.get('MGPU_EXTRA') == '3':
        ok = run_restart_case(1e-9)
        if dist.get_rank() == 0 and ok:
            print('MGPU_RESTART_OK size=%d' % dist.get_world_size())
        sys.exit(0 if ok else 1)
    if os.environ.get('MGPU_EXTRA') == '2':
        ok = run_diag_case(1e-9)
        if dist.get_rank() == 0 and ok:
            print('MGPU_DIAG_OK size=%d' % dist.get_world_size())
        sys.exit(0 if ok else 1)
    if os.environ.get('MGPU_EXTRA') == '1':
        ok = run_pml_antenna_case(1e-8)
        if dist.get_rank() == 0 and ok:
            print('MGPU_EXTRA_OK size=%d' % dist.get_world_size())
        sys.exit(0 if ok else 1)
    ok = run_case(False, 1e-9)
    ok = run_case(True, 5e-4) and ok
    ok = run_window_case(1e-8) and ok
    if dist.get_rank() == 0 and ok:
        print('MGPU_PARITY_OK size=%d' % dist.get_world_size())
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()

"""GPU tests of the diagnostics / checkpoint hooks of `Simulation.step()` (fbpic_b200/diags.py): the openPMD trees
written by the field, particle, charge-density, checkpoint and lab-frame (back-transformed) diagnostics against the
trees the unmodified reference wrote for the same simulations (tests/golden/diags_tree.npz, diags_lab_tree.npz,
harvested through the h5py stand-in of oracle/ref_shim); a diagnostic written during a step holds the data the
simulation itself reports for that iteration; a run restarted from a checkpoint continues like the uninterrupted
one."""
import numpy as np
import pytest
from scipy.constants import c

from conftest import assert_close

pytestmark = pytest.mark.gpu


# slightly faster than one cell per step: the window position never sits a rounding error below a cell boundary,
# so the cell count moved per step does not depend on the sub-cell offset that a restart (here as in the
# reference, which re-creates the MovingWindow at the grid position) does not carry over
V_WINDOW = c * (1 + 1.e-6)


def _sim(fused, window, **kw):
    from fbpic_b200 import Simulation
    Nz, Nr, Nm, zmax, rmax = 32, 12, 2, 16.e-6, 8.e-6
    np.random.seed(4)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, zmax / Nz / c, p_zmin=2.e-6, p_zmax=40.e-6, p_rmin=0, p_rmax=6.e-6,
                     p_nz=2, p_nr=2, p_nt=4, n_e=2.e24, n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 4},
                     boundaries={'z': 'open', 'r': 'reflective'}, fused=fused, **kw)
    sp = sim.ptcl[0]
    # the plasma near the right edge stays exactly at rest during the test (no signal reaches it in 7 cycles): the
    # continuous injector, re-initialised from the particle positions after a restart, then continues the same lattice
    inner = sp.z < 10.e-6
    sp.uz = inner * 0.3 * np.sin(2 * np.pi * sp.z / 8.e-6) * np.exp(-(sp.x**2 + sp.y**2) / (3.e-6)**2)
    sp.ux = inner * 0.05 * sp.x / 3.e-6 * np.cos(2 * np.pi * sp.z / 8.e-6)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    if window:
        sim.set_moving_window(v=V_WINDOW)
    return sim


def _modes(d, Nm):
    """openPMD thetaMode dataset [2 Nm - 1, Nr, Nz] -> list of complex [Nz, Nr] arrays"""
    return [d[0].T + 0.j] + [0.5 * (d[2 * m - 1] + 1.j * d[2 * m]).T for m in range(1, Nm)]


@pytest.mark.parametrize('fused', [False, True])
def test_diagnostics_hold_the_state_of_their_iteration(fused, tmp_path):
    from fbpic_b200.diags import FieldDiagnostic, ParticleDiagnostic
    sim = _sim(fused, window=True)
    sp = sim.ptcl[0]
    sim.diags = [FieldDiagnostic(period=5, fldobject=sim.fld, comm=sim.comm, write_dir=str(tmp_path)),
                 ParticleDiagnostic(period=5, species={'electrons': sp}, comm=sim.comm, select={'uz': [0.05, None]},
                                    particle_data=['position', 'momentum', 'weighting', 'gamma'],
                                    write_dir=str(tmp_path))]
    sim.step(5)                                   # writes iteration 0; the state now is that of iteration 5
    ref = {k: [sim.comm.gather_grid_array(getattr(sim.fld.interp[m], k)) for m in range(2)]
           for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'rho')}
    part = {k: np.array(getattr(sp, k)) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w', 'inv_gamma')}
    zmin_phys, _ = sim.comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
    sim.step(1)                                   # iteration 5 is due at the start of this cycle
    from fbpic_b200.diags import read_diag
    from scipy.constants import m_e
    f = p = read_diag(str(tmp_path), 5)
    assert abs(float(f['time']) - 5 * sim.dt) < 1e-25 and abs(f['zmin'] - zmin_phys) < 1e-12
    assert f['fields/E/r'].shape == (3, 12, 32)
    for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
        got = _modes(f['fields/%s/%s' % (k[0], k[1])], 2)
        scale = max(np.abs(ref[k[0] + c_][m]).max() for c_ in 'rtz' for m in range(2)) + 1e-300
        for m in range(2):
            assert_close(got[m], ref[k][m] if m else ref[k][m].real, 1e-13, 'diag %s m%d' % (k, m), scale=scale)
    got = _modes(f['fields/rho'], 2)
    scale = max(np.abs(ref['rho'][m]).max() for m in range(2))
    for m in range(2):          # re-deposited after a re-sort: summation order differs
        assert_close(got[m], ref['rho'][m] if m else ref['rho'][m].real, 1e-11, 'diag rho m%d' % m, scale=scale)
    assert np.all(np.isfinite(f['fields/J/z']))
    sel = part['uz'] > 0.05
    assert sel.sum() > 0 and len(p['particles/electrons/position/x']) == sel.sum()
    order_ref = np.lexsort((part['z'][sel], part['x'][sel]))
    order_got = np.lexsort((p['particles/electrons/position/z'], p['particles/electrons/position/x']))
    for key, k, unit in (('position/x', 'x', 1.), ('position/z', 'z', 1.), ('momentum/z', 'uz', m_e * c),
                         ('weighting', 'w', 1.)):
        assert_close(p['particles/electrons/' + key][order_got] / unit, part[k][sel][order_ref], 1e-14, 'diag ' + k)
    assert_close(p['particles/electrons/gamma'][order_got], 1. / part['inv_gamma'][sel][order_ref], 1e-14, 'gamma')


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('window', [False, True])
def test_restart_from_checkpoint_continues_the_run(fused, window, tmp_path):
    from fbpic_b200.diags import set_periodic_checkpoint, restart_from_checkpoint
    a = _sim(fused, window)
    set_periodic_checkpoint(a, 4, checkpoint_dir=str(tmp_path))
    a.step(4)                                     # checkpoint of iteration 4
    np.random.seed(77)                            # the azimuths of the injected plasma are drawn from np.random
    a.step(3)
    np.random.seed(9)
    b = _sim(fused, False)
    restart_from_checkpoint(b, checkpoint_dir=str(tmp_path))
    assert b.iteration == 4 and abs(b.time - 4 * b.dt) < 1e-20
    if window:      # as in the reference's lwfa_script.py: restart first, then attach the moving window
        b.set_moving_window(v=V_WINDOW)
    np.random.seed(77)
    b.step(3)
    assert b.iteration == a.iteration and abs(b.fld.interp[0].zmin - a.fld.interp[0].zmin) < 1e-12
    sa, sb = a.ptcl[0], b.ptcl[0]
    assert sa.Ntot == sb.Ntot
    # With the moving window the restarted run re-initialises the continuous injector (as the reference does), so
    # the plasma slices can enter in different batches and draw different azimuths from np.random: the axisymmetric
    # quantities (r, z, momenta along z, weights; mode-0 fields) are the ones a restart must reproduce.
    pa = dict(r=np.hypot(sa.x, sa.y), z=np.array(sa.z), uz=np.array(sa.uz), w=np.array(sa.w))
    pb = dict(r=np.hypot(sb.x, sb.y), z=np.array(sb.z), uz=np.array(sb.uz), w=np.array(sb.w))
    # several particles share one (z, r) (different azimuths): sort on coordinates rounded far above the rounding
    # noise and far below the lattice spacing, then on uz
    order = lambda p: np.lexsort((p['uz'], np.round(p['r'] / 1.e-13), np.round(p['z'] / 1.e-13)))      # noqa: E731
    oa, ob = order(pa), order(pb)
    for k in pa:
        assert_close(pb[k][ob], pa[k][oa], 1e-10, 'restart ' + k)
    if not window:
        oa, ob = np.lexsort((sa.z, sa.y, sa.x, sa.w)), np.lexsort((sb.z, sb.y, sb.x, sb.w))
        for k in ('x', 'y', 'ux', 'uy'):
            assert_close(np.array(getattr(sb, k))[ob], np.array(getattr(sa, k))[oa], 1e-10, 'restart ' + k)
    for m in range(1 if window else 2):
        for grp in ('E', 'B'):
            scale = max(np.abs(getattr(a.fld.interp[mm], grp + c_)).max() for c_ in 'rtz' for mm in range(2))
            for c_ in 'rtz':
                assert_close(getattr(b.fld.interp[m], grp + c_), getattr(a.fld.interp[m], grp + c_), 1e-9,
                             'restart %s%s m%d' % (grp, c_, m), scale=scale)


@pytest.mark.parametrize('fused', [False, True])
def test_tracked_ids_follow_the_particles(fused):
    """Particle tracking (tracking.py:15-130) against the unmodified reference on the same inputs: after 20 cycles
    with a moving window (particles dropped at the left edge, plasma injected at the right one, several cell
    sorts) every id still labels the same particle, and the injected particles got the same new ids."""
    from fbpic_b200 import Simulation
    from conftest import load_golden
    g = load_golden('tracking_window')
    np.random.seed(5)
    sim = Simulation(int(g['Nz']), float(g['zmax']), int(g['Nr']), float(g['rmax']), int(g['Nm']), float(g['dt']),
                     p_zmin=4.e-6, p_zmax=60.e-6, p_rmin=0, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4, n_e=1.e24, n_order=-1,
                     n_guard=12, n_damp={'z': 12, 'r': 4}, boundaries={'z': 'open', 'r': 'reflective'}, fused=fused)
    sp = sim.ptcl[0]
    assert sp.Ntot == int(g['n_in'])
    sp.uz[:] = 0.4 * np.sin(2 * np.pi * sp.z / 8.e-6) * (sp.z < 10.e-6)
    sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.uz**2)
    sp.track(sim.comm)
    assert np.array_equal(sp.tracker.id, g['id_in'])
    sim.set_moving_window(v=c)
    np.random.seed(6)
    sim.step(int(g['nsteps']))
    ids = np.asarray(sp.tracker.id)
    assert ids.dtype == np.uint64 and len(ids) == sp.Ntot == len(g['id_out'])
    assert len(np.unique(ids)) == len(ids)
    assert np.array_equal(np.sort(ids), np.sort(g['id_out']))
    o, ro = np.argsort(ids), np.argsort(g['id_out'])
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        assert_close(np.asarray(getattr(sp, k))[o], g['out_' + k][ro], 1e-10, 'tracked ' + k)


def _namespace():
    import types
    from fbpic_b200 import Simulation
    from fbpic_b200 import openpmd_diag as d
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from fbpic_b200.lpa_utils.bunch import add_elec_bunch_gaussian
    return types.SimpleNamespace(Simulation=Simulation, FieldDiagnostic=d.FieldDiagnostic,
                                 ParticleDiagnostic=d.ParticleDiagnostic,
                                 ParticleChargeDensityDiagnostic=d.ParticleChargeDensityDiagnostic,
                                 BackTransformedFieldDiagnostic=d.BackTransformedFieldDiagnostic,
                                 BackTransformedParticleDiagnostic=d.BackTransformedParticleDiagnostic,
                                 set_periodic_checkpoint=d.set_periodic_checkpoint, add_laser_pulse=add_laser_pulse,
                                 add_elec_bunch_gaussian=add_elec_bunch_gaussian,
                                 GaussianLaser=GaussianLaser, BoostConverter=BoostConverter)


@pytest.mark.parametrize('fused', [False, True])
def test_diagnostic_trees_vs_reference_golden(fused, tmp_path):
    """Field, particle (all records / selection; tracked ids), per-species charge density and checkpoint output of
    the same 5-cycle run, file by file and path by path against what the reference wrote: same groups, datasets,
    shapes, dtypes and openPMD attributes; data within 1e-9 of the scale of its record (particles matched by id
    resp. position)."""
    import diag_cases
    from conftest import load_golden
    g = load_golden('diags_tree')
    ns = _namespace()
    sim, elec, ions = diag_cases.build_diag_sim(ns, fused=fused)
    dirs = diag_cases.attach_diags(ns, sim, elec, ions, str(tmp_path))
    np.random.seed(24)
    sim.step(diag_cases.DIAG_STEPS)
    for d, tag in zip(dirs, diag_cases.DIAG_DIRS):
        ref, got = diag_cases.golden_files(g, tag), diag_cases.written_files(d)
        assert sorted(ref) == sorted(got), (tag, sorted(ref), sorted(got))
        for name in ref:
            diag_cases.compare_trees(got[name], ref[name], 1e-9, '%s/%s' % (tag, name))


@pytest.mark.parametrize('fused', [False, True])
def test_lab_frame_snapshots_vs_reference_golden(fused, tmp_path):
    """BackTransformedFieldDiagnostic / BackTransformedParticleDiagnostic: 4 lab-frame snapshots of E, B, J, rho and
    of the (tracked) electrons, collected during 40 cycles of a boosted-frame run with a moving window -- the grid
    slice and the particles crossing the plane of each snapshot are extracted on the device, the Lorentz
    transformation and the placement in the lab-frame files happen on the host -- against the files of the
    reference; a second particle series with a selection on the lab-frame quantities."""
    import diag_cases
    from conftest import load_golden
    g = load_golden('diags_lab_tree')
    ns = _namespace()
    sim, gamma_boost = diag_cases.build_lab_diag_sim(ns, fused=fused)
    d = diag_cases.attach_lab_diag(ns, sim, gamma_boost, str(tmp_path))
    np.random.seed(30)
    sim.step(diag_cases.LAB_DIAG_STEPS)
    ref, got = diag_cases.golden_files(g, 'lab'), diag_cases.written_files(d)
    assert sorted(ref) == sorted(got) and len(ref) == 4
    caught = 0
    for name in ref:
        diag_cases.compare_trees(got[name], ref[name], 1e-8, 'lab/' + name)
        caught += sum(len(v) for k, v in got[name].items() if k.endswith('/electrons/id'))
    assert caught > 500
    ref, got = diag_cases.golden_files(g, 'labsel'), diag_cases.written_files(str(tmp_path / 'selected'))
    assert sorted(ref) == sorted(got) and len(ref) == 4
    selected = 0
    for name in ref:
        diag_cases.compare_trees(got[name], ref[name], 1e-8, 'labsel/' + name)
        selected += sum(len(v) for k, v in got[name].items() if k.endswith('/electrons/id'))
    assert 0 < selected < caught


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_cpu_gpu_deposition_as_written(shape, fused, tmp_path):
    """The reference's own CPU-vs-GPU parity test (tests/test_cpu_gpu_deposition.py), as written: rho and J of a
    Gaussian bunch after 0, 1, 2 cycles, compared through the FieldDiagnostic files with the reference's tolerance
    1e-13 (max|F_cpu| + max|F_gpu|).  The CPU arm is the unmodified reference (fixture), the GPU arm is this
    library; every thetaMode row of rho, Jr, Jt, Jz is compared (the reference looks at rho, Jx, Jz at theta = 0)."""
    import diag_cases
    from conftest import load_golden
    g = load_golden('cpu_gpu_deposition_' + shape)
    sim = diag_cases.build_cpu_gpu_deposition(_namespace(), shape, str(tmp_path), fused=fused)
    sim.step(3)
    cpu, gpu = diag_cases.golden_files(g, 'cpu'), diag_cases.written_files(str(tmp_path))
    assert sorted(cpu) == sorted(gpu) == ['data%08d' % i for i in range(3)]
    for name in cpu:
        it = int(name[4:])
        for record, comps in (('rho', ('',)), ('J', ('/r', '/t', '/z'))):
            keys = ['/data/%d/fields/%s%s' % (it, record, co) for co in comps]
            tol = 1.e-13 * (max(np.abs(cpu[name][k]).max() for k in keys) + max(np.abs(gpu[name][k]).max() for k in keys))
            assert tol > 0
            for k in keys:
                assert cpu[name][k].shape == gpu[name][k].shape == (3, 50, 100)
                err = np.abs(cpu[name][k] - gpu[name][k]).max()
                assert err <= tol, '%s %s: %.3e > %.3e' % (name, k, err, tol)
        assert np.allclose(gpu[name]['/data/%d/fields/rho@gridGlobalOffset' % it],
                           cpu[name]['/data/%d/fields/rho@gridGlobalOffset' % it], rtol=1e-12, atol=1e-18)

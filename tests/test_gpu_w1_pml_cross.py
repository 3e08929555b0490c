"""GPU parity tests of the solver variants around the hot loop: radial PML
(`boundaries['r']='open'`) and the cross-deposition current correction -- Simulation.step on the B200
against golden outputs of the unmodified reference's CPU path (oracle/gen_golden_ext.py), fused and
unfused.  Tolerances as in test_gpu_step.py: fields 1e-9 of the field-group maximum, particles 1e-10."""
import numpy as np
import pytest
from scipy.constants import c

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu

STATE = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')


def _load_species(sp, g, check_generated=False):
    if check_generated:
        assert sp.Ntot == len(g['s0_in_x'])
    for k in STATE:
        if check_generated:
            assert_close(getattr(sp, k), g['s0_in_' + k], 1e-14, 'initial ' + k)
        setattr(sp, k, g['s0_in_' + k].copy())
    sp.Ntot = len(sp.x)
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(sp.Ntot))


def _check_particles(sp, g, tag, tol=1e-10):
    ref = np.stack([g['s0_out_' + k] for k in STATE])
    got = np.stack([getattr(sp, k) for k in STATE])
    assert got.shape == ref.shape, '%s: particle count %s vs %s' % (tag, got.shape, ref.shape)
    ro, go = np.lexsort((ref[2], ref[1], ref[0], ref[7])), np.lexsort((got[2], got[1], got[0], got[7]))
    for j, k in enumerate(STATE):
        assert_close(got[j][go], ref[j][ro], tol, '%s %s' % (tag, k))


def _check_fields(sim, g, tag, names, tol=1e-9):
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in names:
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], tol,
                         '%s %s m%d' % (tag, k, m), scale=sc)


FIELDS = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')
PML = ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['periodic', 'open', 'galilean', 'window'])
def test_pml_step_vs_reference_golden(tag, fused):
    """A tightly focused mode-1 pulse diffracts into the radial PML over a thin plasma: split components
    (push_eb_pml), radial + longitudinal damping, full transforms around them (main.py:410-415, 732-761);
    'galilean': v_comoving = 0.999 c with the Galilean PSATD; 'window': moving window at c + injection."""
    from fbpic_b200 import Simulation
    from scipy.constants import c
    g = load_golden('step_pml_' + tag)
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    zmax, rmax = float(g['zmax']), float(g['rmax'])
    open_z = bool(g['open_z'])
    V = float(g['v_comoving']) if bool(g['has_v']) else None
    np.random.seed(11)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, float(g['dt']),
                     p_zmin=(4.e-6 if open_z else 0.), p_zmax=(12.e-6 if open_z else zmax),
                     p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2, p_nt=4, n_e=5.e23, n_order=int(g['n_order']),
                     v_comoving=V, use_galilean=bool(g['use_galilean']),
                     n_guard=(12 if open_z else None), n_damp={'z': 12, 'r': 6},
                     boundaries={'z': ('open' if open_z else 'periodic'), 'r': 'open'}, fused=fused)
    if bool(g['window']):
        sim.set_moving_window(v=c)
    assert sim.use_pml and sim.comm.n_guard == int(g['n_guard'])
    assert sim.fld.interp[0].Nz == int(g['Nz_local']) and sim.fld.interp[0].Nr == int(g['Nr_local'])
    _load_species(sim.ptcl[0], g, check_generated=True)
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            getattr(sim.fld.interp[m], k)[:, :] = g['in_%s_m%d' % (k, m)]
    np.random.seed(12)
    sim.step(int(g['nsteps']))
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * abs(zmax)
    _check_particles(sim.ptcl[0], g, 'pml ' + tag, tol=(1e-9 if bool(g['window']) else 1e-10))
    _check_fields(sim, g, 'pml ' + tag, FIELDS + PML)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['std', 'galilean'])
def test_cross_deposition_step_vs_reference_golden(tag, fused):
    """current_correction='cross-deposition' (main.py:512-514, 672-717): two extra charge depositions at
    (z[n], x[n+1]) and (z[n+1], x[n]) per cycle and the cross-deposition correction kernel."""
    from fbpic_b200 import Simulation
    g = load_golden('step_cross_' + tag)
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    V = float(g['v_comoving']) if bool(g['has_v']) else None
    n_order = int(g['n_order'])
    sim = Simulation(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']), n_order=n_order,
                     n_guard=(None if n_order == -1 else 8), v_comoving=V, use_galilean=bool(g['use_galilean']),
                     current_correction='cross-deposition',
                     boundaries={'z': 'periodic', 'r': 'reflective'}, fused=fused)
    sp = sim.add_new_species(q=float(g['s0_q']), m=float(g['s0_m']))
    _load_species(sp, g)
    sim.step(int(g['nsteps']))
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * abs(float(g['zmax']))
    _check_particles(sp, g, 'cross ' + tag)
    _check_fields(sim, g, 'cross ' + tag, FIELDS)


# ------------------------------------------------------------------ the reference's tests/test_pml.py
def _pml_script(tmp_path, restart, **options):
    """tests/unautomated/test_pml.py as written (one domain): a tightly focused pulse (w0 = 1.5 micron) in modes 0
    and 1 diffracts into the radial PML over 40 microns; diagnostics every quarter, a checkpoint at half way."""
    from fbpic_b200 import Simulation
    from fbpic_b200.openpmd_diag import FieldDiagnostic, restart_from_checkpoint, set_periodic_checkpoint
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser, LaguerreGaussLaser
    Nz, zmin, zmax, Nr, Lr, Nm, n_order = 360, -6.e-6, 6.e-6, 50, 4.e-6, 2, 32
    w0, lambda0, tau, a0, zf, z0, L_prop = 1.5e-6, 0.8e-6, 10.e-15, 1., 0., 0., 40.e-6
    dt = (zmax - zmin) * 1. / c / Nz
    sim = Simulation(Nz=Nz, zmax=zmax, Nr=Nr, rmax=Lr, Nm=Nm, dt=dt, n_order=n_order, zmin=zmin, **options)
    profile0 = LaguerreGaussLaser(0, 1, 0.5 * a0, w0, tau, z0, zf=zf, lambda0=lambda0, theta_pol=0., theta0=0.) \
        + LaguerreGaussLaser(0, 1, 0.5 * a0, w0, tau, z0, zf=zf, lambda0=lambda0, theta_pol=np.pi / 2,
                             theta0=np.pi / 2)
    profile1 = GaussianLaser(a0=a0, waist=w0, tau=tau, lambda0=lambda0, z0=z0, zf=zf)
    if not restart:
        add_laser_pulse(sim, profile0)
        add_laser_pulse(sim, profile1)
    else:
        restart_from_checkpoint(sim, checkpoint_dir=str(tmp_path / 'checkpoints'))
    N_step = int(round(L_prop / (c * dt)))
    diag_period = int(round(N_step / 4))
    sim.diags = [FieldDiagnostic(diag_period, sim.fld, fieldtypes=["E"], comm=sim.comm,
                                 write_dir=str(tmp_path / 'diags'))]
    set_periodic_checkpoint(sim, N_step // 2, checkpoint_dir=str(tmp_path / 'checkpoints'))
    sim.step(N_step // 2 + 1)
    return profile0, profile1


@pytest.mark.parametrize('z_boundary,use_galilean', [('periodic', False), ('open', True)])
def test_pml_laser_as_written(z_boundary, use_galilean, tmp_path):
    """tests/test_pml.py::test_laser_periodic / test_laser_galilean (run_parallel + check_theory_pml) on one domain:
    run to half way, restart from the checkpoint, run to the end; at every diagnostic E_x of modes 0 and 1 inside the
    physical domain equals the analytic diffracting pulse within 9 % / 5 % of its maximum, i.e. nothing comes back
    from the radial boundary."""
    from fbpic_b200.diags import read_diag, list_iterations
    options = dict(boundaries={'z': z_boundary, 'r': 'open'})
    if use_galilean:
        options.update(use_galilean=True, v_comoving=0.999 * c)
    _pml_script(tmp_path, False, **options)
    assert list_iterations(str(tmp_path / 'diags')) == [0, 300, 600]
    profile0, profile1 = _pml_script(tmp_path, True, **options)
    iterations = list_iterations(str(tmp_path / 'diags'))
    assert iterations == [0, 300, 600, 900, 1200]
    for iteration in iterations:
        d = read_diag(str(tmp_path / 'diags'), iteration)
        Er = d['fields/E/r']
        t = float(d['time'])
        r = d['dr'] * (0.5 + np.arange(Er.shape[1]))
        z = d['zmin'] + d['dz'] * (0.5 + np.arange(Er.shape[2]))
        rr, zz = np.meshgrid(r, z, indexing='ij')
        # E_x of one mode in the half plane theta = 0 (what openPMD-viewer's get_field('E', 'x', m=m) returns for r > 0)
        for m, E_sim, profile, rtol in ((0, Er[0], profile0, 9.e-2), (1, Er[1], profile1, 5.e-2)):
            if z_boundary == 'periodic':
                Lz = d['dz'] * Er.shape[2]
                n_shift = np.floor(c * t / Lz)
                E_th = profile.E_field(rr, 0, zz + (n_shift + 1) * Lz, t)[0] + profile.E_field(rr, 0, zz + n_shift * Lz, t)[0]
            else:
                E_th = profile.E_field(rr, 0, zz, t)[0]
            relative_error = abs(E_sim - E_th).max() / abs(E_th).max()
            assert relative_error < rtol, (iteration, m, relative_error)

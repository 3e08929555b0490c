"""GPU parity tests of the particle-bunch utilities (fbpic/lpa_utils/bunch.py): bunch generation and the initial
space-charge field -- deposition by the regular kernel, transforms on the device -- against golden outputs of
the unmodified reference (oracle/gen_golden_ext.py)."""
import numpy as np
import pytest

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['uniform', 'gaussian', 'gaussian_boost'])
def test_bunch_space_charge_vs_reference_golden(tag):
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.bunch import add_particle_bunch, add_particle_bunch_gaussian
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from scipy.constants import e, m_e
    g = load_golden('bunch_' + tag)
    gb = float(g['gamma_boost']) or None
    np.random.seed(17)
    sim = Simulation(int(g['Nz']), float(g['zmax']), int(g['Nr']), float(g['rmax']), int(g['Nm']), float(g['dt']),
                     zmin=0., n_order=-1, n_guard=12, n_damp={'z': 10, 'r': 6}, gamma_boost=gb,
                     boundaries={'z': 'open', 'r': 'reflective'})
    boost = BoostConverter(gb) if gb is not None else None
    if bool(g['gaussian']):
        sp = add_particle_bunch_gaussian(sim, -e, m_e, sig_r=2.e-6, sig_z=1.5e-6, n_emit=1.e-6, gamma0=200.,
                                         sig_gamma=2., n_physical_particles=1.e8, n_macroparticles=2000,
                                         tf=10.e-15, zf=10.e-6, boost=boost, symmetrize=True)
    else:
        sp = add_particle_bunch(sim, -e, m_e, 100., 1.e23, 6.e-6, 12.e-6, 0., 5.e-6, boost=boost)
    # the bunch itself (host-side generation, same random draws as the reference); the deposition sorted it
    names = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
    ref = np.stack([g['s0_' + k] for k in names])
    got = np.stack([getattr(sp, k) for k in names])
    assert got.shape == ref.shape
    ro, go = np.lexsort((ref[2], ref[1], ref[0], ref[7])), np.lexsort((got[2], got[1], got[0], got[7]))
    for j, k in enumerate(names):
        assert_close(got[j][go], ref[j][ro], 1e-13, 'bunch %s %s' % (tag, k))
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-10,
                         'space charge %s %s m%d' % (tag, k, m), scale=group_scale(g, 'out_', k[0], Nm))


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'boost'])
def test_bunch_injection_plane_vs_reference_golden(tag, fused):
    """z_injection_plane: the bunch particles that have not yet crossed the plane keep their momenta (ballistic
    motion), the others feel the bunch's space-charge field; 8 cycles against the unmodified reference
    (push_p_after_plane, ballistic_before_plane.py); 'boost': gamma_boost = 3, the plane moves in the boosted frame."""
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.bunch import add_particle_bunch_gaussian
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from scipy.constants import e, m_e
    g = load_golden('bunch_plane_' + tag)
    gb = float(g['gamma_boost']) or None
    np.random.seed(19)
    sim = Simulation(int(g['Nz']), float(g['zmax']), int(g['Nr']), float(g['rmax']), int(g['Nm']), float(g['dt']),
                     zmin=0., n_order=-1, n_guard=12, n_damp={'z': 10, 'r': 6}, gamma_boost=gb,
                     boundaries={'z': 'open', 'r': 'reflective'}, fused=fused)
    boost = BoostConverter(gb) if gb is not None else None
    sp = add_particle_bunch_gaussian(sim, -e, m_e, sig_r=2.e-6, sig_z=1.5e-6, n_emit=1.e-6, gamma0=8., sig_gamma=0.5,
                                     n_physical_particles=5.e9, n_macroparticles=1200, zf=9.e-6, boost=boost,
                                     z_injection_plane=10.e-6)
    assert sp.ballistic_before_plane
    sim.step(int(g['nsteps']))
    names = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
    ref = np.stack([g['out_' + k] for k in names])
    got = np.stack([getattr(sp, k) for k in names])
    assert got.shape == ref.shape
    ro, go = np.lexsort((ref[2], ref[1], ref[0])), np.lexsort((got[2], got[1], got[0]))
    for j, k in enumerate(names):
        assert_close(got[j][go], ref[j][ro], 1e-10, 'bunch plane %s %s' % (tag, k))
    # some particles were pushed by the field, some were not (the plane cuts through the bunch)
    uin = np.sort(g['in_ux'])
    assert 0 < np.sum(np.isin(got[3], uin)) < sp.Ntot
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         'bunch plane %s %s m%d' % (tag, k, m), scale=sc)


def test_bunch_gaussian_as_written(tmp_path):
    """tests/test_space_charge.py::test_bunch_gaussian with tests/unautomated/test_space_charge_gaussian.py, as
    written (single domain): a symmetrized Gaussian bunch of 100000 particles (gamma = 15) with its space-charge
    field, one cycle with moving window and diagnostics; the transverse E and B read from the diagnostics file agree
    with the high-gamma theory of a Gaussian bunch within 10 % of the maximum, and the symmetrized bunch has zero
    mean transverse position and momentum."""
    from fbpic_b200 import Simulation
    from fbpic_b200.openpmd_diag import FieldDiagnostic, ParticleDiagnostic
    from fbpic_b200.lpa_utils.bunch import add_elec_bunch_gaussian
    from fbpic_b200.diags import read_diag
    from scipy.constants import c, epsilon_0
    np.random.seed(0)
    Nz, zmax, zmin, Nr, rmax, Nm, n_order = 400, 0.e-6, -40.e-6, 100, 100.e-6, 2, 32
    sig_r, sig_z, n_emit, gamma0, sig_gamma, Q, N, tf, zf = 3.e-6, 3.e-6, 1.e-6, 15., 1., 10.e-12, 100000, 0, -20.e-6
    dt = (zmax - zmin) / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, 0, 0, 0, 0, 2, 2, 4, 0., zmin=zmin, n_order=n_order,
                     boundaries={'z': 'open', 'r': 'reflective'})
    sim.set_moving_window(v=c)
    sim.ptcl = []
    elec = add_elec_bunch_gaussian(sim, sig_r, sig_z, n_emit, gamma0, sig_gamma, Q, N, tf, zf, symmetrize=True)
    sim.diags = [FieldDiagnostic(10, sim.fld, comm=sim.comm, write_dir=str(tmp_path)),
                 ParticleDiagnostic(10, species={'elec': elec}, comm=sim.comm, write_dir=str(tmp_path))]
    sim.step(1)
    d = read_diag(str(tmp_path), 0)
    # E_x, B_y in the plane theta = 0 / pi, as openPMD-viewer's get_field(..., theta=0) assembles them from the
    # thetaMode rows: x > 0: F_r(0) = m0 + Re m1; x < 0: -F_r(pi) = -(m0 - Re m1) (and B_y = B_t(0) resp. -B_t(pi))
    Er, Bt = d['fields/E/r'], d['fields/B/t']
    r = d['dr'] * (0.5 + np.arange(Nr))
    z = d['zmin'] + d['dz'] * (0.5 + np.arange(Er.shape[2]))
    rr, zz = np.meshgrid(r, z, indexing='ij')
    Eth = -Q / (2 * np.pi)**1.5 / sig_z / epsilon_0 / rr * (1 - np.exp(-0.5 * rr**2 / sig_r**2)) * \
        np.exp(-0.5 * (zz - zf)**2 / sig_z**2)
    for sign in (1., -1.):
        Ex = sign * (Er[0] + sign * Er[1])
        By = sign * (Bt[0] + sign * Bt[1])
        assert np.allclose(Ex, sign * Eth, atol=0.1 * np.abs(Eth).max())
        assert np.allclose(By, sign * Eth / c, atol=0.1 * np.abs(Eth).max() / c)
    assert np.abs(Er[0]).max() > 0.8 * np.abs(Eth).max()
    for key in ('position/x', 'position/y', 'momentum/x', 'momentum/y'):
        q = d['particles/elec/' + key]
        assert len(q) == N and abs(q.mean()) < 1.e-10 * q.std()


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_charge_cylinder_as_written(shape):
    """tests/test_charge_cylinder.py as written: an on-axis cylinder of charge shrunk down to 1 % of a radial cell
    (Ruyten-corrected shapes + corrected cell volumes near the axis); r E_r outside of it, from the deposited and
    filtered charge through `get_space_charge_spect`, equals the analytic lambda / (2 pi eps0) within 1e-3 for every
    radius.  (The reference works on host arrays; here the grid operations run inside `GpuMemoryManager` blocks and
    the element-wise space-charge solve on the host arrays in between.)"""
    from fbpic_b200 import Simulation, GpuMemoryManager, BinomialSmoother
    from fbpic_b200.lpa_utils.bunch import get_space_charge_spect
    from scipy.constants import c, e, epsilon_0
    Nz, zmax, zmin, Nr, rmax, Nm = 10, 10.e-6, -10.e-6, 20, 2.e-6, 1
    p_rmax, n_e = 1.e-6, 4.e18 * 1.e6
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, (zmax - zmin) / Nz / c, -100.e-6, 100.e-6, 0., p_rmax, 1, 8, 1, n_e,
                     zmin=zmin, boundaries={'z': 'periodic', 'r': 'reflective'}, verbose_level=0,
                     smoother=BinomialSmoother(1, False), particle_shape=shape)
    elec = sim.ptcl[0]
    for scale in [1.0, 0.5, 0.25, 0.1, 0.05, 0.025, 0.01]:
        elec.x *= scale
        elec.y *= scale
        with GpuMemoryManager(sim):
            sim.fld.erase('rho')
            sim.fld.erase('E')
            sim.fld.erase('B')
            sim.fld.interp2spect('E')
            sim.fld.interp2spect('B')
            elec.deposit(sim.fld, 'rho')
            sim.fld.sum_reduce_deposition_array('rho')
            sim.fld.divide_by_volume('rho')
            sim.fld.interp2spect('rho_prev')
            sim.fld.filter_spect('rho_prev')
        get_space_charge_spect(sim.fld.spect[0], 1)
        with GpuMemoryManager(sim):
            sim.fld.spect2interp('E')
            sim.fld.spect2interp('B')
            sim.fld.spect2interp('rho_prev')
        elec.x /= scale
        elec.y /= scale
        r = sim.fld.interp[0].r.copy()
        Er = sim.fld.interp[0].Er[5, :].real.copy()
        Er_theory = np.where(r < (p_rmax * scale),
                             r * n_e * e * np.pi * p_rmax**2 / (2 * np.pi * epsilon_0 * (p_rmax * scale)**2),
                             n_e * e * np.pi * p_rmax**2 / (2 * np.pi * epsilon_0 * r))
        assert np.abs(Er).max() > 0
        assert np.allclose((-Er * r)[-5:], (Er_theory * r)[-5:], 1.e-3), scale


def test_bunch_from_openpmd_series(tmp_path):
    """add_elec_bunch_openPMD: a bunch written by a ParticleDiagnostic is read back from the openPMD series into a
    second (boosted-frame) simulation -- same particles and same space-charge field as when the arrays are passed
    directly (`add_elec_bunch_from_arrays`), including the selection and the re-centring at z_off."""
    from fbpic_b200 import Simulation
    from fbpic_b200.openpmd_diag import ParticleDiagnostic
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from fbpic_b200.lpa_utils.bunch import add_elec_bunch_gaussian, add_elec_bunch_openPMD, add_elec_bunch_from_arrays
    from scipy.constants import c
    Nz, zmax, zmin, Nr, rmax, Nm = 64, 0., -32.e-6, 24, 24.e-6, 2
    dt = (zmax - zmin) / Nz / c
    kw = dict(zmin=zmin, n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6}, boundaries={'z': 'open', 'r': 'reflective'})
    np.random.seed(3)
    a = Simulation(Nz, zmax, Nr, rmax, Nm, dt, **kw)
    a.ptcl = []
    src = add_elec_bunch_gaussian(a, 2.e-6, 2.e-6, 1.e-6, 40., 2., 20.e-12, 5000, zf=-16.e-6)
    a.diags = [ParticleDiagnostic(1, {'beam': src}, a.comm, write_dir=str(tmp_path))]
    x, y, z, ux, uy, uz, w = (np.array(getattr(src, k)) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w'))
    a.step(1)                                   # writes iteration 0: the bunch as initialised
    keep = uz > 38.
    z_off = -20.e-6
    z_ref = z[keep] - np.average(z[keep], weights=w[keep]) + z_off
    sims = []
    for loader in ('openPMD', 'arrays'):
        b = Simulation(Nz, zmax, Nr, rmax, Nm, dt, gamma_boost=4., **kw)
        b.ptcl = []
        if loader == 'openPMD':
            sp = add_elec_bunch_openPMD(b, str(tmp_path / 'hdf5'), z_off=z_off, species='beam',
                                        select={'uz': [38., None]}, iteration=0, boost=BoostConverter(4.))
        else:
            sp = add_elec_bunch_from_arrays(b, x[keep], y[keep], z_ref, ux[keep], uy[keep], uz[keep], w[keep],
                                            boost=BoostConverter(4.))
        sims.append((b, sp))
    (b1, s1), (b2, s2) = sims
    assert s1.Ntot == s2.Ntot and 100 < s1.Ntot < 5000
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w', 'inv_gamma'):
        assert_close(np.array(getattr(s1, k)), np.array(getattr(s2, k)), 1e-12, 'openPMD bunch ' + k)
    for m in range(Nm):
        for k in ('Er', 'Ez', 'Bt'):
            scale = np.abs(getattr(b2.fld.interp[0], 'Er')).max()
            assert scale > 0
            assert_close(getattr(b1.fld.interp[m], k), getattr(b2.fld.interp[m], k), 1e-9, 'openPMD bunch %s m%d' % (k, m),
                         scale=scale * (1. if k[0] == 'E' else 1. / c))

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

"""GPU parity tests of laser initialisation and emission against golden outputs of the
unmodified reference (oracle/gen_golden_ext.py): `add_laser_pulse` direct injection (profile sampled on the
grid, Ez / B built in spectral space with the device transforms) and the laser antenna (virtual particles
deposited by the regular deposition kernel every step)."""
import pytest

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu


def _sim(g, **kw):
    from fbpic_b200 import Simulation
    gb = float(g['gamma_boost'])
    return Simulation(int(g['Nz']), float(g['zmax']), int(g['Nr']), float(g['rmax']), int(g['Nm']), float(g['dt']),
                      zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6},
                      gamma_boost=(gb if gb else None), **kw)


@pytest.mark.parametrize('tag', ['gauss', 'lg_pml', 'boost'])
def test_add_laser_direct_vs_reference_golden(tag):
    """direct_injection.py:12-217; 'lg_pml': Laguerre-Gauss (0,1) on a grid with radial PML cells;
    'boost': gamma_boost = 3 (profile given in the lab frame)."""
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser, LaguerreGaussLaser
    g = load_golden('laser_direct_' + tag)
    sim = _sim(g, boundaries={'z': 'open', 'r': ('open' if bool(g['pml']) else 'reflective')})
    assert sim.fld.interp[0].Nz == int(g['Nz_local'])
    if bool(g['lg']):
        prof = LaguerreGaussLaser(0, 1, a0=1., waist=4.e-6, tau=8.e-15, z0=12.e-6, zf=20.e-6, theta_pol=0.4)
    else:
        prof = GaussianLaser(a0=2., waist=4.e-6, tau=8.e-15, z0=12.e-6, zf=25.e-6, theta_pol=0.7,
                             lambda0=1.6e-6, cep_phase=0.3)
    gb = float(g['gamma_boost'])
    add_laser_pulse(sim, prof, gamma_boost=(gb if gb else None))
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-11,
                         'laser %s %s m%d' % (tag, k, m), scale=group_scale(g, 'out_', k[0], Nm))


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'moving', 'boost'])
def test_laser_antenna_vs_reference_golden(tag, fused):
    """antenna_injection.py:24-442: emission into an empty open-z box; 'moving': antenna moving at 0.2 c with
    the cross-deposition correction (the antenna takes part in cross_deposit, main.py:688-713);
    'boost': gamma_boost = 2."""
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser
    g = load_golden('laser_antenna_' + tag)
    sim = _sim(g, boundaries={'z': 'open', 'r': 'reflective'}, fused=fused,
               current_correction=('cross-deposition' if bool(g['cross']) else 'curl-free'))
    prof = GaussianLaser(a0=1., waist=3.e-6, tau=6.e-15, z0=-4.e-6, zf=10.e-6, theta_pol=0.5, lambda0=1.6e-6)
    gb = float(g['gamma_boost'])
    add_laser_pulse(sim, prof, gamma_boost=(gb if gb else None), method='antenna', z0_antenna=6.e-6,
                    v_antenna=float(g['v_antenna']))
    ant = sim.laser_antennas[0]
    assert_close(ant.w, g['w'], 1e-14, 'antenna weights')
    assert abs(ant.mobility_coef - float(g['mobility_coef'])) <= 1e-14 * abs(float(g['mobility_coef']))
    sim.step(int(g['nsteps']))
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * float(g['zmax'])
    assert_close(ant.baseline_z, g['baseline_z'], 1e-13, 'antenna z')
    assert_close(ant.vx, g['vx'], 1e-11, 'antenna vx')
    assert_close(ant.excursion_x, g['excursion_x'], 1e-11, 'antenna excursion x')
    assert_close(ant.excursion_y, g['excursion_y'], 1e-11, 'antenna excursion y')
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         'antenna %s %s m%d' % (tag, k, m), scale=sc)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'pml', 'boost'])
def test_mirror_vs_reference_golden(tag, fused):
    """mirrors.py:10-94: a Gaussian pulse runs into a mirror (E, B zeroed in a slab every cycle) and is reflected;
    'pml': radial PML, only mode 1 is zeroed; 'boost': mirror position given in the lab frame, gamma_boost = 2."""
    from fbpic_b200.lpa_utils.laser import add_laser_pulse, GaussianLaser
    from fbpic_b200.lpa_utils.mirrors import Mirror
    g = load_golden('mirror_' + tag)
    gb = float(g['gamma_boost']) or None
    pml = bool(g['pml'])
    sim = _sim(g, boundaries={'z': 'open', 'r': ('open' if pml else 'reflective')}, fused=fused)
    add_laser_pulse(sim, GaussianLaser(a0=1., waist=4.e-6, tau=6.e-15, z0=8.e-6, lambda0=1.6e-6, theta_pol=0.4),
                    gamma_boost=gb)
    sim.mirrors = [Mirror(16.e-6, 17.5e-6, gamma_boost=gb, m=('all' if not pml else [1]))]
    sim.step(int(g['nsteps']))
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         'mirror %s %s m%d' % (tag, k, m), scale=group_scale(g, 'out_', k[0], Nm))

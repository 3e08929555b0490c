"""CPU tests of the host-side set-up utilities (fbpic_b200/lpa_utils): laser profiles and the boosted-frame
converter against values produced by the unmodified reference (oracle/gen_golden_ext.py)."""
import numpy as np
from scipy.constants import c

from conftest import load_golden, assert_close


def test_laser_profiles_vs_reference():
    from fbpic_b200.lpa_utils.laser import GaussianLaser, LaguerreGaussLaser, DonutLikeLaguerreGaussLaser, \
        FlattenedGaussianLaser, FewCycleLaser
    g = load_golden('laser_profiles')
    x, y, z, t = g['x'], g['y'], g['z'], float(g['t'])
    profs = {
        'gauss': GaussianLaser(a0=2., waist=5.e-6, tau=20.e-15, z0=10.e-6, zf=30.e-6, theta_pol=0.3,
                               lambda0=0.8e-6, cep_phase=0.4, phi2_chirp=150.e-30),
        'gauss_bw': GaussianLaser(a0=1., waist=4.e-6, tau=15.e-15, z0=5.e-6, propagation_direction=-1),
        'lg11': LaguerreGaussLaser(1, 1, a0=1.5, waist=6.e-6, tau=18.e-15, z0=8.e-6, zf=-5.e-6, theta_pol=1.1,
                                   cep_phase=0.2, theta0=0.5),
        'lg20': LaguerreGaussLaser(2, 0, a0=0.7, waist=5.e-6, tau=25.e-15, z0=0.),
        'donut12': DonutLikeLaguerreGaussLaser(1, 2, a0=1.2, waist=6.e-6, tau=18.e-15, z0=8.e-6, zf=20.e-6,
                                               theta_pol=0.6, cep_phase=0.1),
        'donut0m1': DonutLikeLaguerreGaussLaser(0, -1, a0=1., waist=5.e-6, tau=12.e-15, z0=3.e-6,
                                                propagation_direction=-1),
        'flat': FlattenedGaussianLaser(a0=1.3, w0=5.e-6, tau=20.e-15, z0=5.e-6, N=5, zf=60.e-6, theta_pol=0.2,
                                       cep_phase=0.3),
        'fewcycle': FewCycleLaser(a0=2., waist=3.e-6, tau_fwhm=5.e-15, z0=10.e-6, zf=14.e-6, theta_pol=0.9,
                                  cep_phase=0.7),
    }
    for k, p in profs.items():
        Ex, Ey = p.E_field(x, y, z, t)
        assert_close(Ex, g[k + '_Ex'], 1e-13, k + ' Ex')
        assert_close(Ey, g[k + '_Ey'], 1e-13, k + ' Ey')
    Ex, Ey = (profs['gauss'] + profs['lg11']).E_field(x, y, z, t)
    assert_close(Ex, g['sum_Ex'], 1e-13, 'sum Ex')
    assert_close(Ey, g['sum_Ey'], 1e-13, 'sum Ey')


def test_paraxial_profiles_vs_reference():
    """ParaxialApproximationLaser (longitudinal x transverse profile, amplitude from the pulse energy) with the
    Gaussian-chirped and measured-spectrum longitudinal profiles and all transverse profiles, against the reference;
    the spectrum is the reference's own fixture tests/laser_spectrum.csv."""
    import os
    from conftest import GOLDEN
    from fbpic_b200.lpa_utils.laser import ParaxialApproximationLaser, GaussianChirpedLongitudinalProfile, \
        GaussianTransverseProfile, FlattenedGaussianTransverseProfile, DonutLikeLaguerreGaussTransverseProfile, \
        LaguerreGaussTransverseProfile, CustomSpectrumLongitudinalProfile
    g = load_golden('laser_profiles')
    x, y, z, t = g['x'], g['y'], g['z'], float(g['t'])
    custom = CustomSpectrumLongitudinalProfile(z0=4.e-6, spectrum_file=os.path.join(GOLDEN, 'laser_spectrum.csv'),
                                               phi2_chirp=80.e-30, phi3_chirp=2.e-42, subtract_linear_phase=True)
    assert abs(custom.get_mean_wavelength() / float(g['custom_lambda0']) - 1) < 1e-13
    assert abs(custom.squared_profile_integral() / float(g['custom_integral']) - 1) < 1e-12
    chirped = GaussianChirpedLongitudinalProfile(tau=17.e-15, z0=6.e-6, cep_phase=0.3, phi2_chirp=200.e-30)
    parax = {
        'parax_custom': ParaxialApproximationLaser(custom, GaussianTransverseProfile(
            waist=6.e-6, zf=25.e-6, lambda0=custom.get_mean_wavelength()), 0.7, theta_pol=0.2),
        'parax_gauss': ParaxialApproximationLaser(chirped, GaussianTransverseProfile(waist=5.e-6, zf=30.e-6), 1.),
        'parax_flat': ParaxialApproximationLaser(chirped, FlattenedGaussianTransverseProfile(w0=5.e-6, N=8, zf=50.e-6),
                                                 0.5, theta_pol=1.),
        'parax_donut': ParaxialApproximationLaser(chirped, DonutLikeLaguerreGaussTransverseProfile(
            waist=6.e-6, zf=10.e-6, p=2, m=1), 2.),
        'parax_lg': ParaxialApproximationLaser(chirped, LaguerreGaussTransverseProfile(1, 2, 6.e-6, zf=-8.e-6,
                                                                                       theta0=0.4), 1.5),
    }
    for k, p in parax.items():
        Ex, Ey = p.E_field(x, y, z, t)
        tol = 1e-9 if k == 'parax_custom' else 1e-13          # interpolated from a 10^6-point FFT of the spectrum
        assert_close(Ex, g[k + '_Ex'], tol, k + ' Ex')
        assert_close(Ey, g[k + '_Ey'], tol, k + ' Ey')


def test_boost_converter():
    """Lorentz identities of fbpic/lpa_utils/boosted_frame.py."""
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    b = BoostConverter(10.)
    assert abs(b.beta0 - np.sqrt(1 - 0.01)) < 1e-15
    L, = b.static_length([2.])
    assert L == 2. / 10.
    Lc, = b.copropag_length([1.], beta_object=1.)
    assert abs(Lc - 1. / (10. * (1 - b.beta0))) < 1e-12 * Lc
    n1, = b.static_density([3.])
    n2, = b.copropag_density([3.], beta_object=0.5)
    assert n1 == 30. and abs(n2 - 3. * 10. * (1 - 0.5 * b.beta0)) < 1e-13
    v, = b.velocity([c])
    assert abs(v - c) < 1e-14 * c / (1 - b.beta0)   # light stays light (cancellation in 1 - beta0)
    v0, = b.velocity([0.])
    assert abs(v0 + b.beta0 * c) < 1e-7           # a lab-frame object at rest moves backwards
    uz, = b.longitudinal_momentum([0.])
    assert abs(uz + 10. * b.beta0) < 1e-13
    gm, = b.gamma([1.])
    assert abs(gm - 10.) < 1e-13
    k, = b.wavenumber([1.])
    assert abs(k - 1. / (10. * (1 + b.beta0))) < 1e-15
    # a particle at rest at z: after the boost it sits at z/gamma0 at t' = 0 and moves with -beta0 c
    x, y, z = np.zeros(3), np.zeros(3), np.array([1., 2., 3.])
    u0 = np.zeros(3)
    nx, ny, nz, nux, nuy, nuz, nig = b.boost_particle_arrays(x, y, z, u0, u0, u0, np.ones(3))
    assert np.allclose(nz, z / 10., rtol=1e-12) and np.allclose(nuz, -10. * b.beta0) and np.allclose(nig, 0.1)
    T = b.interaction_time(1.e-3, 50.e-6, c)
    Li, lw, vw = 1.e-3 / 10., 50.e-6 / (10. * (1 - b.beta0)), c
    assert abs(T - (Li + lw) / (vw + b.beta0 * c)) < 1e-12 * T


def test_host_tables_option_variants_bit_exact():
    """Smoother passes / compensator, use_ruyten_shapes=False, use_modified_volume=False: the host tables of
    `Simulation` equal the reference's bit for bit (smoothing.py:57-94, interpolation_grid.py:88-138)."""
    from fbpic_b200 import Simulation, BinomialSmoother
    g = load_golden('tables_variants')
    Nr, Nz, rmax, dz, dt = int(g['Nr']), int(g['Nz']), float(g['rmax']), float(g['dz']), float(g['dt'])
    cases = {'p2': dict(n_passes=2, compensator=False), 'p1c': dict(n_passes=1, compensator=True),
             'mixed': dict(n_passes={'z': 3, 'r': 1}, compensator={'z': True, 'r': False})}
    for tag, kw in cases.items():
        sim = Simulation(Nz, Nz * dz, Nr, rmax, 2, dt, zmin=0., smoother=BinomialSmoother(**kw))
        for m in range(2):
            assert np.array_equal(sim.fld.spect[m].filter_array_z, g['%s_fz_m%d' % (tag, m)]), (tag, m)
            assert np.array_equal(sim.fld.spect[m].filter_array_r, g['%s_fr_m%d' % (tag, m)]), (tag, m)
    for tag, kw in {'noruyten': dict(use_ruyten_shapes=False), 'novol': dict(use_modified_volume=False),
                    'neither': dict(use_ruyten_shapes=False, use_modified_volume=False)}.items():
        sim = Simulation(Nz, Nz * dz, Nr, rmax, 3, dt, zmin=0., **kw)
        for m in range(3):
            gr = sim.fld.interp[m]
            assert np.array_equal(gr.invvol, g['%s_invvol_m%d' % (tag, m)]), (tag, m)
            assert np.array_equal(gr.ruyten_linear_coef, g['%s_lin_m%d' % (tag, m)]), (tag, m)
            assert np.array_equal(gr.ruyten_cubic_coef, g['%s_cub_m%d' % (tag, m)]), (tag, m)


def test_adk_tables_vs_reference():
    """Per-level ADK tables (prefactor, power, exponential prefactor) of `Ionizer.initialize_ADK_parameters` for H,
    He, N, Ar, Kr against the reference's (ionizer.py:137-183), which also pins the ionization energies."""
    import types
    from fbpic_b200.ionization import Ionizer, get_ionization_energies
    g = load_golden('ionization')
    for element in ('H', 'He', 'N', 'Ar', 'Kr'):
        ion = types.SimpleNamespace(level_max=None)
        Ionizer.initialize_ADK_parameters(ion, element, float(g['dt']))
        for name, got in (('prefactor', ion.adk_prefactor), ('power', ion.adk_power),
                          ('exp_prefactor', ion.adk_exp_prefactor)):
            ref = g['%s_%s' % (element, name)]
            assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-12, atol=0), (element, name)
    assert get_ionization_energies('Xx') is None and len(get_ionization_energies('Ar')) == 18


def _write_lasy_like(path, data, geometry, spacing, offset, omega, pol, version='0.4.0'):
    """a file with the layout `lasy` writes (openPMD `laserEnvelope` mesh), through fbpic_b200's own container"""
    from fbpic_b200.openpmd_store import File
    f = File(path, 'w')
    f.attrs['software'], f.attrs['softwareVersion'] = np.bytes_('lasy'), np.bytes_(version)
    d = f.create_dataset('/data/0/meshes/laserEnvelope', data=data)
    d.attrs['angularFrequency'], d.attrs['polarization'], d.attrs['geometry'] = omega, pol, np.bytes_(geometry)
    d.attrs['gridSpacing'], d.attrs['gridGlobalOffset'], d.attrs['gridUnitSI'] = spacing, offset, 1.
    f.close()


def test_lasy_file_laser_vs_reference(tmp_path):
    """FromLasyFileLaser on a thetaMode (modes 0, 1) and a cartesian envelope file against the reference's class on
    the same files (points inside and outside of the tables, in space and time); obsolete files are refused."""
    import pytest
    from fbpic_b200.lpa_utils.laser import FromLasyFileLaser
    g = load_golden('lasy_laser')
    omega, pol = float(g['omega']), g['pol']
    x, y, t = g['x'], g['y'], g['t']
    cases = {'theta': (g['theta_data'], 'thetaMode', np.array([1.e-15, 1.e-6]), np.array([-25.e-15, 0.])),
             'cart': (g['cart_data'], 'cartesian', np.array([1.e-15, 2.e-6, 1.5e-6]),
                      np.array([-25.e-15, -23.e-6, -20.e-6]))}
    for tag, (data, geometry, spacing, offset) in cases.items():
        path = str(tmp_path / (tag + '.npz'))
        _write_lasy_like(path, data, geometry, spacing, offset, omega, pol)
        Ex, Ey = FromLasyFileLaser(path, t_start=4.e-15).E_field(x, y, 0. * x, t)
        assert np.count_nonzero(g[tag + '_Ex']) > 100 and np.count_nonzero(g[tag + '_Ex'] == 0) > 20
        assert_close(Ex, g[tag + '_Ex'], 1e-13, tag + ' Ex')
        assert_close(Ey, g[tag + '_Ey'], 1e-13, tag + ' Ey')
    old = str(tmp_path / 'old.npz')
    _write_lasy_like(old, *cases['theta'], omega, pol, version='0.2.1')
    with pytest.raises(RuntimeError):
        FromLasyFileLaser(old)

"""
oracle/gen_golden_ext.py -- fixtures for the solver variants around the hot loop:
radial PML, cross-deposition, laser antenna, external fields, boosted-frame set-up.  Like
oracle/gen_golden.py it runs the UNMODIFIED FBPIC reference (CPU/numba path, through
oracle/ref_shim) and only in the build container:

    NUMBA_THREADING_LAYER=omp OPENBLAS_NUM_THREADS=1 NUMBA_NUM_THREADS=4 \
        python oracle/gen_golden_ext.py [--only NAME]

Test infrastructure.
"""
import math
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import save, ptcl_arrays, field_arrays, impart_momenta   # noqa: E402  (sets sys.path)
from scipy.constants import c, e, m_e                                      # noqa: E402
from fbpic.main import Simulation                                          # noqa: E402

PML_NAMES = ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')


def diverging_mode1_field(sim, z0, w0, amp, lam=1.6e-6, ctau=2.5e-6):
    """A short, tightly focused pulse in mode 1 (x-polarised-laser structure, restated from
    gen_golden.gen_window): it diffracts into the radial boundary within a few steps."""
    g1 = sim.fld.interp[1]
    zz, rr = np.meshgrid(g1.z, g1.r, indexing='ij')
    prof = amp * np.exp(-(zz - z0)**2 / ctau**2) * np.exp(-rr**2 / w0**2) * np.cos(2 * np.pi * (zz - z0) / lam)
    g1.Er[:, :], g1.Et[:, :] = 0.5 * prof, -0.5j * prof
    g1.Br[:, :], g1.Bt[:, :] = 0.5j * prof / c, 0.5 * prof / c


def gen_pml(tag, open_z, v_comoving=None, use_galilean=False, nsteps=6, window=False):
    np.random.seed(11)
    Nz, Nr, Nm, zmax, rmax = 32, 10, 2, 16.e-6, 6.e-6
    dt = zmax / Nz / c
    n_order = 8 if (v_comoving is not None) else -1
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=(4.e-6 if open_z else 0.), p_zmax=(12.e-6 if open_z else zmax),
                     p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2, p_nt=4, n_e=5.e23, n_order=n_order,
                     v_comoving=v_comoving, use_galilean=use_galilean, verbose_level=0,
                     n_guard=(12 if open_z else None), n_damp={'z': 12, 'r': 6},
                     boundaries={'z': ('open' if open_z else 'periodic'), 'r': 'open'})
    if window:
        sim.set_moving_window(v=c)
    assert sim.use_pml
    diverging_mode1_field(sim, 0.5 * zmax, 1.8e-6, 3.e11)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, n_order=n_order, nsteps=nsteps, open_z=open_z,
               nr_damp=6, nz_damp=12, n_guard=sim.comm.n_guard, window=window,
               v_comoving=(0. if v_comoving is None else v_comoving), has_v=(v_comoving is not None),
               use_galilean=use_galilean, Nz_local=sim.fld.interp[0].Nz, Nr_local=sim.fld.interp[0].Nr)
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['s0_q'], out['s0_m'] = sp.q, sp.m
    out.update({'in_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    np.random.seed(12)
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    for m in range(Nm):
        for k in PML_NAMES:
            out['out_%s_m%d' % (k, m)] = getattr(sim.fld.interp[m], k).copy()
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_pml_' + tag, **out)


def gen_cross(tag, v_comoving, use_galilean, nsteps=3):
    """Periodic plasma wave with current_correction='cross-deposition' (main.py:512-514, 672-717)."""
    np.random.seed(0)
    Nz, Nr, Nm, zmax, rmax = 24, 12, 2, 12.e-6, 8.e-6
    dt = zmax / Nz / c
    n_e = 2.e24
    n_order = -1 if v_comoving is None else 16
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4, n_e=n_e, n_order=n_order, n_guard=(None if n_order == -1 else 8),
                     v_comoving=v_comoving, use_galilean=use_galilean, verbose_level=0,
                     current_correction='cross-deposition',
                     boundaries={'z': 'periodic', 'r': 'reflective'})
    k0 = 2 * np.pi / zmax * 2
    wp = np.sqrt(n_e * e**2 / (m_e * 8.8541878128e-12))
    impart_momenta(sim.ptcl[0], 0.05, k0, 3.e-6, wp)
    if v_comoving is not None:
        g = 1. / np.sqrt(1 - (v_comoving / c)**2)
        for sp in sim.ptcl:
            sp.uz += -np.sqrt(g**2 - 1)
            sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, n_order=n_order, nsteps=nsteps,
               v_comoving=(0. if v_comoving is None else v_comoving), has_v=(v_comoving is not None),
               use_galilean=use_galilean, n_species=1, open_z=False, Nz_local=Nz)
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['s0_q'], out['s0_m'] = sp.q, sp.m
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_cross_' + tag, **out)


def gen_laser_profiles():
    """E_field of the analytic profiles on random points (laser_profiles.py)."""
    from fbpic.lpa_utils.laser import GaussianLaser, LaguerreGaussLaser, DonutLikeLaguerreGaussLaser, \
        FlattenedGaussianLaser, FewCycleLaser
    rng = np.random.default_rng(21)
    n = 400
    x, y = rng.normal(size=n) * 6.e-6, rng.normal(size=n) * 6.e-6
    z = rng.uniform(-20.e-6, 40.e-6, n)
    t = 13.e-15
    out = dict(x=x, y=y, z=z, t=t)
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
        out[k + '_Ex'], out[k + '_Ey'] = p.E_field(x, y, z, t)
    s = profs['gauss'] + profs['lg11']
    out['sum_Ex'], out['sum_Ey'] = s.E_field(x, y, z, t)
    # longitudinal x transverse profiles normalised to a pulse energy (laser_profiles.py:105-176); the measured
    # spectrum is the reference's own fixture tests/laser_spectrum.csv (copied to tests/golden/); np.trapz is gone
    # from NumPy 2
    if not hasattr(np, 'trapz'):
        np.trapz = np.trapezoid
    from fbpic.lpa_utils.laser import ParaxialApproximationLaser, GaussianChirpedLongitudinalProfile, \
        GaussianTransverseProfile, FlattenedGaussianTransverseProfile, DonutLikeLaguerreGaussTransverseProfile, \
        LaguerreGaussTransverseProfile, CustomSpectrumLongitudinalProfile
    spectrum = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'laser_spectrum.csv')
    custom = CustomSpectrumLongitudinalProfile(z0=4.e-6, spectrum_file=spectrum, phi2_chirp=80.e-30, phi3_chirp=2.e-42,
                                               subtract_linear_phase=True)
    out['custom_lambda0'] = custom.get_mean_wavelength()
    out['custom_integral'] = custom.squared_profile_integral()
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
        out[k + '_Ex'], out[k + '_Ey'] = p.E_field(x, y, z, t)
    save('laser_profiles', **out)


def gen_laser_direct(tag, gamma_boost=None, lg=False, pml=False):
    """add_laser_pulse(method='direct') on an open-z box: the fields put on the grid
    (direct_injection.py:12-217)."""
    from fbpic.lpa_utils.laser import add_laser_pulse, GaussianLaser, LaguerreGaussLaser
    Nz, Nr, Nm, zmax, rmax = 48, 16, 2, 24.e-6, 16.e-6
    zmin = 0.
    dt = (zmax - zmin) / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6},
                     gamma_boost=gamma_boost, verbose_level=0,
                     boundaries={'z': 'open', 'r': ('open' if pml else 'reflective')})
    if lg:
        prof = LaguerreGaussLaser(0, 1, a0=1., waist=4.e-6, tau=8.e-15, z0=12.e-6, zf=20.e-6, theta_pol=0.4)
    else:
        prof = GaussianLaser(a0=2., waist=4.e-6, tau=8.e-15, z0=12.e-6, zf=25.e-6, theta_pol=0.7,
                             lambda0=1.6e-6, cep_phase=0.3)
    add_laser_pulse(sim, prof, gamma_boost=gamma_boost)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, zmin=zmin, dt=dt, lg=lg, pml=pml,
               gamma_boost=(0. if gamma_boost is None else gamma_boost), Nz_local=sim.fld.interp[0].Nz)
    out.update({'out_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    save('laser_direct_' + tag, **out)


def gen_laser_antenna(tag, gamma_boost=None, v_antenna=0., nsteps=24, cross=False):
    """A laser emitted by an antenna into an empty open-z box (antenna_injection.py:24-442):
    fields after nsteps and the state of the antenna."""
    from fbpic.lpa_utils.laser import add_laser_pulse, GaussianLaser
    Nz, Nr, Nm, zmax, rmax = 48, 12, 2, 24.e-6, 12.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6},
                     gamma_boost=gamma_boost, verbose_level=0,
                     current_correction=('cross-deposition' if cross else 'curl-free'),
                     boundaries={'z': 'open', 'r': 'reflective'})
    prof = GaussianLaser(a0=1., waist=3.e-6, tau=6.e-15, z0=-4.e-6, zf=10.e-6, theta_pol=0.5, lambda0=1.6e-6)
    add_laser_pulse(sim, prof, gamma_boost=gamma_boost, method='antenna', z0_antenna=6.e-6, v_antenna=v_antenna)
    sim.step(nsteps, show_progress=False)
    ant = sim.laser_antennas[0]
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, cross=cross,
               gamma_boost=(0. if gamma_boost is None else gamma_boost), v_antenna=v_antenna,
               Nz_local=sim.fld.interp[0].Nz, excursion_x=ant.excursion_x, excursion_y=ant.excursion_y,
               baseline_z=ant.baseline_z, vx=ant.vx, vy=ant.vy, w=ant.w, mobility_coef=ant.mobility_coef,
               zmin_end=sim.fld.interp[0].zmin)
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    save('laser_antenna_' + tag, **out)


def undulator_field(F, x, y, z, t, amplitude, length_scale):
    return F + amplitude * math.cos(2 * np.pi * z / length_scale)


def focusing_field(F, x, y, z, t, amplitude, length_scale):
    k = 2 * math.pi / length_scale
    if z > 2.e-6:
        g = math.exp(-(x**2 + y**2) / length_scale**2)
    else:
        g = 0.
    return F - amplitude * k * x * g * math.sin(k * (z - 299792458. * t))


def gen_external(tag, gamma_boost=None, nsteps=4):
    """Plasma electrons in external fields (external_fields.py:13-215, main.py:472-473)."""
    from fbpic.lpa_utils.external_fields import ExternalField
    np.random.seed(3)
    Nz, Nr, Nm, zmax, rmax = 24, 12, 2, 12.e-6, 8.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4, n_e=1.e23, n_order=-1, gamma_boost=gamma_boost, verbose_level=0,
                     boundaries={'z': 'periodic', 'r': 'reflective'})
    sim.external_fields = [
        ExternalField(undulator_field, 'By', 40., 5.e-6, gamma_boost=gamma_boost),
        ExternalField(focusing_field, 'Ex', 3.e10, 4.e-6, species=sim.ptcl[0], gamma_boost=gamma_boost)]
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps,
               gamma_boost=(0. if gamma_boost is None else gamma_boost))
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['s0_q'], out['s0_m'] = sp.q, sp.m
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_external_' + tag, **out)


def gen_bunch(tag, gaussian=False, gamma_boost=None):
    """Relativistic bunch + its space-charge field on the grid (bunch.py:18-1007)."""
    from fbpic.lpa_utils.bunch import add_particle_bunch, add_particle_bunch_gaussian
    from fbpic.lpa_utils.boosted_frame import BoostConverter
    np.random.seed(17)
    Nz, Nr, Nm, zmax, rmax = 40, 16, 2, 20.e-6, 16.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 10, 'r': 6},
                     gamma_boost=gamma_boost, verbose_level=0, boundaries={'z': 'open', 'r': 'reflective'})
    boost = BoostConverter(gamma_boost) if gamma_boost is not None else None
    if gaussian:
        sp = add_particle_bunch_gaussian(sim, -e, m_e, sig_r=2.e-6, sig_z=1.5e-6, n_emit=1.e-6, gamma0=200.,
                                         sig_gamma=2., n_physical_particles=1.e8, n_macroparticles=2000,
                                         tf=10.e-15, zf=10.e-6, boost=boost, symmetrize=True)
    else:
        sp = add_particle_bunch(sim, -e, m_e, 100., 1.e23, 6.e-6, 12.e-6, 0., 5.e-6, boost=boost)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, gaussian=gaussian,
               gamma_boost=(0. if gamma_boost is None else gamma_boost))
    out.update({'s0_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    save('bunch_' + tag, **out)


def gen_script(tag):
    """The scaled-down documented input scripts of tests/script_cases.py, run by the reference."""
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import script_cases
    from fbpic.lpa_utils.laser import add_laser_pulse, GaussianLaser
    from fbpic.lpa_utils.bunch import add_particle_bunch
    from fbpic.lpa_utils.boosted_frame import BoostConverter
    ns = types.SimpleNamespace(Simulation=Simulation, add_laser_pulse=add_laser_pulse, GaussianLaser=GaussianLaser,
                               add_particle_bunch=add_particle_bunch, BoostConverter=BoostConverter)
    np.random.seed(31)
    sim, species, nsteps = script_cases.CASES[tag](ns, verbose_level=0)
    out = dict(nsteps=nsteps, Nm=sim.fld.Nm, Nz_local=sim.fld.interp[0].Nz, dt=sim.dt,
               species=np.array(sorted(species)))
    out.update({'%s_n_in' % name: sp.Ntot for name, sp in species.items()})
    np.random.seed(32)
    sim.step(nsteps, show_progress=False)
    for name, sp in species.items():
        out.update({'%s_out_%s' % (name, k): v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    out['time_end'] = sim.time
    save('script_' + tag, **out)


def gen_tables_variants():
    """Host tables for the non-default options: smoother passes / compensator, Ruyten shapes and modified
    volumes switched off (smoothing.py:57-94, interpolation_grid.py:88-138)."""
    from fbpic.fields.smoothing import BinomialSmoother
    Nr, Nz, rmax, dz = 12, 16, 15.e-6, 0.4e-6
    dt = dz / c
    out = dict(Nr=Nr, Nz=Nz, rmax=rmax, dz=dz, dt=dt)
    cases = {'p2': dict(n_passes=2, compensator=False), 'p1c': dict(n_passes=1, compensator=True),
             'mixed': dict(n_passes={'z': 3, 'r': 1}, compensator={'z': True, 'r': False})}
    for tag, kw in cases.items():
        sim = Simulation(Nz, Nz * dz, Nr, rmax, 2, dt, zmin=0., smoother=BinomialSmoother(**kw), verbose_level=0)
        for m in range(2):
            out['%s_fz_m%d' % (tag, m)] = sim.fld.spect[m].filter_array_z
            out['%s_fr_m%d' % (tag, m)] = sim.fld.spect[m].filter_array_r
    for tag, kw in {'noruyten': dict(use_ruyten_shapes=False), 'novol': dict(use_modified_volume=False),
                    'neither': dict(use_ruyten_shapes=False, use_modified_volume=False)}.items():
        sim = Simulation(Nz, Nz * dz, Nr, rmax, 3, dt, zmin=0., verbose_level=0, **kw)
        for m in range(3):
            g = sim.fld.interp[m]
            out['%s_invvol_m%d' % (tag, m)] = g.invvol
            out['%s_lin_m%d' % (tag, m)] = g.ruyten_linear_coef
            out['%s_cub_m%d' % (tag, m)] = g.ruyten_cubic_coef
    save('tables_variants', **out)


STEP_OPTIONS = {
    'true_rho': dict(step=dict(use_true_rho=True), sim=dict(initialize_ions=True)),
    'true_rho_galilean': dict(step=dict(use_true_rho=True),
                              sim=dict(initialize_ions=True, v_comoving=-0.995 * c, use_galilean=True, n_order=16,
                                       n_guard=8)),
    'no_correction': dict(step=dict(correct_currents=False), sim=dict()),
    'correct_divE': dict(step=dict(correct_divE=True, correct_currents=False), sim=dict(initialize_ions=True)),
    'no_filter': dict(step=dict(), sim=dict(filter_currents=False)),
    'no_push_x': dict(step=dict(move_positions=False), sim=dict()),
    'no_push_p': dict(step=dict(move_momenta=False), sim=dict()),
    'nm1': dict(step=dict(), sim=dict(), Nm=1),
    'nm1_cubic_galilean': dict(step=dict(), sim=dict(particle_shape='cubic', v_comoving=-0.995 * c, use_galilean=True,
                                                     n_order=16, n_guard=8), Nm=1),
    'cubic_true_rho_nm3': dict(step=dict(use_true_rho=True), sim=dict(initialize_ions=True, particle_shape='cubic'),
                               Nm=3),
}


def gen_step_options(tag, nsteps=3):
    """Periodic plasma wave advanced with the non-default options of Simulation / step() (main.py:346-586)."""
    opt = STEP_OPTIONS[tag]
    np.random.seed(0)
    Nz, Nr, Nm, zmax, rmax = 24, 12, opt.get('Nm', 2), 12.e-6, 8.e-6
    dt = zmax / Nz / c
    n_e = 2.e24
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2,
                     p_nt=4 * max(Nm - 1, 1), n_e=n_e, verbose_level=0, boundaries={'z': 'periodic', 'r': 'reflective'},
                     **opt['sim'])
    k0 = 2 * np.pi / zmax * 2
    wp = np.sqrt(n_e * e**2 / (m_e * 8.8541878128e-12))
    impart_momenta(sim.ptcl[0], 0.05, k0, 3.e-6, wp)
    V = opt['sim'].get('v_comoving')
    if V is not None:
        g = 1. / np.sqrt(1 - (V / c)**2)
        for sp in sim.ptcl:
            sp.uz += -np.sqrt(g**2 - 1)
            sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, n_species=len(sim.ptcl))
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_in_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
        out['s%d_q' % i], out['s%d_m' % i] = sp.q, sp.m
    sim.step(nsteps, show_progress=False, **opt['step'])
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_out_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_opt_' + tag, **out)


def gen_mirror(tag, pml=False, gamma_boost=None, nsteps=30):
    """A laser pulse reflected by a Mirror (mirrors.py:10-94; main.py:751-753)."""
    from fbpic.lpa_utils.laser import add_laser_pulse, GaussianLaser
    from fbpic.lpa_utils.mirrors import Mirror
    Nz, Nr, Nm, zmax, rmax = 48, 12, 2, 24.e-6, 12.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 6},
                     gamma_boost=gamma_boost, verbose_level=0,
                     boundaries={'z': 'open', 'r': ('open' if pml else 'reflective')})
    add_laser_pulse(sim, GaussianLaser(a0=1., waist=4.e-6, tau=6.e-15, z0=8.e-6, lambda0=1.6e-6, theta_pol=0.4),
                    gamma_boost=gamma_boost)
    sim.mirrors = [Mirror(16.e-6, 17.5e-6, gamma_boost=gamma_boost, m=('all' if not pml else [1]))]
    sim.step(nsteps, show_progress=False)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, pml=pml,
               gamma_boost=(0. if gamma_boost is None else gamma_boost))
    out.update({'out_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    save('mirror_' + tag, **out)


def gen_tracking(nsteps=20):
    """Tracked electrons in a moving window with continuous injection: ids follow the particles and new ids are
    drawn for the injected plasma (tracking.py:15-130; particles.py:367-368, 376-392)."""
    np.random.seed(5)
    Nz, Nr, Nm, zmax, rmax = 32, 10, 2, 16.e-6, 8.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=4.e-6, p_zmax=60.e-6, p_rmin=0, p_rmax=6.e-6, p_nz=2, p_nr=2,
                     p_nt=4, n_e=1.e24, n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 4}, verbose_level=0,
                     boundaries={'z': 'open', 'r': 'reflective'})
    sp = sim.ptcl[0]
    sp.uz[:] = 0.4 * np.sin(2 * np.pi * sp.z / 8.e-6) * (sp.z < 10.e-6)
    sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.uz**2)
    sp.track(sim.comm)
    sim.set_moving_window(v=c)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, n_in=sp.Ntot, id_in=sp.tracker.id.copy())
    out.update({'in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    np.random.seed(6)
    sim.step(nsteps, show_progress=False)
    out.update({'out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['id_out'] = sp.tracker.id.copy()
    save('tracking_window', **out)


def gen_species_mix(nsteps=4):
    """Several species at once: thermal electrons, heavier positive ions with another sampling, a tracer species
    (pushed, never deposited) and a neutral one (q = 0: neither gathered nor pushed in momentum)
    (main.py:792-1001; particles.py:557-560, 690-693, 860-864)."""
    from scipy.constants import m_p
    np.random.seed(8)
    Nz, Nr, Nm, zmax, rmax = 24, 12, 2, 12.e-6, 8.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, verbose_level=0, boundaries={'z': 'periodic', 'r': 'reflective'})
    kw = dict(p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax)
    sim.add_new_species(q=-e, m=m_e, n=2.e24, p_nz=2, p_nr=2, p_nt=4, ux_th=0.02, uy_th=0.01, uz_th=0.05, uz_m=0.1, **kw)
    sim.add_new_species(q=2 * e, m=4 * m_p, n=1.e24, p_nz=1, p_nr=2, p_nt=6, **kw)
    sim.add_new_species(q=-e, m=m_e, n=1.e20, p_nz=1, p_nr=1, p_nt=4, is_tracer=True, ux_m=0.3, **kw)
    sim.add_new_species(q=0., m=m_e, n=1.e24, p_nz=1, p_nr=1, p_nt=4, uz_m=2., uy_th=0.1, **kw)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, n_species=len(sim.ptcl))
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_in_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
    sim.step(nsteps, show_progress=False)
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_out_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    save('step_species_mix', **out)


def gen_bunch_plane(tag, gamma_boost=None, nsteps=8):
    """A Gaussian bunch with z_injection_plane: particles before the plane move ballistically, the others feel the
    bunch's own field (bunch.py:117-119; push/numba_methods.py push_p_after_plane; ballistic_before_plane.py)."""
    from fbpic.lpa_utils.bunch import add_particle_bunch_gaussian
    from fbpic.lpa_utils.boosted_frame import BoostConverter
    np.random.seed(19)
    Nz, Nr, Nm, zmax, rmax = 40, 16, 2, 20.e-6, 16.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 10, 'r': 6},
                     gamma_boost=gamma_boost, verbose_level=0, boundaries={'z': 'open', 'r': 'reflective'})
    boost = BoostConverter(gamma_boost) if gamma_boost is not None else None
    sp = add_particle_bunch_gaussian(sim, -e, m_e, sig_r=2.e-6, sig_z=1.5e-6, n_emit=1.e-6, gamma0=8., sig_gamma=0.5,
                                     n_physical_particles=5.e9, n_macroparticles=1200, zf=9.e-6, boost=boost,
                                     z_injection_plane=10.e-6)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps,
               gamma_boost=(0. if gamma_boost is None else gamma_boost))
    out.update({'in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    sim.step(nsteps, show_progress=False)
    out.update({'out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    save('bunch_plane_' + tag, **out)


def _harvest(write_dir, prefix):
    """Every shim-h5py file under write_dir/hdf5 as '<prefix>/<file>:<path>' -> dataset, '...@attr' -> attribute."""
    import h5py
    out = {}
    d = os.path.join(write_dir, 'hdf5')
    for name in sorted(os.listdir(d)):
        for k, v in h5py.flatten(os.path.join(d, name)).items():
            if isinstance(v, bytes):
                v = np.bytes_(v)
            out['%s/%s:%s' % (prefix, name, k)] = v
    return out


def _diag_namespace():
    import types
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    if not hasattr(np, 'string_'):          # gone from NumPy 2; the reference's writers use it for text attributes
        np.string_ = np.bytes_
    from fbpic.openpmd_diag import (FieldDiagnostic, ParticleDiagnostic, ParticleChargeDensityDiagnostic,
                                    BackTransformedFieldDiagnostic, BackTransformedParticleDiagnostic,
                                    set_periodic_checkpoint)
    from fbpic.lpa_utils.laser import add_laser_pulse, GaussianLaser
    from fbpic.lpa_utils.boosted_frame import BoostConverter
    from fbpic.lpa_utils.bunch import add_elec_bunch_gaussian
    return types.SimpleNamespace(Simulation=Simulation, FieldDiagnostic=FieldDiagnostic,
                                 add_elec_bunch_gaussian=add_elec_bunch_gaussian,
                                 ParticleDiagnostic=ParticleDiagnostic,
                                 ParticleChargeDensityDiagnostic=ParticleChargeDensityDiagnostic,
                                 BackTransformedFieldDiagnostic=BackTransformedFieldDiagnostic,
                                 BackTransformedParticleDiagnostic=BackTransformedParticleDiagnostic,
                                 set_periodic_checkpoint=set_periodic_checkpoint, add_laser_pulse=add_laser_pulse,
                                 GaussianLaser=GaussianLaser, BoostConverter=BoostConverter)


def gen_diags():
    """The trees written by the reference's FieldDiagnostic, ParticleDiagnostic (with a selection and a tracked
    species), ParticleChargeDensityDiagnostic and checkpoints (fbpic/openpmd_diag/*.py) for tests/diag_cases.py,
    through the h5py stand-in of oracle/ref_shim."""
    import shutil
    import tempfile
    ns = _diag_namespace()
    import diag_cases
    sim, elec, ions = diag_cases.build_diag_sim(ns, verbose_level=0)
    tmp = tempfile.mkdtemp()
    dirs = diag_cases.attach_diags(ns, sim, elec, ions, tmp)
    np.random.seed(24)
    sim.step(diag_cases.DIAG_STEPS, show_progress=False)
    out = dict(nsteps=diag_cases.DIAG_STEPS)
    for d, tag in zip(dirs, diag_cases.DIAG_DIRS):
        out.update(_harvest(d, tag))
    shutil.rmtree(tmp)
    save('diags_tree', **out)


def gen_cpu_gpu_deposition(shape):
    """The CPU arm of the reference's tests/test_cpu_gpu_deposition.py: its FieldDiagnostic files of rho and J."""
    import shutil
    import tempfile
    ns = _diag_namespace()
    import diag_cases
    tmp = tempfile.mkdtemp()
    sim = diag_cases.build_cpu_gpu_deposition(ns, shape, tmp, use_cuda=False, verbose_level=0)
    sim.step(3, show_progress=False)
    tree = _harvest(tmp, 'cpu')
    shutil.rmtree(tmp)
    # the data and the grid attributes are what the test compares; keep the fixture small
    save('cpu_gpu_deposition_' + shape, **{k: v for k, v in tree.items() if '@' not in k or k.endswith(('@time', '@gridSpacing', '@gridGlobalOffset'))})


def gen_lab_diags():
    """Lab-frame snapshots of a boosted-frame run written by the reference's BackTransformedFieldDiagnostic
    (fbpic/openpmd_diag/boosted_field_diag.py) for tests/diag_cases.py."""
    import shutil
    import tempfile
    ns = _diag_namespace()
    import diag_cases
    sim, gamma_boost = diag_cases.build_lab_diag_sim(ns, verbose_level=0)
    tmp = tempfile.mkdtemp()
    d = diag_cases.attach_lab_diag(ns, sim, gamma_boost, tmp)
    np.random.seed(30)
    sim.step(diag_cases.LAB_DIAG_STEPS, show_progress=False)
    out = dict(nsteps=diag_cases.LAB_DIAG_STEPS)
    out.update(_harvest(d, 'lab'))
    out.update(_harvest(os.path.join(d, 'selected'), 'labsel'))
    shutil.rmtree(tmp)
    save('diags_lab_tree', **out)


def gen_ionization():
    """ADK tables of the reference's Ionizer for a few elements and its per-particle probability on random inputs
    (ionization/ionizer.py:137-183, inline_functions.py:9-45)."""
    import types
    from fbpic.particles.elementary_process.ionization.ionizer import Ionizer
    from fbpic.particles.elementary_process.ionization.inline_functions import get_E_amplitude, \
        get_ionization_probability
    out = dict(dt=1.3e-16)
    for element in ('H', 'He', 'N', 'Ar', 'Kr'):
        ion = types.SimpleNamespace(level_max=None)
        Ionizer.initialize_ADK_parameters(ion, element, 1.3e-16)
        out[element + '_prefactor'], out[element + '_power'], out[element + '_exp_prefactor'] = \
            ion.adk_prefactor, ion.adk_power, ion.adk_exp_prefactor
    rng = np.random.default_rng(51)
    n = 600
    u = rng.normal(size=(3, n)) * np.array([0.5, 0.5, 3.])[:, None]
    E = rng.normal(size=(3, n)) * 4.e12
    B = rng.normal(size=(3, n)) * 1.e4
    level = rng.integers(0, 7, n)
    ion = types.SimpleNamespace(level_max=None)
    Ionizer.initialize_ADK_parameters(ion, 'N', 1.3e-16)
    p = np.zeros(n)
    amp = np.zeros(n)
    for i in range(n):
        amp[i], g = get_E_amplitude(u[0, i], u[1, i], u[2, i], E[0, i], E[1, i], E[2, i], c * B[0, i], c * B[1, i],
                                    c * B[2, i])
        p[i] = get_ionization_probability(amp[i], g, ion.adk_prefactor[level[i]], ion.adk_power[level[i]],
                                          ion.adk_exp_prefactor[level[i]])
    out.update(u=u, E=E, B=B, level=level, amplitude=amp, probability=p)
    save('ionization', **out)


def gen_lasy_laser():
    """The reference's FromLasyFileLaser (laser_profiles.py:841-1065) on two synthetic lasy files (thetaMode with
    modes 0 and 1, and cartesian), written through the h5py stand-in; the fixture keeps the file contents so that the
    tests can rebuild the files."""
    import tempfile
    import h5py
    from fbpic.lpa_utils.laser import FromLasyFileLaser
    rng = np.random.default_rng(71)
    omega = 2 * np.pi * c / 0.8e-6
    pol = np.array([np.cos(0.3), np.sin(0.3) * np.exp(0.5j)])
    out = dict(omega=omega, pol=pol)
    tmp = tempfile.mkdtemp()

    def write(name, data, geometry, spacing, offset):
        path = os.path.join(tmp, name)
        with h5py.File(path, 'w') as f:
            f.attrs['software'], f.attrs['softwareVersion'] = np.bytes_('lasy'), np.bytes_('0.4.0')
            d = f.create_dataset('/data/0/meshes/laserEnvelope', data=data)
            d.attrs['angularFrequency'], d.attrs['polarization'] = omega, pol
            d.attrs['geometry'] = np.bytes_(geometry)
            d.attrs['gridSpacing'], d.attrs['gridGlobalOffset'], d.attrs['gridUnitSI'] = spacing, offset, 1.
        return path
    # thetaMode: [2 Nm - 1, nt, nr]
    nt, nr, dt_, dr_ = 60, 40, 1.e-15, 1.e-6
    tt, rr = np.meshgrid(dt_ * np.arange(nt), dr_ * np.arange(nr), indexing='ij')
    g = np.exp(-(tt - 30.e-15)**2 / (10.e-15)**2 - rr**2 / (12.e-6)**2) * np.exp(1.j * 2.e13 * tt)
    rt = 1.e12 * np.stack([g, 0.2 * g * rr / 12.e-6, 0.1j * g * rr / 12.e-6])
    p1 = write('theta.h5', rt, 'thetaMode', np.array([dt_, dr_]), np.array([-25.e-15, 0.]))
    # cartesian: [nt, ny, nx]
    ny, nx, dy_, dx_ = 24, 28, 2.e-6, 1.5e-6
    t3, y3, x3 = np.meshgrid(dt_ * np.arange(nt), -23.e-6 + dy_ * np.arange(ny), -20.e-6 + dx_ * np.arange(nx),
                             indexing='ij')
    xyz = 1.e12 * np.exp(-(t3 - 30.e-15)**2 / (10.e-15)**2 - (x3**2 + 1.5 * y3**2) / (12.e-6)**2 + 1.j * x3 / 5.e-6)
    p2 = write('cart.h5', xyz, 'cartesian', np.array([dt_, dy_, dx_]), np.array([-25.e-15, -23.e-6, -20.e-6]))
    n = 500
    x, y = rng.uniform(-22.e-6, 22.e-6, n), rng.uniform(-26.e-6, 26.e-6, n)
    t = rng.uniform(-5.e-15, 70.e-15, n)
    out.update(x=x, y=y, t=t, theta_data=rt, cart_data=xyz)
    for tag, path in (('theta', p1), ('cart', p2)):
        prof = FromLasyFileLaser(path, t_start=4.e-15)
        out[tag + '_Ex'], out[tag + '_Ey'] = prof.E_field(x, y, 0. * x, t)
    save('lasy_laser', **out)


GENERATORS = {
    'lasy_laser': gen_lasy_laser,
    'ionization': gen_ionization,
    'diags_tree': gen_diags,
    'cpu_gpu_deposition_linear': lambda: gen_cpu_gpu_deposition('linear'),
    'cpu_gpu_deposition_cubic': lambda: gen_cpu_gpu_deposition('cubic'),
    'diags_lab_tree': gen_lab_diags,
    'bunch_plane_lab': lambda: gen_bunch_plane('lab'),
    'bunch_plane_boost': lambda: gen_bunch_plane('boost', gamma_boost=3.),
    'species_mix': gen_species_mix,
    'tracking_window': gen_tracking,
    'mirror_lab': lambda: gen_mirror('lab'),
    'mirror_pml': lambda: gen_mirror('pml', pml=True),
    'mirror_boost': lambda: gen_mirror('boost', gamma_boost=2., nsteps=40),
    'tables_variants': gen_tables_variants,
    'script_lwfa': lambda: gen_script('lwfa'),
    'script_boosted': lambda: gen_script('boosted'),
    'bunch_uniform': lambda: gen_bunch('uniform'),
    'bunch_gaussian': lambda: gen_bunch('gaussian', gaussian=True),
    'bunch_gaussian_boost': lambda: gen_bunch('gaussian_boost', gaussian=True, gamma_boost=5.),
    'external_lab': lambda: gen_external('lab'),
    'external_boost': lambda: gen_external('boost', gamma_boost=4.),
    'laser_profiles': gen_laser_profiles,
    'laser_direct_gauss': lambda: gen_laser_direct('gauss'),
    'laser_direct_lg_pml': lambda: gen_laser_direct('lg_pml', lg=True, pml=True),
    'laser_direct_boost': lambda: gen_laser_direct('boost', gamma_boost=3.),
    'laser_antenna_lab': lambda: gen_laser_antenna('lab'),
    'laser_antenna_moving': lambda: gen_laser_antenna('moving', v_antenna=0.2 * c, cross=True, nsteps=16),
    'laser_antenna_boost': lambda: gen_laser_antenna('boost', gamma_boost=2.),
    'pml_periodic': lambda: gen_pml('periodic', False),
    'pml_open': lambda: gen_pml('open', True),
    'pml_galilean': lambda: gen_pml('galilean', True, v_comoving=0.999 * c, use_galilean=True),
    'pml_window': lambda: gen_pml('window', True, window=True, nsteps=10),
    'cross_std': lambda: gen_cross('std', None, False),
    'cross_galilean': lambda: gen_cross('galilean', -0.995 * c, True),
}

for _tag in STEP_OPTIONS:
    GENERATORS['step_opt_' + _tag] = (lambda t: (lambda: gen_step_options(t)))(_tag)

if __name__ == '__main__':
    only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
    for name, fn in GENERATORS.items():
        if only is None or only == name:
            fn()

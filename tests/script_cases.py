"""Scaled-down versions of the reference's two documented input scripts
(docs/source/example_input/lwfa_script.py and boosted_frame_script.py), written ONCE against a namespace `ns`
that provides `Simulation`, `add_laser_pulse`, `GaussianLaser`, `add_particle_bunch`, `BoostConverter`:
oracle/gen_golden_ext.py runs them with the unmodified reference's objects, the tests with fbpic_b200's -- the
same user code on both sides of the drop-in boundary.  (The diagnostics of the scripts are exercised separately:
tests/diag_cases.py, examples/.)"""
import numpy as np
from scipy.constants import c, e, m_e, m_p


def lwfa_dens_func(z, r):
    """density up-ramp as in lwfa_script.py:72-78"""
    ramp_start, ramp_length = 24.e-6, 12.e-6
    n = np.ones_like(z)
    n = np.where(z < ramp_start + ramp_length, (z - ramp_start) / ramp_length, n)
    n = np.where(z < ramp_start, 0., n)
    return n


def build_lwfa(ns, **sim_kw):
    """Laser-wakefield acceleration, lab frame: Gaussian pulse put on the grid, moving window at c, the plasma
    (with an up-ramp) enters through the right edge by continuous injection."""
    Nz, zmax, zmin, Nr, rmax, Nm = 64, 24.e-6, -8.e-6, 16, 16.e-6, 2
    dt = (zmax - zmin) / Nz / c
    sim = ns.Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, n_order=-1, n_guard=16, n_damp={'z': 16, 'r': 8},
                        boundaries={'z': 'open', 'r': 'reflective'}, **sim_kw)
    elec = sim.add_new_species(q=-e, m=m_e, n=4.e24, dens_func=lwfa_dens_func, p_zmin=24.e-6, p_zmax=400.e-6,
                               p_rmax=14.e-6, p_nz=2, p_nr=2, p_nt=4)
    ns.add_laser_pulse(sim, ns.GaussianLaser(2., 4.e-6, 8.e-15, 10.e-6, lambda0=1.6e-6))
    sim.set_moving_window(v=c)
    return sim, {'electrons': elec}, 56


RAMP_UP, PLATEAU, RAMP_DOWN = 40.e-6, 200.e-6, 40.e-6
REL_DELTA_N_OVER_W2 = 1. / (np.pi * 2.81e-15 * (20.e-6)**4 * 3.e24)


def boosted_dens_func(z, r):
    """ramps + plateau + parabolic channel, as in boosted_frame_script.py:97-121 (z in the lab frame)"""
    n = np.ones_like(z)
    n = np.where(z < RAMP_UP, z / RAMP_UP, n)
    n = np.where((z >= RAMP_UP + PLATEAU) & (z < RAMP_UP + PLATEAU + RAMP_DOWN),
                 -(z - (RAMP_UP + PLATEAU + RAMP_DOWN)) / RAMP_DOWN, n)
    n = np.where(z >= RAMP_UP + PLATEAU + RAMP_DOWN, 0, n)
    return n * (1. + REL_DELTA_N_OVER_W2 * r**2)


def build_boosted(ns, **sim_kw):
    """Boosted-frame LWFA (gamma_boost = 4): Galilean PSATD comoving with the plasma, electrons + ions flowing
    backwards with a lab-frame density profile, an externally injected electron bunch with its space charge,
    the laser emitted by an antenna at the plasma entrance, moving window."""
    gamma_boost = 4.
    boost = ns.BoostConverter(gamma_boost)
    Nz, zmax, zmin, Nr, rmax, Nm = 64, 0.e-6, -24.e-6, 16, 48.e-6, 2
    dt = min(rmax / (2 * boost.gamma0 * Nr) / c, (zmax - zmin) / Nz / c)
    n_e = 3.e24
    v_window = c * (1 - 0.5 * n_e / 1.75e27)
    v_comoving = -c * np.sqrt(1. - 1. / boost.gamma0**2)
    sim = ns.Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, v_comoving=v_comoving, gamma_boost=boost.gamma0,
                        n_order=-1, n_guard=16, n_damp={'z': 16, 'r': 8},
                        boundaries={'z': 'open', 'r': 'reflective'}, **sim_kw)
    kw = dict(n=n_e, dens_func=boosted_dens_func, boost_positions_in_dens_func=True, p_zmin=0.,
              p_zmax=RAMP_UP + PLATEAU + RAMP_DOWN, p_rmax=36.e-6, p_nz=2, p_nr=2, p_nt=4)
    elec = sim.add_new_species(q=-e, m=m_e, **kw)
    ions = sim.add_new_species(q=e, m=m_p, **kw)
    bunch = ns.add_particle_bunch(sim, -e, m_e, 400., 5.e23, -21.e-6, -18.e-6, 0, 9.e-6, boost=boost)
    ns.add_laser_pulse(sim, ns.GaussianLaser(2., 16.e-6, 16.e-15, -8.e-6, lambda0=3.2e-6, zf=0.),
                       gamma_boost=boost.gamma0, method='antenna', z0_antenna=0)
    v_window_boosted, = boost.velocity([v_window])
    sim.set_moving_window(v=v_window_boosted)
    return sim, {'electrons': elec, 'ions': ions, 'bunch': bunch}, 40


CASES = {'lwfa': build_lwfa, 'boosted': build_boosted}

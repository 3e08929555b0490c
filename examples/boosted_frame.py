"""Laser-wakefield acceleration in a Lorentz-boosted frame (gamma_boost = 10) with the Galilean PSATD solver: the
set-up of FBPIC's documented example (docs/source/example_input/boosted_frame_script.py) written against fbpic_b200 --
plasma electrons and ions with a lab-frame density profile flowing backwards, an externally injected electron
bunch with its space-charge field, the laser emitted by an antenna at the plasma entrance, moving window.

    python examples/boosted_frame.py [--steps N] [--out DIR]
"""
import argparse
import numpy as np
from scipy.constants import c, e, m_e, m_p

from fbpic_b200 import Simulation
from fbpic_b200.lpa_utils.laser import add_laser_pulse
from fbpic_b200.lpa_utils.laser.laser_profiles import GaussianLaser
from fbpic_b200.lpa_utils.bunch import add_particle_bunch
from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
from fbpic_b200.openpmd_diag import (FieldDiagnostic, ParticleDiagnostic, BackTransformedFieldDiagnostic,
                                     BackTransformedParticleDiagnostic)

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=None, help='default: the whole interaction')
ap.add_argument('--out', default='./diags')
ap.add_argument('--n-order', type=int, default=-1)
args = ap.parse_args()

gamma_boost = 10.
boost = BoostConverter(gamma_boost)
Nz, zmin, zmax, Nr, rmax, Nm = 600, -30.e-6, 0.e-6, 75, 150.e-6, 2
dt = min(rmax / (2 * boost.gamma0 * Nr) / c, (zmax - zmin) / Nz / c)
n_e, w_matched = 3.e24, 50.e-6
ramp_up, plateau, ramp_down = .5e-3, 3.5e-3, .5e-3
L_plasma = ramp_up + plateau + ramp_down
rel_delta_n_over_w2 = 1. / (np.pi * 2.81e-15 * w_matched**4 * n_e)


def dens_func(z, r):
    """ramps and plateau along z (lab frame), parabolic guiding channel along r"""
    n = np.ones_like(z)
    n = np.where(z < ramp_up, z / ramp_up, n)
    n = np.where((z >= ramp_up + plateau) & (z < L_plasma), -(z - L_plasma) / ramp_down, n)
    n = np.where(z >= L_plasma, 0, n)
    return n * (1. + rel_delta_n_over_w2 * r**2)


v_window = c * (1 - 0.5 * n_e / 1.75e27)
sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, v_comoving=-c * np.sqrt(1. - 1. / boost.gamma0**2),
                 gamma_boost=boost.gamma0, n_order=args.n_order, boundaries={'z': 'open', 'r': 'reflective'})
plasma = dict(n=n_e, dens_func=dens_func, boost_positions_in_dens_func=True, p_zmin=0., p_zmax=L_plasma,
              p_rmax=100.e-6, p_nz=2, p_nr=2, p_nt=6)
elec = sim.add_new_species(q=-e, m=m_e, **plasma)
ions = sim.add_new_species(q=e, m=m_p, **plasma)
bunch = add_particle_bunch(sim, -e, m_e, 400., 5.e23, -25.e-6, -22.e-6, 0, 10.e-6, boost=boost)
add_laser_pulse(sim, GaussianLaser(2., 50.e-6, 16.e-15, -10.e-6, lambda0=0.8e-6, zf=0.), gamma_boost=boost.gamma0,
                method='antenna', z0_antenna=0)
v_window_boosted, = boost.velocity([v_window])
sim.set_moving_window(v=v_window_boosted)
T_interact = boost.interaction_time(L_plasma, zmax - zmin, v_window)
N_lab_diag = 10 + 1
dt_lab_diag_period = (L_plasma + (zmax - zmin)) / v_window / (N_lab_diag - 1)
lab_dir = args.out.rstrip('/') + '_lab'
sim.diags = [  # in the boosted frame
             FieldDiagnostic(dt_period=T_interact / 15, fldobject=sim.fld, comm=sim.comm, write_dir=args.out),
             ParticleDiagnostic(dt_period=T_interact / 15, species={'electrons': elec, 'bunch': bunch}, comm=sim.comm,
                                write_dir=args.out),
             # in the lab frame (back-transformed), written every 50 cycles
             BackTransformedFieldDiagnostic(zmin, zmax, v_window, dt_lab_diag_period, N_lab_diag, boost.gamma0,
                                            fieldtypes=['rho', 'E', 'B'], period=50, fldobject=sim.fld, comm=sim.comm,
                                            write_dir=lab_dir),
             BackTransformedParticleDiagnostic(zmin, zmax, v_window, dt_lab_diag_period, N_lab_diag, boost.gamma0, 50,
                                               sim.fld, select={'uz': [0., None]}, species={'bunch': bunch},
                                               comm=sim.comm, write_dir=lab_dir)]
N_step = args.steps or int(T_interact / sim.dt)
sim.step(N_step)
if sim.comm.rank == 0:
    print('%d steps of %.3e s in the boosted frame; %d plasma electrons, %d bunch particles on rank 0; output in %s'
          % (N_step, sim.dt, elec.Ntot, bunch.Ntot, args.out))

"""Laser-wakefield acceleration with ionization injection on one B200: the set-up of FBPIC's documented example
(docs/source/example_input/ionization_script.py) written against fbpic_b200 -- a helium / nitrogen gas mix, pre-ionized
to He+ and N5+, whose remaining electrons are freed by the laser (ADK); the electrons from the inner shells of nitrogen
go to a species of their own.

    python examples/ionization_injection.py [--steps N] [--out DIR]
"""
import argparse
import numpy as np
from scipy.constants import c, e, m_e, m_p

from fbpic_b200 import Simulation, set_random_seed
from fbpic_b200.lpa_utils.laser import add_laser_pulse
from fbpic_b200.lpa_utils.laser.laser_profiles import GaussianLaser
from fbpic_b200.openpmd_diag import (FieldDiagnostic, ParticleDiagnostic, ParticleChargeDensityDiagnostic,
                                     set_periodic_checkpoint)

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=None, help='default: 50 microns of interaction + one window length')
ap.add_argument('--out', default='./diags')
ap.add_argument('--n-order', type=int, default=-1)
ap.add_argument('--checkpoint-period', type=int, default=0)
args = ap.parse_args()

Nz, zmax, zmin, Nr, rmax, Nm = 800, 10.e-6, -30.e-6, 50, 20.e-6, 2
dt = (zmax - zmin) / Nz / c
n_He, n_N, ramp_length = 2.e24, 1.e24, 20.e-6


def dens_func(z, r):
    n = np.ones_like(z)
    n = np.where(z < ramp_length, np.sin(np.pi / 2 * z / ramp_length)**2, n)
    return np.where(z < 0, 0., n)


set_random_seed(0)
sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, n_order=args.n_order,
                 boundaries={'z': 'open', 'r': 'reflective'})
gas = dict(dens_func=dens_func, p_nz=1, p_nr=2, p_nt=4, p_zmin=0.)
atoms_He = sim.add_new_species(q=e, m=4. * m_p, n=n_He, **gas)
atoms_N = sim.add_new_species(q=5 * e, m=14. * m_p, n=n_N, **gas)
elec = sim.add_new_species(q=-e, m=m_e, n=n_He + 5 * n_N, **gas)        # the electrons of the pre-ionized levels
atoms_He.make_ionizable('He', target_species=elec, level_start=1)
elec_from_N = sim.add_new_species(q=-e, m=m_e)
atoms_N.make_ionizable('N', target_species=elec_from_N, level_start=5)
add_laser_pulse(sim, GaussianLaser(4., 5.e-6, 16.e-15, -5.e-6, zf=20.e-6))
sim.set_moving_window(v=c)
sim.diags = [FieldDiagnostic(50, sim.fld, comm=sim.comm, write_dir=args.out),
             ParticleDiagnostic(50, {"electrons from N": elec_from_N, "electrons": elec}, comm=sim.comm,
                                write_dir=args.out),
             ParticleChargeDensityDiagnostic(50, sim, {"electrons": elec}, write_dir=args.out)]
if args.checkpoint_period:
    set_periodic_checkpoint(sim, args.checkpoint_period)
N_step = args.steps or int((50.e-6 + (zmax - zmin)) / c / sim.dt)
sim.step(N_step)
if sim.comm.rank == 0:
    print('%d steps; %d electrons freed from the K shell of nitrogen on rank 0; output in %s'
          % (N_step, elec_from_N.Ntot, args.out))

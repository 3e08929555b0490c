"""Laser-wakefield acceleration in the lab frame on one B200 (or on N slabs: torchrun --nproc-per-node N).
The set-up of FBPIC's documented example (docs/source/example_input/lwfa_script.py: a0 = 4 Gaussian pulse, plasma
with a 40 micron up-ramp entering a window that moves at c), written against fbpic_b200.

    python examples/lwfa.py [--steps N] [--out DIR]
"""
import argparse
import numpy as np
from scipy.constants import c, e, m_e

from fbpic_b200 import Simulation
from fbpic_b200.lpa_utils.laser import add_laser_pulse
from fbpic_b200.lpa_utils.laser.laser_profiles import GaussianLaser
from fbpic_b200.openpmd_diag import FieldDiagnostic, ParticleDiagnostic, set_periodic_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=None, help='default: one window length + 50 microns of plasma')
ap.add_argument('--out', default='./diags')
ap.add_argument('--n-order', type=int, default=-1, help='finite stencil order (needed with more than one GPU, e.g. 32)')
ap.add_argument('--checkpoint-period', type=int, default=0)
args = ap.parse_args()

Nz, zmin, zmax, Nr, rmax, Nm = 800, -10.e-6, 30.e-6, 50, 20.e-6, 2
dt = (zmax - zmin) / Nz / c
ramp_start, ramp_length = 30.e-6, 40.e-6


def dens_func(z, r):
    n = np.ones_like(z)
    n = np.where(z < ramp_start + ramp_length, (z - ramp_start) / ramp_length, n)
    return np.where(z < ramp_start, 0., n)


sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, n_order=args.n_order,
                 boundaries={'z': 'open', 'r': 'reflective'})
elec = sim.add_new_species(q=-e, m=m_e, n=4.e24, dens_func=dens_func, p_zmin=30.e-6, p_zmax=500.e-6, p_rmax=18.e-6,
                           p_nz=2, p_nr=2, p_nt=4)
add_laser_pulse(sim, GaussianLaser(a0=4., waist=5.e-6, tau=16.e-15, z0=15.e-6))
sim.set_moving_window(v=c)
sim.diags = [FieldDiagnostic(50, sim.fld, comm=sim.comm, write_dir=args.out),
             ParticleDiagnostic(50, {'electrons': elec}, select={'uz': [1., None]}, comm=sim.comm, write_dir=args.out)]
if args.checkpoint_period:
    set_periodic_checkpoint(sim, args.checkpoint_period)
N_step = args.steps or int((50.e-6 + (zmax - zmin)) / c / sim.dt)
sim.step(N_step)
if sim.comm.rank == 0:
    t = sim.last_step_timing
    print('%d steps, %d electrons on rank 0: %.2f ms/step on the device (+ %.0f ms H2D, %.0f ms D2H); output in %s'
          % (N_step, elec.Ntot, 1e3 * t['cycles_s'] / N_step, 1e3 * t['h2d_s'], 1e3 * t['d2h_s'], args.out))

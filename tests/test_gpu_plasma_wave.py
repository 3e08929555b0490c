"""The reference's acceptance test of the whole PIC loop, restated for the B200 implementation:
a linear periodic plasma wave in modes 0, 1, 2 (tests/test_periodic_plasma_wave.py of FBPIC:
Nz=200, Nr=64, Nm=3, 2x2x8 particles per cell, n_order=16, 0.75 plasma period).
Pass criteria are the reference's own (:359-362, :407-409):
  * E_z and E_r in the theta=0 half-plane agree with linear theory (atol 1.1e6 V/m, rtol 2e-2);
  * div E - rho/eps0 vanishes in spectral space to a relative RMS below 1e-11 in every mode."""
import numpy as np
import pytest
from scipy.constants import c, e, m_e, epsilon_0

pytestmark = pytest.mark.gpu

Nz, zmax, Nr, rmax, Nm, n_order = 200, 40.e-6, 64, 20.e-6, 3, 16
dt = zmax / Nz / c
n_e, w0, n_periods = 2.e24, 5.e-6, 3
eps = (1.e-3, 1.e-3, 1.e-3)
k0 = 2 * np.pi / zmax * n_periods
wp = np.sqrt(n_e * e**2 / (m_e * epsilon_0))
n_step = int(2 * np.pi / (wp * dt) * 0.75)


def envelope(x, y):
    """F(x, y) and its gradient: the wave potential is F(x,y) sin(k0 z); mode m contributes
    eps_m (2/w0)^m Re[(x + i y)^m] exp(-r^2/w0^2)."""
    g = np.exp(-(x**2 + y**2) / w0**2)
    p = eps[0] + eps[1] * 2 * x / w0 + eps[2] * 4 * (x**2 - y**2) / w0**2
    dpx = eps[1] * 2 / w0 + eps[2] * 8 * x / w0**2
    dpy = -eps[2] * 8 * y / w0**2
    return p * g, (dpx - 2 * x / w0**2 * p) * g, (dpy - 2 * y / w0**2 * p) * g


def set_wave_momenta(sp):
    """u = (c/wp) grad(-F sin k0 z) at t = 0 (velocities lead the field by a quarter period)."""
    F, Fx, Fy = envelope(sp.x, sp.y)
    sp.ux = -c / wp * Fx * np.sin(k0 * sp.z)
    sp.uy = -c / wp * Fy * np.sin(k0 * sp.z)
    sp.uz = -c / wp * k0 * F * np.cos(k0 * sp.z)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_periodic_plasma_wave(shape):
    from fbpic_b200 import Simulation
    from fbpic_b200.fields import Fields
    np.random.seed(0)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0., p_zmax=41.e-6, p_rmin=0., p_rmax=18.e-6,
                     p_nz=2, p_nr=2, p_nt=8, n_e=n_e, n_order=n_order, particle_shape=shape)
    # the immobile ions are the negative of the initial electron density
    sim.send_data_to_gpu()
    sim.deposit('rho_prev', exchange=True)
    sim.fld.spect2interp('rho_prev')
    sim.receive_data_from_gpu()
    rho_ions = [-sim.fld.interp[m].rho.copy() for m in range(Nm)]
    set_wave_momenta(sim.ptcl[0])
    sim.step(n_step, correct_currents=True)

    # ---- fields vs linear theory in the half-plane theta = 0 (y = 0, x = r)
    g0 = sim.fld.interp[0]
    z, r = np.meshgrid(g0.z, g0.r, indexing='ij')
    F, Fx, _ = envelope(r, np.zeros_like(r))
    amp = m_e * c**2 / e * np.sin(wp * sim.time)
    for name, theory in (('Ez', -amp * k0 * F * np.cos(k0 * z)), ('Er', -amp * Fx * np.sin(k0 * z))):
        sim_field = getattr(sim.fld.interp[0], name).real.copy()
        for m in range(1, Nm):
            sim_field += 2 * getattr(sim.fld.interp[m], name).real
        assert np.allclose(theory, sim_field, atol=1.1e6, rtol=2e-2), name

    # ---- charge conservation in spectral space
    chk = Fields(Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=n_order)
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez'):
            setattr(chk.interp[m], k, getattr(sim.fld.interp[m], k).copy())
        chk.interp[m].rho = sim.fld.interp[m].rho + rho_ions[m]
    chk.send_fields_to_gpu()
    chk.interp2spect('E')
    chk.interp2spect('rho_prev')
    chk.receive_fields_from_gpu()
    for m in range(Nm):
        s = chk.spect[m]
        divE = s.kr * (s.Ep - s.Em) + 1.j * s.kz * s.Ez
        rho_eps0 = s.rho_prev / epsilon_0
        rel = np.sqrt(np.sum(abs(divE - rho_eps0)**2) / np.sum(abs(rho_eps0)**2))
        assert rel < 1.e-11, 'mode %d: relative error on div E = %.3e' % (m, rel)

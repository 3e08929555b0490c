"""GPU tests of Compton scattering (fbpic_b200/compton.py; fbpic/particles/elementary_process/compton/): the reference's
own acceptance test (tests/test_compton.py) as written -- a relativistic electron bunch crosses a Gaussian laser pulse,
lab frame and boosted frame; only the position push and the scattering run, as in the reference's test."""
import numpy as np
import pytest
from scipy.constants import e, c, h, m_e, epsilon_0

pytestmark = pytest.mark.gpu

Q_bunch = 2080.5031144200598 * 30000 * e
gamma_bunch_mean, gamma_bunch_rms = 30.205798028084185, 0.58182474907848347
laser_energy, laser_radius, laser_duration = 1., 33.e-6, 2.e-12
laser_waist, laser_ctau = laser_radius * (2.)**.5, c * laser_duration
laser_wavelength = h * c / e            # 1 eV photons
laser_initial_z0 = c * 4 * laser_duration


def _run(gamma_boost, ratio_w_electron_photon, N_bunch):
    """tests/test_compton.py:46-118"""
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from fbpic_b200.lpa_utils.bunch import add_elec_bunch_gaussian
    Nz, zmax_lab, zmin_lab, Nr, rmax, Nm = 200, 20.e-6, -20.e-6, 50, 20.e-6, 2
    bunch_sigma_z = 1.e-6
    boost = BoostConverter(gamma_boost)
    N_step = 101
    laser_duration_boosted, = boost.copropag_length([laser_duration], beta_object=-1)
    bunch_sigma_z_boosted, = boost.copropag_length([bunch_sigma_z], beta_object=1)
    dt = (4 * laser_duration_boosted + bunch_sigma_z_boosted / c) / N_step
    zmax, zmin = boost.copropag_length([zmax_lab, zmin_lab], beta_object=1.)
    np.random.seed(0)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, dens_func=None, zmin=zmin,
                     boundaries={'z': 'periodic', 'r': 'reflective'})
    sim.ptcl = []
    add_elec_bunch_gaussian(sim, sig_r=1.e-6, sig_z=bunch_sigma_z, n_emit=0., gamma0=gamma_bunch_mean,
                            sig_gamma=gamma_bunch_rms, Q=Q_bunch, N=N_bunch, tf=0.0, zf=0.5 * (zmax + zmin),
                            boost=boost)
    elec = sim.ptcl[0]
    photons = sim.add_new_species(q=0, m=0)
    elec.activate_compton(target_species=photons, laser_energy=laser_energy, laser_wavelength=laser_wavelength,
                          laser_waist=laser_waist, laser_ctau=laser_ctau, laser_initial_z0=laser_initial_z0,
                          ratio_w_electron_photon=ratio_w_electron_photon, boost=boost)
    p_init = [(elec.w * getattr(elec, k)).sum() * m_e * c for k in ('ux', 'uy', 'uz')]
    for species in sim.ptcl:
        species.send_particles_to_gpu()
    for i_step in range(N_step):
        for species in sim.ptcl:
            species.push_x(0.5 * sim.dt)
        elec.handle_elementary_processes(sim.time + 0.5 * sim.dt)
        for species in sim.ptcl:
            species.push_x(0.5 * sim.dt)
        sim.time += sim.dt
        sim.iteration += 1
    for species in sim.ptcl:
        species.receive_particles_from_gpu()
    return boost, elec, photons, p_init


def _check_photon_fraction(simulated_frac):
    """within 10 % of the estimate from the Thomson limit of the Klein-Nishina formula (test_compton.py:171-189)"""
    beta_bunch_mean = np.sqrt(1 - 1. / gamma_bunch_mean**2)
    k = gamma_bunch_mean * (1 + beta_bunch_mean) * h / laser_wavelength / (m_e * c)
    assert k < 1.e-3
    r_e = 1. / (4 * np.pi * epsilon_0) * e**2 / (m_e * c**2)
    sigma = 8. / 3 * np.pi * r_e**2
    nphoton_per_surface = laser_energy / (np.pi / 2 * laser_waist**2) / (h * c / laser_wavelength)
    expected_frac = sigma * nphoton_per_surface
    assert abs(simulated_frac - expected_frac) < 0.1 * expected_frac, (simulated_frac, expected_frac)


@pytest.mark.parametrize('gamma_boost', [1., 10.])
def test_compton_as_written(gamma_boost):
    """test_compton_labframe / test_compton_boostedframe: the number of photons per electron after the crossing is the
    Klein-Nishina estimate within 10 %; the photons are emitted in a 1 / gamma cone around the bunch direction at up
    to 4 gamma^2 times the laser frequency."""
    boost, elec, photons, _ = _run(gamma_boost, 50, 300000)
    _check_photon_fraction(photons.w.sum() / elec.w.sum())
    photon_u = 1. / photons.inv_gamma
    lab_pz = boost.gamma0 * (photons.uz + boost.beta0 * photon_u)
    lab_p = boost.gamma0 * (photon_u + boost.beta0 * photons.uz)
    scaled_freq = lab_p * c / (h * 4 * gamma_bunch_mean**2 * c / laser_wavelength)
    gamma_theta = gamma_bunch_mean * np.arccos(np.clip(lab_pz / lab_p, -1., 1.))
    # (4 gamma^2 with the energy spread of the bunch: gamma up to about 30.2 + 4 x 0.58)
    assert photons.Ntot > 1000 and 0.9 < scaled_freq.max() < 1.3
    assert np.median(gamma_theta) < 1.5 and np.mean(lab_pz > 0) > 0.99
    # on axis the photon carries the full Doppler shift: omega / omega_max = 1 / (1 + (gamma theta)^2)
    near = gamma_theta < 1.
    assert np.abs(scaled_freq[near] * (1 + gamma_theta[near]**2) - 1.).mean() < 0.06


@pytest.mark.parametrize('gamma_boost', [1., 10.])
def test_compton_momentum_conservation(gamma_boost):
    """With one photon macroparticle per electron weight (ratio 1) the recoil is applied for every photon: the momentum
    lost by the electrons is the momentum gained by the photons relative to the incoming ones
    (test_compton.py:152-169, check_momentum_conservation), to 1e-9 of the bunch momentum."""
    boost, elec, photons, (px_init, py_init, pz_init) = _run(gamma_boost, 1, 60000)
    assert photons.Ntot > 0
    elec_p = [(elec.w * getattr(elec, k)).sum() * m_e * c for k in ('ux', 'uy', 'uz')]
    d_px, d_py = (photons.w * photons.ux).sum(), (photons.w * photons.uy).sum()
    incoming_pz = elec.compton_scatterer.photon_pz
    d_pz = (photons.w * (photons.uz - incoming_pz)).sum()
    atol = 1.e-9 * abs(pz_init)
    assert np.allclose(px_init, elec_p[0] + d_px, atol=atol)
    assert np.allclose(py_init, elec_p[1] + d_py, atol=atol)
    assert np.allclose(pz_init, elec_p[2] + d_pz, atol=atol)

"""
Compton scattering of a counter-propagating Gaussian laser pulse off an electron species, Monte-Carlo with the
Klein-Nishina cross-section (fbpic/particles/elementary_process/compton/{compton,numba_methods,cuda_methods,
inline_functions}.py).  The laser is not on the grid: it is a unidirectional photon flux whose density is evaluated
analytically at the electrons; scattered photons become macroparticles of `target_species` (q = 0, m = 0, their
`ux, uy, uz` hold the momentum in kg m/s and `inv_gamma` = 1 / |p|, as in the reference).  Works in a boosted frame.

Two kernels per cycle (`b2_compton_count`, `b2_compton_scatter`): the first one decides how many photons each electron
emits and returns the total, the host makes room at the end of the photon arrays, the second one writes the photons --
every electron reserves its slots with one atomic -- and applies the recoil.  The reference needs per-batch counts, a
host cumulative sum and per-batch random states for the same (compton.py:150-240); here all draws are counter-based
per (electron, draw), so the result does not depend on the execution order.
"""
import ctypes
import numpy as np
from scipy.constants import c, h, m_e, physical_constants

from . import _lib
from ._lib import DeviceArray, call, ptr_array

r_e = physical_constants['classical electron radius'][0]


class ComptonScatterer(object):
    """compton.py:30-147"""

    def __init__(self, source_species, target_species, laser_energy, laser_wavelength, laser_waist, laser_ctau,
                 laser_initial_z0, ratio_w_electron_photon, boost):
        assert target_species.q == 0
        assert ratio_w_electron_photon >= 1
        self.target_species = target_species
        self.ratio_w_electron_photon = ratio_w_electron_photon
        self.inv_ratio_w_elec_photon = 1. / ratio_w_electron_photon
        self.gamma_boost, self.beta_boost = (boost.gamma0, boost.beta0) if boost is not None else (1., 0.)
        # momentum of the incoming photons: along -z in the lab frame, boosted to the frame of the simulation
        photon_lab_pz = -h / laser_wavelength
        photon_lab_p = abs(photon_lab_pz)
        self.photon_px = self.photon_py = 0.
        self.photon_pz = self.gamma_boost * (photon_lab_pz - self.beta_boost * photon_lab_p)
        self.photon_p = abs(self.photon_pz)
        self.photon_beta_x, self.photon_beta_y = 0., 0.
        self.photon_beta_z = self.photon_pz / self.photon_p
        self.laser_initial_z0 = laser_initial_z0
        self.inv_laser_waist2, self.inv_laser_ctau2 = 1. / laser_waist**2, 1. / laser_ctau**2
        # peak photon density of the pulse (lab frame): energy / (effective volume x photon energy)
        effective_volume = (np.pi / 2.)**(3. / 2) * laser_waist**2 * laser_ctau
        self.photon_n_lab_peak = laser_energy / (effective_volume * photon_lab_p * c)
        self.seed = int(np.random.randint(0, 2**31 - 1))
        self.n_calls = 0
        self._nscatter = self._scalars = None

    def _params(self, elec, t):
        return np.array([c * t, self.photon_n_lab_peak, self.inv_laser_waist2, self.inv_laser_ctau2,
                         self.laser_initial_z0, self.gamma_boost, self.beta_boost, self.photon_p, self.photon_px,
                         self.photon_py, self.photon_pz, self.photon_beta_x, self.photon_beta_y, self.photon_beta_z,
                         elec.dt, self.ratio_w_electron_photon, self.inv_ratio_w_elec_photon, np.pi * r_e**2,
                         1. / (m_e * c), c], dtype=np.float64)

    def handle_scattering(self, elec, t):
        """compton.py:150-240"""
        n = elec.Ntot
        if n == 0:
            return
        elec._need_gpu()
        photons = self.target_species
        photons._need_gpu()
        ctx = _lib.context().handle
        if self._scalars is None:
            self._scalars = DeviceArray(2, np.int64)
        if self._nscatter is None or self._nscatter.size < n:
            self._nscatter = DeviceArray(elec._capacity_for(n), np.int32)
        params = self._params(elec, t)
        self.n_calls += 1
        seed = (self.seed * 1000003 + self.n_calls) & (2**64 - 1)
        created = ctypes.c_int64(0)
        call.b2_compton_count(ctx, n, elec.x.ptr, elec.y.ptr, elec.z.ptr, elec.ux.ptr, elec.uy.ptr, elec.uz.ptr,
                              elec.inv_gamma.ptr, params.ctypes.data, seed, self._nscatter.ptr, self._scalars.ptr,
                              ctypes.byref(created), None)
        N_created = int(created.value)
        if N_created == 0:
            return
        old = photons.Ntot
        photons.grow_device_arrays(old + N_created)
        from .particles import FLOAT_ATTRS
        assert FLOAT_ATTRS == ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
        slots = ptr_array([getattr(photons, a).ptr + 8 * old for a in FLOAT_ATTRS])
        call.b2_compton_scatter(ctx, n, self._nscatter.ptr, elec.x.ptr, elec.y.ptr, elec.z.ptr, elec.ux.ptr,
                                elec.uy.ptr, elec.uz.ptr, elec.inv_gamma.ptr, elec.w.ptr, params.ctypes.data, seed,
                                slots, self._scalars.ptr + 8, None)
        elec.sorted = False

"""
Laser initialisation and emission, same entry points as `fbpic.lpa_utils.laser`
(fbpic/lpa_utils/laser/{laser.py, laser_profiles.py, direct_injection.py, antenna_injection.py}):

* laser profiles (`GaussianLaser`, `LaguerreGaussLaser`, `DonutLikeLaguerreGaussLaser`, `FlattenedGaussianLaser`,
  `FewCycleLaser`, `ParaxialApproximationLaser` over longitudinal / transverse profile classes, sums with `+`):
  analytic E(x, y, z, t), NumPy; `FromLasyFileLaser`: envelope interpolated from a lasy file;
* `add_laser_pulse(sim, profile, method='direct')`: the profile is sampled on the global grid, Ez and B
  follow from div E = 0 and Faraday's law in spectral space.  The transforms of that one-off set-up run on
  the GPU (cuFFT + DMMA Hankel kernels of the hot path) -- the reference does them on the CPU even in GPU
  runs (direct_injection.py:61-64);
* `method='antenna'`: `LaserAntenna`, a sheet of virtual macroparticle pairs whose prescribed motion
  deposits the emitting current every step.  Their state lives in HBM; the deposition is the regular
  order-independent deposition kernel (linear shapes) fed by `b2_antenna_particles` -- the reference needs
  separate per-particle-atomics kernels for this (deposition/cuda_methods_unsorted.py).
"""
import numpy as np
from math import factorial
from scipy.constants import c, m_e, e, epsilon_0, physical_constants
from scipy.special import genlaguerre, binom

from ... import _lib
from ..._lib import DeviceArray, call, ptr_array
from ..boosted_frame import BoostConverter

r_e = physical_constants['classical electron radius'][0]


# =============================================================================
# profiles (laser_profiles.py, longitudinal_laser_profiles.py, transverse_laser_profiles.py)
# =============================================================================
class LaserProfile(object):
    """Base class: E_field(x, y, z, t) -> (Ex, Ey) in the lab frame; profiles add with `+`
    (laser_profiles.py:20-75)."""

    def __init__(self, propagation_direction, gpu_capable=False):
        assert propagation_direction in [-1, 1]
        self.propag_direction = float(propagation_direction)
        self.gpu_capable = gpu_capable

    def E_field(self, x, y, z, t):
        return np.zeros_like(x), np.zeros_like(x)

    def __add__(self, other):
        return SummedLaserProfile(self, other)


class SummedLaserProfile(LaserProfile):
    """laser_profiles.py:77-102"""

    def __init__(self, profile1, profile2):
        assert profile1.propag_direction == profile2.propag_direction
        LaserProfile.__init__(self, profile1.propag_direction)
        self.profile1, self.profile2 = profile1, profile2

    def E_field(self, x, y, z, t):
        a, b = self.profile1.E_field(x, y, z, t), self.profile2.E_field(x, y, z, t)
        return a[0] + b[0], a[1] + b[1]


# ---- longitudinal profiles (longitudinal_laser_profiles.py) ----
class LaserLongitudinalProfile(object):
    """Complex E(z, t) on the axis; `squared_profile_integral()` = integral of |E|^2 dz
    (longitudinal_laser_profiles.py:16-91)."""

    def __init__(self, propagation_direction, gpu_capable=False):
        assert propagation_direction in [-1, 1]
        self.propag_direction = float(propagation_direction)
        self.gpu_capable = gpu_capable
        self.lambda0 = 0.8e-6
        self.k0 = 2 * np.pi / self.lambda0

    def evaluate(self, z, t):
        return np.zeros_like(z, dtype='complex')

    def squared_profile_integral(self):
        return 0


class GaussianChirpedLongitudinalProfile(LaserLongitudinalProfile):
    """(Chirped) Gaussian pulse (longitudinal_laser_profiles.py:94-187)."""

    def __init__(self, tau, z0, lambda0=0.8e-6, cep_phase=0., phi2_chirp=0., propagation_direction=1):
        LaserLongitudinalProfile.__init__(self, propagation_direction, gpu_capable=True)
        self.lambda0, self.k0 = lambda0, 2 * np.pi / lambda0
        self.z0, self.cep_phase, self.phi2_chirp = z0, cep_phase, phi2_chirp
        self.inv_ctau2 = 1. / (c * tau)**2

    def evaluate(self, z, t):
        stretch = 1 - 2j * self.phi2_chirp * c**2 * self.inv_ctau2
        xi = self.propag_direction * (z - self.z0) - c * t
        return np.exp(-1j * self.cep_phase + 1j * self.k0 * xi - 1. / stretch * self.inv_ctau2 * xi**2) / stretch**0.5

    def squared_profile_integral(self):
        return (0.5 * np.pi / self.inv_ctau2)**.5


class CustomSpectrumLongitudinalProfile(LaserLongitudinalProfile):
    """Pulse shape computed from a measured spectrum: a tab-separated file with the wavelength (m), the spectral
    intensity per wavelength and, optionally, the spectral phase (rad); E(z, t) is the Fourier transform of
    sqrt(I lambda^2) exp(i phi) over omega, normalised to a peak of 1 (longitudinal_laser_profiles.py:190-354)."""

    def __init__(self, z0, spectrum_file, phi2_chirp=0., phi3_chirp=0., phi4_chirp=0., subtract_linear_phase=False,
                 propagation_direction=1):
        from scipy.interpolate import interp1d
        LaserLongitudinalProfile.__init__(self, propagation_direction, gpu_capable=False)
        trapezoid = getattr(np, 'trapezoid', None) or np.trapz
        data = np.loadtxt(spectrum_file, delimiter='\t')
        wavelength, intensity = data[:, 0], data[:, 1]
        phase = np.zeros_like(wavelength) if data.shape[1] < 3 else data[:, 2].copy()
        omega = 2 * np.pi * c / wavelength
        if subtract_linear_phase:
            phase -= omega * np.polyfit(omega, phase, 4)[-2]
        self.lambda0 = trapezoid(wavelength * intensity, wavelength) / trapezoid(intensity, wavelength)
        self.k0 = 2 * np.pi / self.lambda0
        d_omega = omega - self.k0 * c
        phase += phi2_chirp / 2. * d_omega**2 + phi3_chirp / 6. * d_omega**3 + phi4_chirp / 24. * d_omega**4
        intensity_fn = interp1d(omega, intensity * wavelength**2, fill_value=0, bounds_error=False)
        phase_fn = interp1d(omega, phase, fill_value=0, bounds_error=False)
        # a time window of lambda0^2 / (c d_lambda) sampled at d_lambda / c, d_lambda = lambda0 / 1000
        d_lambda = self.lambda0 / 1000
        dt = d_lambda / c
        window = self.lambda0 * self.lambda0 / c / d_lambda
        Nt = int(np.round(window / dt))
        time = -0.5 * window + dt * np.arange(Nt)
        w = 2 * np.pi * np.fft.fftfreq(Nt, dt)
        field = np.fft.fftshift(np.fft.fft(np.sqrt(intensity_fn(w)) * np.exp(1j * phase_fn(w))))
        field = field / abs(field).max()
        self.interp_Efield_function = interp1d(self.propag_direction * z0 - c * time, field, fill_value=0,
                                               bounds_error=False)
        self.squared_field_integral = trapezoid(abs(field)**2, c * time)

    def get_mean_wavelength(self):
        return self.lambda0

    def squared_profile_integral(self):
        return self.squared_field_integral

    def evaluate(self, z, t):
        return self.interp_Efield_function(self.propag_direction * z - c * t)


# ---- transverse profiles (transverse_laser_profiles.py) ----
class LaserTransverseProfile(object):
    """Complex E(x, y, z) of a monochromatic paraxial beam; `squared_profile_integral()` = integral of |E|^2 over
    the transverse plane (transverse_laser_profiles.py:14-88)."""

    def __init__(self, propagation_direction, gpu_capable=False):
        assert propagation_direction in [-1, 1]
        self.propag_direction = float(propagation_direction)
        self.gpu_capable = gpu_capable
        self.lambda0 = 0.8e-6
        self.k0 = 2 * np.pi / self.lambda0

    def _focus(self, waist, zf, lambda0):
        self.lambda0, self.k0 = lambda0, 2 * np.pi / lambda0
        self.w0, self.zf = waist, zf
        self.inv_zr = 1. / (0.5 * self.k0 * waist**2)

    def _diffract(self, z):
        return 1. + 1j * self.propag_direction * (z - self.zf) * self.inv_zr

    def evaluate(self, x, y, z):
        return np.zeros_like(x, dtype='complex')

    def squared_profile_integral(self):
        return 0


class GaussianTransverseProfile(LaserTransverseProfile):
    """Gaussian beam with Gouy phase / wavefront curvature away from focus (transverse_laser_profiles.py:91-166)."""

    def __init__(self, waist, zf=0., lambda0=0.8e-6, propagation_direction=1):
        LaserTransverseProfile.__init__(self, propagation_direction, gpu_capable=True)
        self._focus(waist, zf, lambda0)

    def evaluate(self, x, y, z):
        d = self._diffract(z)
        return np.exp(-(x**2 + y**2) / (self.w0**2 * d)) / d

    def squared_profile_integral(self):
        return 0.5 * np.pi * self.w0**2


class LaguerreGaussTransverseProfile(LaserTransverseProfile):
    """Laguerre-Gauss mode (p, m), azimuthal dependence cos(m (theta - theta0)), energy independent of p and m
    (transverse_laser_profiles.py:169-310)."""

    def __init__(self, p, m, waist, zf=0., lambda0=0.8e-6, theta0=0., propagation_direction=1):
        if m < 0 or type(m) is not int:
            raise ValueError("m should be an integer positive number.")
        LaserTransverseProfile.__init__(self, propagation_direction)
        self._focus(waist, zf, lambda0)
        self.p, self.m, self.theta0 = p, m, theta0
        self.scaled_amplitude = 1. if m == 0 else np.sqrt(factorial(p) / factorial(m + p)) * 2**.5
        self.laguerre_pm = genlaguerre(p, m)

    def evaluate(self, x, y, z):
        d = self._diffract(z)
        w = self.w0 * abs(d)
        psi = np.angle(d)
        r2 = x**2 + y**2
        s2 = 2 * r2 / w**2
        theta = np.angle(x + 1.j * y)
        prof = np.exp(-r2 / (self.w0**2 * d) - 1.j * (2 * self.p + self.m) * psi) / d \
            * np.sqrt(s2)**self.m * self.laguerre_pm(s2) * np.cos(self.m * (theta - self.theta0))
        return prof * self.scaled_amplitude

    def squared_profile_integral(self):
        return 0.5 * np.pi * self.w0**2


class DonutLikeLaguerreGaussTransverseProfile(LaserTransverseProfile):
    """Laguerre-Gauss mode (p, m) with the helical phase exp(-i m theta): a donut-like intensity profile
    (transverse_laser_profiles.py:312-432)."""

    def __init__(self, p, m, waist, zf=0., lambda0=0.8e-6, propagation_direction=1):
        LaserTransverseProfile.__init__(self, propagation_direction)
        self._focus(waist, zf, lambda0)
        self.p, self.m = p, m
        self.scaled_amplitude = np.sqrt(factorial(p) / factorial(abs(m) + p))
        self.laguerre_pm = genlaguerre(p, abs(m))

    def evaluate(self, x, y, z):
        d = self._diffract(z)
        w = self.w0 * abs(d)
        psi = np.angle(d)
        r2 = x**2 + y**2
        s2 = 2 * r2 / w**2
        theta = np.angle(x + 1.j * y)
        arg = -1.j * self.m * theta - r2 / (self.w0**2 * d) - 1.j * (2 * self.p + abs(self.m)) * psi
        return np.exp(arg) / d * np.sqrt(s2)**abs(self.m) * self.laguerre_pm(s2) * self.scaled_amplitude

    def squared_profile_integral(self):
        return 0.5 * np.pi * self.w0**2


class FlattenedGaussianTransverseProfile(LaserTransverseProfile):
    """Flat-top-like intensity far from focus (a Gaussian times a polynomial of order N there), built as a sum of
    N+1 Laguerre-Gauss modes of waist w0 sqrt(N+1) (transverse_laser_profiles.py:434-565)."""

    def __init__(self, w0, N=6, zf=0., lambda0=0.8e-6, propagation_direction=1):
        LaserTransverseProfile.__init__(self, propagation_direction)
        self.N = int(round(N))
        self.w_foc = w0 * (self.N + 1)**.5
        self._focus(self.w_foc, zf, lambda0)
        self.cn = np.empty(self.N + 1)
        for n in range(self.N + 1):
            mv = np.arange(n, self.N + 1)
            self.cn[n] = np.sum((1. / 2)**mv * binom(mv, n)) / (self.N + 1)

    def evaluate(self, x, y, z):
        d = self._diffract(z)
        w = self.w_foc * np.abs(d)
        psi = np.angle(d)
        r2 = x**2 + y**2
        s2 = 2 * r2 / w**2
        total = np.zeros_like(x, dtype=np.complex128)
        L_prev, L = 0., 1.                       # three-term recurrence of the Laguerre polynomials
        for n in range(self.N + 1):
            if n == 1:
                L_prev, L = L, 1. - s2
            elif n > 1:
                L_prev, L = L, (((2 * n - 1) - s2) * L - (n - 1) * L_prev) / n
            total += self.cn[n] * np.exp(-(2j * n) * psi) * L
        return total * np.exp(-r2 / (self.w_foc**2 * d)) / d

    def squared_profile_integral(self):
        return 0.5 * np.pi * self.w_foc**2 * sum(self.cn**2)


# ---- full profiles (laser_profiles.py) ----
class ParaxialApproximationLaser(LaserProfile):
    """E = Re[E0 longitudinal(z, t) transverse(x, y, z)], E0 such that the pulse carries the energy E_laser (J)
    (laser_profiles.py:105-176)."""

    def __init__(self, longitudinal_profile, transverse_profile, E_laser, theta_pol=0.):
        LaserProfile.__init__(self, 1)
        self.longitudinal_profile, self.transverse_profile = longitudinal_profile, transverse_profile
        self.propag_direction = longitudinal_profile.propag_direction
        assert self.propag_direction == transverse_profile.propag_direction
        assert longitudinal_profile.k0 == transverse_profile.k0
        self.k0 = longitudinal_profile.k0
        self.gpu_capable = longitudinal_profile.gpu_capable and transverse_profile.gpu_capable
        E0 = np.sqrt(2 * E_laser / (epsilon_0 * longitudinal_profile.squared_profile_integral()
                                    * transverse_profile.squared_profile_integral()))
        self.E0x, self.E0y = E0 * np.cos(theta_pol), E0 * np.sin(theta_pol)

    def E_field(self, x, y, z, t):
        prof = self.longitudinal_profile.evaluate(z, t) * self.transverse_profile.evaluate(x, y, z)
        return (self.E0x * prof).real, (self.E0y * prof).real


class _ParaxialLaser(LaserProfile):
    """A Gaussian (possibly chirped) pulse times a transverse profile, amplitude given by a0, polarised at
    theta_pol."""

    def __init__(self, a0, tau, z0, theta_pol, lambda0, cep_phase, phi2_chirp, propagation_direction, transverse):
        LaserProfile.__init__(self, propagation_direction)
        self.k0 = 2 * np.pi / lambda0
        E0 = a0 * m_e * c**2 * self.k0 / e
        self.E0x, self.E0y = E0 * np.cos(theta_pol), E0 * np.sin(theta_pol)
        self.longitudinal_profile = GaussianChirpedLongitudinalProfile(tau, z0, lambda0, cep_phase, phi2_chirp,
                                                                       propagation_direction)
        self.transverse_profile = transverse

    def E_field(self, x, y, z, t):
        prof = self.longitudinal_profile.evaluate(z, t) * self.transverse_profile.evaluate(x, y, z)
        return (self.E0x * prof).real, (self.E0y * prof).real


class GaussianLaser(_ParaxialLaser):
    """Linearly polarised Gaussian pulse (laser_profiles.py:179-293)."""

    def __init__(self, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 phi2_chirp=0., propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, tau, z0, theta_pol, lambda0, cep_phase, phi2_chirp, propagation_direction,
                                GaussianTransverseProfile(waist, z0 if zf is None else zf, lambda0,
                                                          propagation_direction))
        self.gpu_capable = True


class LaguerreGaussLaser(_ParaxialLaser):
    """Laguerre-Gauss pulse (laser_profiles.py:296-445)."""

    def __init__(self, p, m, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 theta0=0., propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, tau, z0, theta_pol, lambda0, cep_phase, 0., propagation_direction,
                                LaguerreGaussTransverseProfile(p, m, waist, z0 if zf is None else zf, lambda0,
                                                               theta0, propagation_direction))


class DonutLikeLaguerreGaussLaser(_ParaxialLaser):
    """Donut-like Laguerre-Gauss pulse (laser_profiles.py:448-584)."""

    def __init__(self, p, m, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, tau, z0, theta_pol, lambda0, cep_phase, 0., propagation_direction,
                                DonutLikeLaguerreGaussTransverseProfile(p, m, waist, z0 if zf is None else zf,
                                                                        lambda0, propagation_direction))


class FlattenedGaussianLaser(_ParaxialLaser):
    """Flattened Gaussian pulse (laser_profiles.py:587-710)."""

    def __init__(self, a0, w0, tau, z0, N=6, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, tau, z0, theta_pol, lambda0, cep_phase, 0., propagation_direction,
                                FlattenedGaussianTransverseProfile(w0, N, z0 if zf is None else zf, lambda0,
                                                                   propagation_direction))


class FromLasyFileLaser(LaserProfile):
    """Laser whose envelope E(x, y, t) in the emission plane comes from a `lasy` file (openPMD `laserEnvelope` mesh,
    `thetaMode` or `cartesian` geometry; laser_profiles.py:841-1065): multilinear interpolation of the envelope times
    exp(-i omega t), along the polarisation vector of the file.  The time axis of the file starts at `t_start`
    (its own offset is ignored, as in the reference).  Only for `method='antenna'`.  The file is read with the tree
    reader of the diagnostics (`.h5` through h5py, or an `.npz` archive of the same tree)."""

    def __init__(self, filename, t_start=0.):
        from ...openpmd_store import read_tree
        LaserProfile.__init__(self, propagation_direction=1, gpu_capable=False)
        self.t_start = t_start
        tree = read_tree(filename)

        def text(v):
            v = v[()] if isinstance(v, np.ndarray) and v.ndim == 0 else v
            return v.decode() if isinstance(v, bytes) else str(v)
        valid = False
        if '/@software' in tree and '/@softwareVersion' in tree:
            version = tuple(int(n) for n in text(tree['/@softwareVersion']).split('.'))
            valid = (text(tree['/@software']) == 'lasy') and version >= (0, 3, 0)
        if not valid:
            raise RuntimeError("The `lasy` version that was used to create the file %s is obsolete and not supported "
                               "by FBPIC.\nPlease upgrade your lasy version to at least 0.3.0 (e.g. with `pip install "
                               "--upgrade lasy`) and re-create the file %s." % (filename, filename))
        key = '/data/0/meshes/laserEnvelope'
        self.env_data = np.asarray(tree[key])
        self.omega = float(tree[key + '@angularFrequency'])
        self.pol = np.asarray(tree[key + '@polarization'])
        offset = np.asarray(tree[key + '@gridGlobalOffset'], dtype=np.float64)
        spacing = np.asarray(tree[key + '@gridSpacing'], dtype=np.float64) * float(tree[key + '@gridUnitSI'])
        self.t_min_lasy = offset[0]
        self.geometry = text(tree[key + '@geometry'])
        if self.geometry == 'thetaMode':
            self.inv_dt, self.inv_dr = 1. / spacing[0], 1. / spacing[1]
        elif self.geometry == 'cartesian':
            self.inv_dt, self.inv_dy, self.inv_dx = 1. / spacing[0], 1. / spacing[1], 1. / spacing[2]
            self.y_min, self.x_min = offset[1], offset[2]
        else:
            raise RuntimeError("Unknown geometry for lasy file %s: %s" % (filename, self.geometry))

    @staticmethod
    def _cell(pos, n):
        """lower index, weight of the upper node, inside-the-table flag of a linear interpolation on n nodes"""
        i = np.floor(pos).astype(np.int64)
        inside = (i >= 0) & (i + 1 <= n - 1)
        return np.where(inside, i, 0), pos - i, inside

    def envelope(self, x, y, t):
        x, y, t = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64),
                                      np.asarray(t, dtype=np.float64))
        d = self.env_data
        if self.geometry == 'thetaMode':
            _, nt, nr = d.shape
            r = (x**2 + y**2)**.5
            ir, Sr, ok_r = self._cell(r * self.inv_dr, nr)
            it, St, ok_t = self._cell(t * self.inv_dt, nt)

            def interp(a):
                return (1 - Sr) * (1 - St) * a[it, ir] + Sr * (1 - St) * a[it, ir + 1] \
                    + (1 - Sr) * St * a[it + 1, ir] + Sr * St * a[it + 1, ir + 1]
            env = interp(d[0]).astype(np.complex128)
            with np.errstate(invalid='ignore', divide='ignore'):
                phase = (x + 1.j * y) / r
            for m in range(1, d.shape[0] // 2 + 1):
                e_m = phase**m
                env = env + interp(d[2 * m - 1]) * e_m.real + interp(d[2 * m]) * e_m.imag
            return np.where(ok_r & ok_t, env, 0.)
        nt, ny, nx = d.shape
        ix, Sx, ok_x = self._cell((x - self.x_min) * self.inv_dx, nx)
        iy, Sy, ok_y = self._cell((y - self.y_min) * self.inv_dy, ny)
        it, St, ok_t = self._cell(t * self.inv_dt, nt)
        env = 0.
        for jt, wt in ((it, 1 - St), (it + 1, St)):
            for jy, wy in ((iy, 1 - Sy), (iy + 1, Sy)):
                for jx, wx in ((ix, 1 - Sx), (ix + 1, Sx)):
                    env = env + wx * wy * wt * d[jt, jy, jx]
        return np.where(ok_x & ok_y & ok_t, env, 0.)

    def E_field(self, x, y, z, t):
        E = self.envelope(x, y, t - self.t_start) * np.exp(-1.j * self.omega * (t - self.t_start + self.t_min_lasy))
        return (E * self.pol[0]).real, (E * self.pol[1]).real


class FewCycleLaser(LaserProfile):
    """Ultra-short, tightly focused pulse: an exact solution of the paraxial equation with a Poisson-like
    spectrum, valid down to a few cycles (laser_profiles.py:713-838)."""

    def __init__(self, a0, waist, tau_fwhm, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 propagation_direction=1):
        from scipy.optimize import fsolve
        LaserProfile.__init__(self, propagation_direction, gpu_capable=True)
        self.k0 = 2 * np.pi / lambda0
        E0 = a0 * m_e * c**2 * self.k0 / e
        self.zr = 0.5 * self.k0 * waist**2
        self.zf = z0 if zf is None else zf
        self.z0, self.w0, self.cep_phase = z0, waist, cep_phase
        self.E0x, self.E0y = E0 * np.cos(theta_pol), E0 * np.sin(theta_pol)
        w_tau = c * self.k0 * tau_fwhm          # the Poisson parameter s follows from the FWHM duration
        self.s = fsolve(lambda s: s * (2 * (4**(1 / (s + 1)) - 1))**.5 - w_tau, 1.)[0]

    def E_field(self, x, y, z, t):
        d = self.propag_direction
        inv_q = 1. / (d * (z - self.zf) + 1.j * self.zr)
        arg = 1. + 1.j * self.k0 / self.s * (d * (z - self.z0) - c * t + 0.5 * (x**2 + y**2) * inv_q)
        prof = np.exp(1.j * self.cep_phase) * 1.j * self.zr * inv_q * arg**(-self.s - 1)
        return (self.E0x * prof).real, (self.E0y * prof).real


# =============================================================================
# user entry points (laser.py:13-229)
# =============================================================================
def add_laser_pulse(sim, laser_profile, gamma_boost=None, method='direct', z0_antenna=None, v_antenna=0.):
    """Introduce a laser pulse: added to the grid at once (`direct`) or emitted progressively by an antenna
    at z0_antenna (`antenna`).  With `gamma_boost`, the profile is given in the lab frame and converted
    (laser.py:13-109)."""
    if (gamma_boost is not None) and (gamma_boost != 1.):
        if laser_profile.propag_direction == 1:
            boost = BoostConverter(gamma_boost)
        else:
            raise ValueError('For now, backward-propagating lasers cannot be used in the boosted-frame.')
    else:
        boost = None
    if method == 'direct':
        add_laser_direct(sim, laser_profile, boost)
    elif method == 'antenna':
        if z0_antenna is None:
            raise ValueError('You need to provide `z0_antenna`.')
        g0 = sim.fld.interp[0]
        sim.laser_antennas.append(LaserAntenna(laser_profile, z0_antenna, v_antenna, g0.dr, g0.Nr, sim.fld.Nm,
                                               boost, use_cuda=True))
    else:
        raise ValueError('Unknown laser method: %s' % method)


def add_laser(sim, a0, w0, ctau, z0, zf=None, lambda0=0.8e-6, cep_phase=0., phi2_chirp=0., theta_pol=0.,
              gamma_boost=None, method='direct', fw_propagating=True, update_spectral=True,
              z0_antenna=None, v_antenna=0.):
    """Linearly polarised Gaussian laser (laser.py:111-229)."""
    profile = GaussianLaser(a0, waist=w0, tau=ctau / c, z0=z0, zf=zf, theta_pol=theta_pol, lambda0=lambda0,
                            cep_phase=cep_phase, phi2_chirp=phi2_chirp,
                            propagation_direction=(1 if fw_propagating else -1))
    add_laser_pulse(sim, profile, gamma_boost=gamma_boost, method=method, z0_antenna=z0_antenna,
                    v_antenna=v_antenna)


# =============================================================================
# direct injection (direct_injection.py:12-217)
# =============================================================================
def get_laser_Er_Et(z, r, Nm, time, laser_profile, boost):
    """Sample the laser on (z, r, 2 Nm angles) and decompose into azimuthal modes; arrays [Nz, Nr, 2 Nm],
    the first Nm entries of the last axis are the modes m >= 0 (direct_injection.py:106-158)."""
    ntheta = 2 * Nm
    theta = (2 * np.pi / ntheta) * np.arange(ntheta)
    z3, r3, t3 = np.meshgrid(z, r, theta, indexing='ij')
    cs, sn = np.cos(t3), np.sin(t3)
    if boost is not None:
        zlab = boost.gamma0 * (z3 + boost.beta0 * c * time)
        tlab = boost.gamma0 * (time + (boost.beta0 * 1. / c) * z3)
    else:
        zlab, tlab = z3, time
    Ex, Ey = laser_profile.E_field(r3 * cs, r3 * sn, zlab, tlab)
    Er, Et = cs * Ex + sn * Ey, -sn * Ex + cs * Ey
    if boost is not None:
        scale = 1. / (boost.gamma0 * (1 + boost.beta0))
        Er, Et = Er * scale, Et * scale
    return np.fft.ifft(Er, axis=-1), np.fft.ifft(Et, axis=-1)


def laser_spectral_fields(spect, Nz, dz, propag_direction):
    """Given Ep, Em of the laser in spectral space (host arrays), fill Ez (div E = 0) and Bp, Bm, Bz
    (d_t B = -curl E for a pulse travelling along propag_direction), after a z-filter that removes the
    Nyquist noise (direct_injection.py:160-217).  Element-wise host arithmetic on the one-off set-up."""
    kz_true = 2 * np.pi * np.fft.fftfreq(Nz, dz)
    s2 = np.sin(0.5 * kz_true * dz)**2
    filt = ((1. - s2) * (1. + s2))[:, np.newaxis]
    for sg in spect:
        sg.Ep *= filt
        sg.Em *= filt
        kz, kr = sg.kz, sg.kr
        inv_kz = np.where(kz == 0, 0, 1. / np.where(kz == 0, 1., kz))
        sg.Ez[:, :] = 1.j * kr * (sg.Ep - sg.Em) * inv_kz
        w = c * np.sqrt(kz**2 + kr**2) * np.sign(kz) * propag_direction
        inv_w = np.where(w == 0, 0., 1. / np.where(w == 0, 1., w))
        sg.Bp[:, :] = -1.j * inv_w * (kz * sg.Ep - 0.5j * kr * sg.Ez)
        sg.Bm[:, :] = -1.j * inv_w * (-kz * sg.Em - 0.5j * kr * sg.Ez)
        sg.Bz[:, :] = inv_w * kr * (sg.Ep + sg.Em)


def add_laser_direct(sim, laser_profile, boost):
    """Add the laser to the interpolation grids of `sim` (host arrays, before the first step).

    Every rank samples the profile on the GLOBAL grid (damp cells included, guard cells not) -- the values
    the reference gathers from the ranks (direct_injection.py:45-75) -- runs the forward transform, the
    spectral construction of Ez and B and the inverse transforms of that global grid on its own GPU and adds
    its slab (direct_injection.py:85-101): no communication."""
    from ...fields import Fields
    comm, fld = sim.comm, sim.fld
    if fld.data_is_on_gpu:
        raise _lib.B200Error('add_laser_pulse(method="direct") acts on the host copy of the fields: call it '
                             'before step() or after receive_data_from_gpu()')
    Nz_g, iz_g = comm.get_Nz_and_iz(local=False, with_damp=True, with_guard=False)
    zmin_g, zmax_g = comm.get_zmin_zmax(local=False, with_damp=True, with_guard=False)
    gfld = Fields(Nz_g, zmax_g, fld.Nr, fld.rmax, fld.Nm, fld.dt, zmin=zmin_g, n_order=fld.n_order)
    Er_m, Et_m = get_laser_Er_Et(gfld.interp[0].z, gfld.interp[0].r, fld.Nm, sim.time, laser_profile, boost)
    for m in range(fld.Nm):
        gfld.interp[m].Er[:, :] = Er_m[:, :, m]
        gfld.interp[m].Et[:, :] = Et_m[:, :, m]
    gfld.send_fields_to_gpu()
    gfld.interp2spect('E')
    gfld.receive_fields_from_gpu()
    laser_spectral_fields(gfld.spect, Nz_g, gfld.interp[0].dz, laser_profile.propag_direction)
    gfld.send_fields_to_gpu()
    gfld.spect2interp('E')
    gfld.spect2interp('B')
    gfld.receive_fields_from_gpu()
    Nz_loc, iz_dom = comm.get_Nz_and_iz(local=True, with_damp=True, with_guard=False, rank=comm.rank)
    _, iz_arr = comm.get_Nz_and_iz(local=True, with_damp=True, with_guard=True, rank=comm.rank)
    i_loc, i_glob = iz_dom - iz_arr, iz_dom - iz_g
    for m in range(fld.Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            getattr(fld.interp[m], k)[i_loc:i_loc + Nz_loc, :] += \
                getattr(gfld.interp[m], k)[i_glob:i_glob + Nz_loc, :]


# =============================================================================
# antenna (antenna_injection.py:24-442)
# =============================================================================
class LaserAntenna(object):
    """A sheet of virtual macroparticle pairs at z = z0_antenna whose prescribed transverse motion
    v = mobility * E_laser(t) deposits the current j = 2 eps0 c E that emits the laser
    (antenna_injection.py:24-60).  Only the excursion of the positive particles is stored; the negative ones
    mirror it.  Linear shape factors whatever the shape of the plasma species."""

    def __init__(self, laser_profile, z0_antenna, v_antenna, dr_grid, Nr_grid, Nm, boost, npr=2,
                 epsilon=0.01, use_cuda=True):
        self.laser_profile, self.boost, self.use_cuda = laser_profile, boost, True
        if (v_antenna != 0) and (boost is not None) and boost.gamma0 != 1.:
            raise ValueError("For now, the boosted frame is incompatible with non-zero v_antenna.")
        nptheta = 2 * Nm
        # weights such that the maximal excursion is epsilon * dr (antenna_injection.py:110-125)
        alpha_weights = 2 * np.pi / (nptheta * npr * epsilon) * dr_grid / r_e * e
        self.mobility_coef = 2 * np.pi * dr_grid**2 / (nptheta * npr * alpha_weights) * epsilon_0 * c
        if boost is not None:
            self.mobility_coef = self.mobility_coef / boost.gamma0
        elif v_antenna is not None:
            self.mobility_coef *= (1. - laser_profile.propag_direction * v_antenna / c)
        Npr = Nr_grid * npr
        self.Ntot = Ntot = Npr * nptheta
        r_reg = dr_grid / npr * (np.arange(Npr) + 0.5)
        theta_reg = 2 * np.pi / nptheta * np.arange(nptheta)
        rp, thetap = np.meshgrid(r_reg, theta_reg, copy=True)
        self.baseline_r = rp.flatten()
        theta0 = thetap.flatten()
        self.baseline_x = self.baseline_r * np.cos(theta0)
        self.baseline_y = self.baseline_r * np.sin(theta0)
        self.baseline_z = z0_antenna * np.ones(Ntot)
        self.w = alpha_weights * self.baseline_r / dr_grid
        # host mirrors of the velocities: the profile is evaluated on the host every step (the profile is a
        # Python object); the excursions and the deposition state live on the device
        self.vx, self.vy, self.vz = np.zeros(Ntot), np.zeros(Ntot), np.zeros(Ntot)
        if boost is not None:
            self.baseline_z, = boost.static_length([self.baseline_z])
            self.vz, = boost.velocity([self.vz])
        elif v_antenna != 0:
            self.vz += v_antenna
        self.deposit_on_this_rank = False
        self._dev = None

    # -- device state
    def _device(self):
        if self._dev is None:
            n = self.Ntot
            d = dict(bx=DeviceArray.from_numpy(self.baseline_x), by=DeviceArray.from_numpy(self.baseline_y),
                     bz=DeviceArray.from_numpy(self.baseline_z), w=DeviceArray.from_numpy(self.w),
                     ex=DeviceArray.zeros(n, np.float64), ey=DeviceArray.zeros(n, np.float64),
                     vx=DeviceArray.from_numpy(self.vx), vy=DeviceArray.from_numpy(self.vy),
                     vz=DeviceArray.from_numpy(self.vz), one=DeviceArray.from_numpy(np.ones(n)))
            for k in ('x', 'y', 'ux', 'uy', 'uz'):
                d[k] = DeviceArray(n, np.float64)
            self._dev = d
        return self._dev

    @property
    def excursion_x(self):
        return self._device()['ex'].get()

    @property
    def excursion_y(self):
        return self._device()['ey'].get()

    def update_current_rank(self, comm):
        """antenna_injection.py:171-194"""
        zmin_local, zmax_local = comm.get_zmin_zmax(local=True, with_damp=True, with_guard=False, rank=comm.rank)
        z_antenna = self.baseline_z[0]
        self.deposit_on_this_rank = bool((z_antenna >= zmin_local) and (z_antenna < zmax_local))

    def push_x(self, dt, x_push=1., y_push=1., z_push=1.):
        """antenna_injection.py:196-218"""
        d, ctx = self._device(), _lib.context()
        call.b2_axpy(ctx.handle, self.Ntot, dt * x_push, d['vx'].ptr, d['ex'].ptr, None)
        call.b2_axpy(ctx.handle, self.Ntot, dt * y_push, d['vy'].ptr, d['ey'].ptr, None)
        call.b2_axpy(ctx.handle, self.Ntot, dt * z_push, d['vz'].ptr, d['bz'].ptr, None)
        self.baseline_z += (dt * z_push) * self.vz

    def update_v(self, t, dt):
        """Velocities at time t from the laser field at the antenna (antenna_injection.py:220-272): the
        profile is evaluated on the host, the two velocity arrays (16 B per virtual particle) go to HBM."""
        x = self.baseline_x + self.vx * 0.5 * dt
        y = self.baseline_y + self.vy * 0.5 * dt
        z = self.baseline_z + self.vz * 0.5 * dt
        if self.boost is not None:
            b = self.boost
            zlab = b.gamma0 * (z + (c * b.beta0) * t)
            tlab = b.gamma0 * (t + ((1. / c) * b.beta0) * z)
        else:
            zlab, tlab = z, t
        Ex, Ey = self.laser_profile.E_field(x, y, zlab, tlab)
        self.vx = self.mobility_coef * Ex
        self.vy = self.mobility_coef * Ey
        d = self._device()
        d['vx'].set(self.vx)
        d['vy'].set(self.vy)

    def deposit(self, fld, fieldtype):
        """Charge or current of the positive and negative virtual particles (antenna_injection.py:274-391)."""
        if not self.deposit_on_this_rank:
            return
        d, ctx = self._device(), _lib.context()
        grid = fld.interp
        g0, Nm = grid[0], len(grid)
        r0 = grid[0].d_ruyten_linear_coef
        rh = grid[1 if Nm > 1 else 0].d_ruyten_linear_coef
        for q in (-1, 1):
            call.b2_antenna_particles(ctx.handle, self.Ntot, d['bx'].ptr, d['by'].ptr, d['ex'].ptr, d['ey'].ptr,
                                      d['vx'].ptr, d['vy'].ptr, d['vz'].ptr, float(q), d['x'].ptr, d['y'].ptr,
                                      d['ux'].ptr, d['uy'].ptr, d['uz'].ptr, None)
            if fieldtype == 'rho':
                call.b2_deposit_rho(ctx.handle, self.Ntot, d['x'].ptr, d['y'].ptr, d['bz'].ptr, d['w'].ptr,
                                    float(q), g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm,
                                    ptr_array([g.rho for g in grid]), None, r0.ptr, rh.ptr, 0, None)
            elif fieldtype == 'J':
                call.b2_deposit_J(ctx.handle, self.Ntot, d['x'].ptr, d['y'].ptr, d['bz'].ptr, d['w'].ptr,
                                  float(q), d['ux'].ptr, d['uy'].ptr, d['uz'].ptr, d['one'].ptr,
                                  g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm,
                                  ptr_array([getattr(g, k) for g in grid for k in ('Jr', 'Jt', 'Jz')]),
                                  None, r0.ptr, rh.ptr, 0, None)
            else:
                raise ValueError('Unknown fieldtype: %s' % fieldtype)

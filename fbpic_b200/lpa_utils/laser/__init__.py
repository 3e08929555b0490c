"""
Laser initialisation and emission, same entry points as `fbpic.lpa_utils.laser`
(fbpic/lpa_utils/laser/{laser.py, laser_profiles.py, direct_injection.py, antenna_injection.py}):

* laser profiles (`GaussianLaser`, `LaguerreGaussLaser`, `DonutLikeLaguerreGaussLaser`, `FlattenedGaussianLaser`,
  `FewCycleLaser`, sums with `+`): analytic E(x, y, z, t), NumPy (`FromLasyFileLaser` needs h5py: not built);
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


def _gaussian_chirped_envelope(z, t, direction, z0, k0, tau, cep_phase, phi2_chirp):
    """Complex longitudinal profile of a (chirped) Gaussian pulse
    (longitudinal_laser_profiles.py:162-181)."""
    inv_ctau2 = 1. / (c * tau)**2
    stretch = 1 - 2j * phi2_chirp * c**2 * inv_ctau2
    xi = direction * (z - z0) - c * t
    return np.exp(-1j * cep_phase + 1j * k0 * xi - 1. / stretch * inv_ctau2 * xi**2) / stretch**0.5


class _ParaxialLaser(LaserProfile):
    """E = Re[ E0 * longitudinal(z, t) * transverse(x, y, z) ], polarised at theta_pol."""

    def __init__(self, a0, waist, tau, z0, zf, theta_pol, lambda0, cep_phase, phi2_chirp, propagation_direction):
        LaserProfile.__init__(self, propagation_direction)
        self.k0 = 2 * np.pi / lambda0
        E0 = a0 * m_e * c**2 * self.k0 / e
        self.E0x, self.E0y = E0 * np.cos(theta_pol), E0 * np.sin(theta_pol)
        self.w0, self.tau, self.z0 = waist, tau, z0
        self.zf = z0 if zf is None else zf
        self.inv_zr = 1. / (0.5 * self.k0 * waist**2)
        self.cep_phase, self.phi2_chirp = cep_phase, phi2_chirp

    def _diffract(self, z):
        return 1. + 1j * self.propag_direction * (z - self.zf) * self.inv_zr

    def transverse(self, x, y, z):
        raise NotImplementedError

    def E_field(self, x, y, z, t):
        prof = _gaussian_chirped_envelope(z, t, self.propag_direction, self.z0, self.k0, self.tau,
                                          self.cep_phase, self.phi2_chirp) * self.transverse(x, y, z)
        return (self.E0x * prof).real, (self.E0y * prof).real


class GaussianLaser(_ParaxialLaser):
    """Linearly polarised Gaussian pulse with Gouy phase / wavefront curvature away from focus
    (laser_profiles.py:179-293, transverse_laser_profiles.py:143-160)."""

    def __init__(self, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 phi2_chirp=0., propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, waist, tau, z0, zf, theta_pol, lambda0, cep_phase, phi2_chirp,
                                propagation_direction)
        self.gpu_capable = True

    def transverse(self, x, y, z):
        d = self._diffract(z)
        return np.exp(-(x**2 + y**2) / (self.w0**2 * d)) / d


class LaguerreGaussLaser(_ParaxialLaser):
    """Laguerre-Gauss mode (p, m), azimuthal dependence cos(m (theta - theta0)), pulse energy independent of
    p and m (laser_profiles.py:296-445, transverse_laser_profiles.py:169-310)."""

    def __init__(self, p, m, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 theta0=0., propagation_direction=1):
        if m < 0 or type(m) is not int:
            raise ValueError("m should be an integer positive number.")
        _ParaxialLaser.__init__(self, a0, waist, tau, z0, zf, theta_pol, lambda0, cep_phase, 0.,
                                propagation_direction)
        self.p, self.m, self.theta0 = p, m, theta0
        self.scaled_amplitude = 1. if m == 0 else np.sqrt(factorial(p) / factorial(m + p)) * 2**.5
        self.laguerre_pm = genlaguerre(p, m)

    def transverse(self, x, y, z):
        d = self._diffract(z)
        w = self.w0 * abs(d)
        psi = np.angle(d)
        r2 = x**2 + y**2
        s2 = 2 * r2 / w**2
        theta = np.angle(x + 1.j * y)
        prof = np.exp(-r2 / (self.w0**2 * d) - 1.j * (2 * self.p + self.m) * psi) / d \
            * np.sqrt(s2)**self.m * self.laguerre_pm(s2) * np.cos(self.m * (theta - self.theta0))
        return prof * self.scaled_amplitude


class DonutLikeLaguerreGaussLaser(_ParaxialLaser):
    """Laguerre-Gauss mode (p, m) with the helical phase exp(-i m theta): a donut-like intensity profile
    (laser_profiles.py:448-584, transverse_laser_profiles.py:312-432)."""

    def __init__(self, p, m, a0, waist, tau, z0, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 propagation_direction=1):
        _ParaxialLaser.__init__(self, a0, waist, tau, z0, zf, theta_pol, lambda0, cep_phase, 0.,
                                propagation_direction)
        self.p, self.m = p, m
        self.scaled_amplitude = np.sqrt(factorial(p) / factorial(abs(m) + p))
        self.laguerre_pm = genlaguerre(p, abs(m))

    def transverse(self, x, y, z):
        d = self._diffract(z)
        w = self.w0 * abs(d)
        psi = np.angle(d)
        r2 = x**2 + y**2
        s2 = 2 * r2 / w**2
        theta = np.angle(x + 1.j * y)
        arg = -1.j * self.m * theta - r2 / (self.w0**2 * d) - 1.j * (2 * self.p + abs(self.m)) * psi
        return np.exp(arg) / d * np.sqrt(s2)**abs(self.m) * self.laguerre_pm(s2) * self.scaled_amplitude


class FlattenedGaussianLaser(_ParaxialLaser):
    """Flat-top-like intensity far from focus (a Gaussian times a polynomial of order N there), built as a
    sum of N+1 Laguerre-Gauss modes of waist w0 sqrt(N+1) (laser_profiles.py:587-710,
    transverse_laser_profiles.py:434-565)."""

    def __init__(self, a0, w0, tau, z0, N=6, zf=None, theta_pol=0., lambda0=0.8e-6, cep_phase=0.,
                 propagation_direction=1):
        self.N = int(round(N))
        w_foc = w0 * (self.N + 1)**.5
        _ParaxialLaser.__init__(self, a0, w_foc, tau, z0, zf, theta_pol, lambda0, cep_phase, 0.,
                                propagation_direction)
        self.w_foc = w_foc
        self.cn = np.empty(self.N + 1)
        for n in range(self.N + 1):
            mv = np.arange(n, self.N + 1)
            self.cn[n] = np.sum((1. / 2)**mv * binom(mv, n)) / (self.N + 1)

    def transverse(self, x, y, z):
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

"""
`Fields` and its per-mode containers, backed by libfbpic_b200.so.

Mirrors the operator surface of fbpic/fields/fields.py:20-625 (`Fields`),
interpolation_grid.py:20-306 (`InterpolationGrid`), spectral_grid.py:33-478
(`SpectralGrid`), spectral_transform/spectral_transformer.py:21-223
(`SpectralTransformer`), hankel.py:25 (`DHT`), fourier.py:27 (`FFT`) and
psatd_coefs.py:15 (`PsatdCoeffs`): same class, method and attribute names.
Arrays are NumPy on the host until `send_fields_to_gpu()`, `DeviceArray`s (HBM)
afterwards, exactly like the reference swaps NumPy for CuPy arrays.

Radial PML split fields (`use_pml`) and the cross-deposition current correction
 are built, and so is correct_divE (NumPy-only in the reference, a device kernel here).
"""
import ctypes
import os
import numpy as np
from scipy.constants import mu_0, epsilon_0

from . import _lib
from . import host_tables as ht
from ._lib import DeviceArray, call, ptr_array, SpectralMode, DhtJob

INTERP_FIELDS = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')
SPECT_FIELDS = ('Ep', 'Em', 'Ez', 'Bp', 'Bm', 'Bz', 'Jp', 'Jm', 'Jz', 'rho_prev', 'rho_next')
RHO_TYPES = ('rho_prev', 'rho_next', 'rho_next_z', 'rho_next_xy')  # fields.py:361
PML_INTERP_FIELDS = ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')        # interpolation_grid.py:152-156
PML_SPECT_FIELDS = ('Ep_pml', 'Em_pml', 'Bp_pml', 'Bm_pml')         # spectral_grid.py:102-106
CROSS_SPECT_FIELDS = ('rho_next_z', 'rho_next_xy')                  # spectral_grid.py:97-99


class BinomialSmoother(object):
    """Binomial smoother description (fbpic/fields/smoothing.py:10-94)."""

    def __init__(self, n_passes=1, compensator=False):
        if type(n_passes) is int:
            self.n_passes = {'z': n_passes, 'r': n_passes}
        elif type(n_passes) is dict:
            self.n_passes = n_passes
        else:
            raise ValueError('Invalid argument `n_passes`')
        if type(compensator) is bool:
            self.compensator = {'z': compensator, 'r': compensator}
        elif type(compensator) is dict:
            self.compensator = compensator
        else:
            raise ValueError('Invalid argument `compensator`')

    def get_filter_array(self, kz, kr, dz, dr):
        return ht.binomial_filters(kz, kr, dz, dr, self.n_passes, self.compensator)


def _on_gpu(a):
    return isinstance(a, DeviceArray)


def _need_gpu(a):
    if not _on_gpu(a):
        raise _lib.B200Error('field data is on the host: call send_fields_to_gpu() '
                             '(fbpic_b200 has no CPU path)')


# =============================================================================
class InterpolationGrid(object):
    """Real-space (z, r) grid of one azimuthal mode (interpolation_grid.py:20-306)."""

    def __init__(self, Nz, Nr, m, zmin, zmax, rmax, use_pml=False, use_cuda=True,
                 use_ruyten_shapes=True, use_modified_volume=True):
        self.Nz, self.Nr, self.m, self.use_pml = Nz, Nr, m, bool(use_pml)
        self._fields = INTERP_FIELDS + (PML_INTERP_FIELDS if self.use_pml else ())
        dr = rmax / Nr
        dz = (zmax - zmin) / Nz
        self.dr, self.dz = dr, dz
        self.invdr, self.invdz = 1. / dr, 1. / dz
        self.rmin, self.rmax = 0., rmax
        self.zmin, self.zmax = zmin, zmax
        vol = ht.cell_volumes(m, Nr, rmax, dz, use_modified_volume)
        self.invvol = 1. / vol
        self.ruyten_linear_coef, self.ruyten_cubic_coef = ht.ruyten_coefs(vol, dr, dz, use_ruyten_shapes)
        for k in self._fields:
            setattr(self, k, np.zeros((Nz, Nr), dtype='complex'))
        self.use_cuda = True
        self.d_invvol = self.d_ruyten_linear_coef = self.d_ruyten_cubic_coef = None

    @property
    def z(self):
        return self.zmin + (0.5 + np.arange(self.Nz)) * self.dz

    @property
    def r(self):
        return self.rmin + (0.5 + np.arange(self.Nr)) * self.dr

    def send_fields_to_gpu(self, sources_recomputed=False):
        """`sources_recomputed` (set by Simulation.step): J and rho are erased and deposited afresh before
        anything reads them (main.py:446-454), so their host copies are not uploaded -- the device arrays start
        as zeros."""
        for k in self._fields:
            a = getattr(self, k)
            if sources_recomputed and k in ('Jr', 'Jt', 'Jz', 'rho') and not isinstance(a, DeviceArray):
                setattr(self, k, DeviceArray.zeros(a.shape, np.complex128))
            else:
                setattr(self, k, _lib.to_device(a))
        if self.d_invvol is None:
            self.d_invvol = DeviceArray.from_numpy(self.invvol)
            self.d_ruyten_linear_coef = DeviceArray.from_numpy(self.ruyten_linear_coef)
            self.d_ruyten_cubic_coef = DeviceArray.from_numpy(self.ruyten_cubic_coef)

    def receive_fields_from_gpu(self):
        for k in self._fields:
            setattr(self, k, _lib.to_host(getattr(self, k)))

    def _names(self, fieldtype):
        if fieldtype == 'rho':
            return ('rho',)
        if fieldtype in ('E', 'B', 'J'):
            return (fieldtype + 'r', fieldtype + 't', fieldtype + 'z')
        raise ValueError('Invalid string for fieldtype: %s' % fieldtype)

    def erase(self, fieldtype):
        for k in self._names(fieldtype):
            a = getattr(self, k)
            _need_gpu(a)
            a.fill(0)

    def divide_by_volume(self, fieldtype):
        if fieldtype not in ('rho', 'J'):
            raise ValueError('Invalid string for fieldtype: %s' % fieldtype)
        arrs = [getattr(self, k) for k in self._names(fieldtype)]
        _need_gpu(arrs[0])
        call.b2_scale_rows_by_r(_lib.context().handle, len(arrs), ptr_array(arrs), self.d_invvol.ptr,
                                self.Nz, self.Nr, None)


# =============================================================================
class FFT(object):
    """z-FFT of [Nz, Nr] arrays: cuFFT on the native layout (fourier.py:27-168)."""

    def __init__(self, Nr, Nz, use_cuda=True, nthreads=None):
        self.Nr, self.Nz = Nr, Nz
        self.use_cuda = True

    def transform(self, array_in, array_out):
        _need_gpu(array_in)
        call.b2_fft_z(_lib.context().handle, array_in.ptr, array_out.ptr, self.Nz, self.Nr, 0, None)

    def inverse_transform(self, array_in, array_out):
        _need_gpu(array_in)
        call.b2_fft_z(_lib.context().handle, array_in.ptr, array_out.ptr, self.Nz, self.Nr, 1, None)


class DHT(object):
    """Discrete Hankel transform of order p for mode m (hankel.py:25-243): host-built
    matrices, fp64 tensor-core GEMM on the device."""

    def __init__(self, p, m, Nr, Nz, rmax, use_cuda=True):
        self.p, self.m, self.Nr, self.Nz, self.rmax = p, m, Nr, Nz, rmax
        self.M, self.invM, self.nu = ht.hankel_matrices(p, m, Nr, rmax)
        self.r = (rmax * 1. / Nr) * (np.arange(Nr) + 0.5)
        self.use_cuda = True
        self.d_M = self.d_invM = None

    def get_r(self):
        return self.r

    def get_nu(self):
        return self.nu

    def _upload(self):
        if self.d_M is None:
            self.d_M = DeviceArray.from_numpy(self.M)
            self.d_invM = DeviceArray.from_numpy(self.invM)

    def transform(self, F, G):
        _need_gpu(F)
        self._upload()
        call.b2_dht(_lib.context().handle, F.ptr, G.ptr, self.d_M.ptr, None, self.Nz, self.Nr, None)

    def inverse_transform(self, G, F):
        _need_gpu(G)
        self._upload()
        call.b2_dht(_lib.context().handle, G.ptr, F.ptr, self.d_invM.ptr, None, self.Nz, self.Nr, None)


class SpectralTransformer(object):
    """FFT + DHT of one mode (spectral_transformer.py:21-223).  The (r,t)<->(p,m)
    combinations are fused into the Hankel GEMM (prologue / epilogue)."""

    def __init__(self, Nz, Nr, m, rmax, use_cuda=True):
        self.use_cuda = True
        self.Nz, self.Nr, self.m = Nz, Nr, m
        self.dht0 = DHT(m, m, Nr, Nz, rmax)
        self.dhtp = DHT(m + 1, m, Nr, Nz, rmax)
        self.dhtm = DHT(m - 1, m, Nr, Nz, rmax)
        self.fft = FFT(Nr, Nz)
        self.spect_buffer_r = self.spect_buffer_t = None

    def _buffers(self):
        if self.spect_buffer_r is None:
            self.spect_buffer_r = DeviceArray((self.Nz, self.Nr), np.complex128)
            self.spect_buffer_t = DeviceArray((self.Nz, self.Nr), np.complex128)
            self.spect_buffer_p, self.spect_buffer_m = self.spect_buffer_r, self.spect_buffer_t
            for d in (self.dht0, self.dhtp, self.dhtm):
                d._upload()

    def spect2interp_scal(self, spect_array, interp_array):
        self._buffers()
        self.dht0.inverse_transform(spect_array, self.spect_buffer_r)
        self.fft.inverse_transform(self.spect_buffer_r, interp_array)

    def spect2interp_vect(self, spect_array_p, spect_array_m, interp_array_r, interp_array_t):
        self._buffers()
        _need_gpu(spect_array_p)
        call.b2_dht_pm_to_rt(_lib.context().handle, spect_array_p.ptr, spect_array_m.ptr,
                             self.spect_buffer_r.ptr, self.spect_buffer_t.ptr,
                             self.dhtp.d_invM.ptr, self.dhtm.d_invM.ptr, None, self.Nz, self.Nr, None)
        self.fft.inverse_transform(self.spect_buffer_r, interp_array_r)
        self.fft.inverse_transform(self.spect_buffer_t, interp_array_t)

    def interp2spect_scal(self, interp_array, spect_array):
        self._buffers()
        self.fft.transform(interp_array, self.spect_buffer_r)
        self.dht0.transform(self.spect_buffer_r, spect_array)

    def interp2spect_vect(self, interp_array_r, interp_array_t, spect_array_p, spect_array_m):
        self._buffers()
        self.fft.transform(interp_array_r, self.spect_buffer_r)
        self.fft.transform(interp_array_t, self.spect_buffer_t)
        call.b2_dht_rt_to_pm(_lib.context().handle, self.spect_buffer_r.ptr, self.spect_buffer_t.ptr,
                             spect_array_p.ptr, spect_array_m.ptr, self.dhtp.d_M.ptr, self.dhtm.d_M.ptr,
                             None, self.Nz, self.Nr, None)


# =============================================================================
class PsatdCoeffs(object):
    """PSATD coefficient tables of one mode (psatd_coefs.py:15-177); built on the
    host from the 1-D kz, kr axes and uploaded once."""

    NAMES_STD = ('C', 'S_w', 'j_coef', 'rho_prev_coef', 'rho_next_coef')
    NAMES_COM = ('T_eb', 'T_cc', 'T_rho', 'j_corr_coef')

    def __init__(self, kz, kr, m, dt, Nz, Nr, V=None, use_galilean=False, use_cuda=True):
        kz1 = kz[:, 0] if np.ndim(kz) == 2 else kz
        kr1 = kr[0, :] if np.ndim(kr) == 2 else kr
        self.m, self.dt, self.V = m, dt, V
        t = ht.psatd_coefficients(kz1, kr1, dt, V, use_galilean)
        for k, v in t.items():
            setattr(self, k, v)
        self._device = {}

    def device(self, name):
        if name not in self._device:
            a = getattr(self, name)
            if self.V is not None and name not in ('C', 'S_w'):
                a = a.astype(np.complex128)       # comoving kernels read complex coefficients
            self._device[name] = DeviceArray.from_numpy(a)
        return self._device[name]


class SpectralGrid(object):
    """Spectral (kz, kr) grid of one mode (spectral_grid.py:33-478)."""

    def __init__(self, kz_modified, kr, m, kz_true, dz, dr, current_correction, smoother,
                 use_pml=False, use_cuda=True):
        if current_correction not in ('curl-free', 'cross-deposition'):
            raise ValueError('Unkown current correction:%s' % current_correction)
        Nz, Nr = len(kz_modified), len(kr)
        self.Nz, self.Nr, self.m, self.use_pml = Nz, Nr, m, bool(use_pml)
        self._fields = SPECT_FIELDS + (PML_SPECT_FIELDS if self.use_pml else ()) + \
            (CROSS_SPECT_FIELDS if current_correction == 'cross-deposition' else ())
        for k in self._fields:
            setattr(self, k, np.zeros((Nz, Nr), dtype='complex'))
        self.kz, self.kr = np.meshgrid(kz_modified, kr, indexing='ij')
        self.kz_1d, self.kr_1d = np.ascontiguousarray(kz_modified), np.ascontiguousarray(kr)
        self.filter_array_z, self.filter_array_r = smoother.get_filter_array(kz_true, kr, dz, dr)
        self.inv_k2 = ht.inverse_k2(self.kz_1d, self.kr_1d)
        self.field_shift = np.exp(1.j * kz_true * dz)
        self.use_cuda = True
        self.d_kz = self.d_kr = self.d_inv_k2 = self.d_filter_array_z = self.d_filter_array_r = None

    def send_fields_to_gpu(self, recomputed=False):
        """`recomputed` (set by Simulation.step): every spectral array is rebuilt from the interpolation grid or
        from a fresh deposition before it is read (main.py:408-415, 446-454), so nothing is uploaded."""
        for k in self._fields:
            a = getattr(self, k)
            if recomputed and not isinstance(a, DeviceArray):
                setattr(self, k, DeviceArray.zeros(a.shape, np.complex128))
            else:
                setattr(self, k, _lib.to_device(a))
        if self.d_kz is None:
            self.d_kz = DeviceArray.from_numpy(self.kz_1d)
            self.d_kr = DeviceArray.from_numpy(self.kr_1d)
            self.d_inv_k2 = DeviceArray.from_numpy(self.inv_k2)
            self.d_filter_array_z = DeviceArray.from_numpy(self.filter_array_z)
            self.d_filter_array_r = DeviceArray.from_numpy(self.filter_array_r)

    def receive_fields_from_gpu(self):
        for k in self._fields:
            setattr(self, k, _lib.to_host(getattr(self, k)))

    def _mode_struct(self, ps):
        s = SpectralMode()
        for k in SPECT_FIELDS:
            a = getattr(self, k)
            _need_gpu(a)
            setattr(s, k, a.ptr)
        s.kz, s.kr, s.inv_k2 = self.d_kz.ptr, self.d_kr.ptr, self.d_inv_k2.ptr
        for k in PsatdCoeffs.NAMES_STD:
            setattr(s, k, ps.device(k).ptr)
        if ps.V is not None:
            for k in PsatdCoeffs.NAMES_COM:
                setattr(s, k, ps.device(k).ptr)
        s.mu_0, s.epsilon_0 = mu_0, epsilon_0
        return s

    def correct_currents(self, dt, ps, current_correction):
        """spectral_grid.py:198-296"""
        s = self._mode_struct(ps)
        if current_correction == 'curl-free':
            call.b2_correct_currents(_lib.context().handle, ctypes.byref(s), int(ps.V is not None),
                                     1. / dt, self.Nz, self.Nr, None)
        elif current_correction == 'cross-deposition':
            _need_gpu(self.rho_next_z)
            call.b2_correct_currents_cross(_lib.context().handle, ctypes.byref(s), self.rho_next_z.ptr,
                                           self.rho_next_xy.ptr, int(ps.V is not None), 1. / dt,
                                           self.Nz, self.Nr, None)
        else:
            raise ValueError('Unkown current correction:%s' % current_correction)

    def correct_divE(self, ps):
        """E <- E - grad-like correction such that div E = rho_prev/eps0 (spectral_grid.py:299-314)"""
        s = self._mode_struct(ps)
        call.b2_correct_divE(_lib.context().handle, ctypes.byref(s), self.Nz, self.Nr, None)

    def push_eb_pml_with(self, ps):
        """Split PML components (spectral_grid.py:343-348, 358-363); reads the Ez, Bz of BEFORE the
        regular push, so it is issued first."""
        _need_gpu(self.Ep_pml)
        call.b2_push_eb_pml(_lib.context().handle, self.Ep_pml.ptr, self.Em_pml.ptr, self.Bp_pml.ptr,
                            self.Bm_pml.ptr, self.Ez.ptr, self.Bz.ptr, ps.device('C').ptr, ps.device('S_w').ptr,
                            ps.device('T_eb').ptr if ps.V is not None else None, self.d_kr.ptr,
                            self.Nz, self.Nr, None)

    def push_eb_with(self, ps, use_true_rho=False):
        """E, B push; the kernel also performs push_rho (rho_prev <- rho_next, rho_next <- 0),
        so `push_rho()` below is a no-op marker when called right after it."""
        assert self.m == ps.m
        if self.use_pml:
            self.push_eb_pml_with(ps)
        s = self._mode_struct(ps)
        call.b2_push_eb(_lib.context().handle, ctypes.byref(s), int(ps.V is not None), ps.dt,
                        0. if ps.V is None else ps.V, int(bool(use_true_rho)), self.Nz, self.Nr, None)
        self._rho_pushed = True

    def correct_and_push(self, ps, use_true_rho=False):
        """Fused correct_currents + push_eb + push_rho: one pass over the mode's arrays."""
        if self.use_pml:
            self.push_eb_pml_with(ps)
        s = self._mode_struct(ps)
        call.b2_correct_push(_lib.context().handle, ctypes.byref(s), int(ps.V is not None), ps.dt,
                             0. if ps.V is None else ps.V, int(bool(use_true_rho)), self.Nz, self.Nr, None)
        self._rho_pushed = True

    def push_rho(self):
        if getattr(self, '_rho_pushed', False):
            self._rho_pushed = False
            return
        _need_gpu(self.rho_prev)
        self.rho_prev.copy_from(self.rho_next)
        self.rho_next.fill(0)

    def filter(self, fieldtype):
        if fieldtype in ('J', 'E', 'B'):
            arrs = [getattr(self, fieldtype + c) for c in ('p', 'm', 'z')]
        elif fieldtype in RHO_TYPES:
            arrs = [getattr(self, fieldtype)]
        else:
            raise ValueError('Invalid string for fieldtype: %s' % fieldtype)
        _need_gpu(arrs[0])
        call.b2_filter(_lib.context().handle, len(arrs), ptr_array(arrs), self.d_filter_array_z.ptr,
                       self.d_filter_array_r.ptr, self.Nz, self.Nr, None)


# =============================================================================
class Fields(object):
    """All field data of the simulation (fields.py:20-625)."""

    def __init__(self, Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1, v_comoving=None,
                 use_pml=False, use_galilean=True, current_correction='curl-free', use_cuda=True,
                 smoother=None, create_threading_buffers=False, use_ruyten_shapes=True,
                 use_modified_volume=True):
        if current_correction not in ('curl-free', 'cross-deposition'):
            raise ValueError('Unkown current correction:%s' % current_correction)
        self.Nz, self.Nr, self.rmax, self.Nm, self.dt = Nz, Nr, rmax, Nm, dt
        self.n_order, self.v_comoving, self.use_galilean = n_order, v_comoving, use_galilean
        self.smoother = smoother if smoother is not None else BinomialSmoother(1, False)
        self.use_cuda, self.use_pml = True, bool(use_pml)
        self.data_is_on_gpu = False
        self.current_correction = current_correction
        self.trans = [SpectralTransformer(Nz, Nr, m, rmax) for m in range(Nm)]
        self.interp = [InterpolationGrid(Nz, Nr, m, zmin, zmax, rmax, use_pml=use_pml,
                                         use_ruyten_shapes=use_ruyten_shapes,
                                         use_modified_volume=use_modified_volume) for m in range(Nm)]
        dz = (zmax - zmin) / Nz
        kz_true = 2 * np.pi * np.fft.fftfreq(Nz, dz)
        kz_modified = ht.modified_kz(kz_true, n_order, dz)
        self.spect, self.psatd = [], []
        for m in range(Nm):
            kr = 2 * np.pi * self.trans[m].dht0.get_nu()
            self.spect.append(SpectralGrid(kz_modified, kr, m, kz_true, self.interp[m].dz,
                                           self.interp[m].dr, current_correction, self.smoother,
                                           use_pml=use_pml))
            self.psatd.append(PsatdCoeffs(kz_modified, kr, m, dt, Nz, Nr, V=v_comoving,
                                          use_galilean=use_galilean))
        self.exchanged_source = {'J': False, 'rho_prev': False, 'rho_new': False,
                                 'rho_next_xy': False, 'rho_next_z': False}

    def send_fields_to_gpu(self, step_entry=False):
        """fields.py:221-232.  `step_entry`: called by Simulation.step(N >= 1), which recomputes the sources and
        all spectral arrays before their first use -- only E and B of the interpolation grid cross the bus."""
        for m in range(self.Nm):
            self.interp[m].send_fields_to_gpu(sources_recomputed=step_entry)
            self.spect[m].send_fields_to_gpu(recomputed=step_entry)
        self.data_is_on_gpu = True

    def receive_fields_from_gpu(self):
        self.join_side()
        for m in range(self.Nm):
            self.interp[m].receive_fields_from_gpu()
            self.spect[m].receive_fields_from_gpu()
        self.data_is_on_gpu = False

    # ---- spectral solver ----
    def push(self, use_true_rho=False, check_exchanges=False):
        if check_exchanges:
            assert self.exchanged_source['J'] is True
            if use_true_rho:
                assert self.exchanged_source['rho_prev'] is True
                assert self.exchanged_source['rho_next'] is True
        self.join_side()       # the spectral E, B may still be on their way (partial_interp2spect(side=True))
        for m in range(self.Nm):
            self.spect[m].push_eb_with(self.psatd[m], use_true_rho)
            self.spect[m].push_rho()

    def correct_currents(self, check_exchanges=False):
        if check_exchanges:
            assert self.exchanged_source['rho_prev'] is False
            assert self.exchanged_source['rho_next'] is False
            assert self.exchanged_source['J'] is False
            if self.current_correction == 'cross-deposition':
                assert self.exchanged_source['rho_next_xy'] is False
                assert self.exchanged_source['rho_next_z'] is False
        for m in range(self.Nm):
            self.spect[m].correct_currents(self.dt, self.psatd[m], self.current_correction)

    def correct_divE(self):
        """fields.py:298-311 (the reference runs this one with NumPy; here it is a device kernel)"""
        self.join_side()
        for m in range(self.Nm):
            self.spect[m].correct_divE(self.psatd[m])

    def correct_currents_and_push(self, use_true_rho=False):
        """Fused Fields.correct_currents() + Fields.push() (single-domain fast path)."""
        self.join_side()
        for m in range(self.Nm):
            self.spect[m].correct_and_push(self.psatd[m], use_true_rho)
            self.spect[m].push_rho()

    # ---- transforms ----
    def _vec(self, fieldtype):
        return fieldtype in ('E', 'B', 'J')

    def interp2spect(self, fieldtype):
        self.join_side()
        for m in range(self.Nm):
            g, s, tr = self.interp[m], self.spect[m], self.trans[m]
            if self._vec(fieldtype):
                f = fieldtype
                tr.interp2spect_scal(getattr(g, f + 'z'), getattr(s, f + 'z'))
                tr.interp2spect_vect(getattr(g, f + 'r'), getattr(g, f + 't'),
                                     getattr(s, f + 'p'), getattr(s, f + 'm'))
            elif fieldtype in RHO_TYPES:
                tr.interp2spect_scal(g.rho, getattr(s, fieldtype))
            elif fieldtype in ('E_pml', 'B_pml'):          # fields.py:341-352
                f = fieldtype[0]
                tr.interp2spect_vect(getattr(g, f + 'r_pml'), getattr(g, f + 't_pml'),
                                     getattr(s, f + 'p_pml'), getattr(s, f + 'm_pml'))
            else:
                raise ValueError('Invalid string for fieldtype: %s' % fieldtype)

    def spect2interp(self, fieldtype):
        self.join_side()
        for m in range(self.Nm):
            g, s, tr = self.interp[m], self.spect[m], self.trans[m]
            if self._vec(fieldtype):
                f = fieldtype
                tr.spect2interp_scal(getattr(s, f + 'z'), getattr(g, f + 'z'))
                tr.spect2interp_vect(getattr(s, f + 'p'), getattr(s, f + 'm'),
                                     getattr(g, f + 'r'), getattr(g, f + 't'))
            elif fieldtype in RHO_TYPES:
                tr.spect2interp_scal(getattr(s, fieldtype), g.rho)
            elif fieldtype in ('E_pml', 'B_pml'):          # fields.py:398-409
                f = fieldtype[0]
                tr.spect2interp_vect(getattr(s, f + 'p_pml'), getattr(s, f + 'm_pml'),
                                     getattr(g, f + 'r_pml'), getattr(g, f + 't_pml'))
            else:
                raise ValueError('Invalid string for fieldtype: %s' % fieldtype)

    def _partial_pairs(self, m, fieldtype):
        g, s = self.interp[m], self.spect[m]
        if fieldtype == 'EB':       # both at once (fused step): same arrays as 'E' then 'B'
            return self._partial_pairs(m, 'E') + self._partial_pairs(m, 'B')
        if self._vec(fieldtype):
            f = fieldtype
            return [(getattr(s, f + 'z'), getattr(g, f + 'z')), (getattr(s, f + 'p'), getattr(g, f + 'r')),
                    (getattr(s, f + 'm'), getattr(g, f + 't'))]
        if fieldtype in RHO_TYPES:
            return [(getattr(s, fieldtype), g.rho)]
        raise ValueError('Invalid string for fieldtype: %s' % fieldtype)

    # ---- second stream: transforms off the critical path of the step ----
    def side_allowed(self):
        """The z-FFTs that follow the guard-cell exchange of E and B rebuild the SPECTRAL arrays, which nothing reads
        before the next field push: they may run on a second stream, under the particle kernels of the next step
        (which leave half of the DRAM bandwidth idle).  Only with the library's own transform (a cuFFT plan must
        not serve two streams at once); B2_OVERLAP_EB=0 switches it off."""
        if getattr(self, '_side_ok', None) is None:
            self._side_ok = os.environ.get('B2_OVERLAP_EB', '1') != '0' and \
                bool(_lib.load().b2_fft_has_plan(self.Nz, 23))
        return self._side_ok

    def _fft_many_side(self, pairs, inverse):
        """`_fft_many` on the second stream, ordered after everything issued so far on the context stream;
        `join_side()` orders the context stream after it."""
        ctx = _lib.context()
        if getattr(self, '_side_stream', None) is None:
            p, e0, e1 = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            call.b2_stream_create(ctypes.byref(p))
            call.b2_event_create(ctypes.byref(e0))
            call.b2_event_create(ctypes.byref(e1))
            self._side_stream, self._side_fork, self._side_join = p, e0, e1
        call.b2_event_record(self._side_fork, ctx.stream)
        call.b2_stream_wait_event(self._side_stream, self._side_fork)
        call.b2_fft_z_multi(ctx.handle, len(pairs), ptr_array([a for a, _ in pairs]),
                            ptr_array([b for _, b in pairs]), self.Nz, self.Nr, inverse, self._side_stream)
        call.b2_event_record(self._side_join, self._side_stream)
        self._side_pending = True

    def join_side(self):
        """Order the context stream after the transforms issued on the second stream (no host wait)."""
        if getattr(self, '_side_pending', False):
            call.b2_stream_wait_event(_lib.context().stream, self._side_join)
            self._side_pending = False

    def _fft_many(self, pairs, inverse):
        """z-FFTs of independent arrays in one call (spread over concurrent lanes by the library)."""
        if not pairs:
            return
        call.b2_fft_z_multi(_lib.context().handle, len(pairs), ptr_array([a for a, _ in pairs]),
                            ptr_array([b for _, b in pairs]), self.Nz, self.Nr, inverse, None)

    def spect2partial_interp(self, fieldtype):
        """iFFT along z only (fields.py:431-485)."""
        if fieldtype != 'J':
            self.join_side()
        self._fft_many([(sp, it) for m in range(self.Nm) for sp, it in self._partial_pairs(m, fieldtype)], 1)

    def partial_interp2spect(self, fieldtype, side=False):
        """FFT along z only (fields.py:487-537).  `side`: on the second stream (see side_allowed); the caller
        must not touch the spectral arrays of `fieldtype` before join_side()."""
        pairs = [(it, sp) for m in range(self.Nm) for sp, it in self._partial_pairs(m, fieldtype)]
        if side and pairs and self.side_allowed():
            self.join_side()
            self._fft_many_side(pairs, 0)
            return
        if fieldtype != 'J':
            self.join_side()
        self._fft_many(pairs, 0)

    # ---- fused transforms (single-domain fast path of Simulation.step) ----
    def _fused_tables(self, filter_currents):
        """Device copies of the Hankel matrices with the neighbouring element-wise passes folded in:
        forward (deposited sources): diag(invvol) . M . diag(filter_r)  -- divide_by_volume
        (interpolation_grid.py:271) and the radial half of filter_spect (spectral_grid.py:423) cost
        nothing; the z half of the filter is the row scale of the GEMM epilogue;
        inverse (E, B): invM / Nz -- the normalisation pass of the inverse FFT (fourier.py:157)."""
        key = bool(filter_currents)
        if getattr(self, '_fused_key', None) == key:
            return self._fused
        out = []
        for m in range(self.Nm):
            tr, g, sp = self.trans[m], self.interp[m], self.spect[m]
            fr = sp.filter_array_r if filter_currents else np.ones(self.Nr)
            fwd = lambda M: DeviceArray.from_numpy(g.invvol[:, None] * M * fr[None, :])
            inv = lambda M: DeviceArray.from_numpy(M / self.Nz)
            tr._buffers()
            bufs = [DeviceArray((self.Nz, self.Nr), np.complex128) for _ in range(6)]
            out.append(dict(F0=fwd(tr.dht0.M), Fp=fwd(tr.dhtp.M), Fm=fwd(tr.dhtm.M),
                            I0=inv(tr.dht0.invM), Ip=inv(tr.dhtp.invM), Im=inv(tr.dhtm.invM),
                            fz=(sp.d_filter_array_z if filter_currents else None), buf=bufs))
        self._fused, self._fused_key = out, key
        return out

    def fused_deposit2spect(self, fieldtype, filter_currents=True):
        """divide_by_volume + interp2spect + filter_spect of freshly deposited sources
        (main.py:640,657,666-668) as 1 (rho) or 3 (J) FFTs per mode + ONE batched Hankel launch.
        `fieldtype` may be a list (e.g. ['J', 'rho_next']: the current deposited at the half step is
        transformed together with the charge deposited at the end of the step -- nothing reads the
        spectral current in between), all of it in one FFT call and one Hankel launch."""
        fieldtypes = [fieldtype] if isinstance(fieldtype, str) else list(fieldtype)
        T = self._fused_tables(filter_currents)
        ctx = _lib.context()
        jobs, ffts = [], []
        for m in range(self.Nm):
            g, s, t = self.interp[m], self.spect[m], T[m]
            fz = t['fz'].ptr if t['fz'] is not None else None
            nbuf = 0
            for ft in fieldtypes:
                if ft == 'J':
                    bz, br, bt = t['buf'][nbuf:nbuf + 3]
                    nbuf += 3
                    ffts += [(g.Jz, bz), (g.Jr, br), (g.Jt, bt)]
                    jobs.append(DhtJob(bz.ptr, None, s.Jz.ptr, None, t['F0'].ptr, None, fz, _lib.DHT_SCALAR))
                    jobs.append(DhtJob(br.ptr, bt.ptr, s.Jp.ptr, s.Jm.ptr, t['Fp'].ptr, t['Fm'].ptr,
                                       fz, _lib.DHT_RT_TO_PM))
                elif ft in RHO_TYPES:
                    b = t['buf'][nbuf]
                    nbuf += 1
                    ffts.append((g.rho, b))
                    jobs.append(DhtJob(b.ptr, None, getattr(s, ft).ptr, None, t['F0'].ptr, None,
                                       fz, _lib.DHT_SCALAR))
                else:
                    raise ValueError('Invalid string for fieldtype: %s' % ft)
        self._fft_many(ffts, 0)
        self._dht_batch(jobs)

    def fused_spect2interp_EB(self):
        """spect2interp('E') + spect2interp('B') (main.py:768-769): batched inverse Hankel
        transforms of all modes (1/Nz folded in), then the 6*Nm unscaled inverse FFTs."""
        self.join_side()
        T = self._fused_tables(getattr(self, '_fused_key', True))
        ctx = _lib.context()
        jobs, ffts = [], []
        for m in range(self.Nm):
            g, s, t = self.interp[m], self.spect[m], T[m]
            for f, o in (('E', 0), ('B', 3)):
                bz, br, bt = t['buf'][o], t['buf'][o + 1], t['buf'][o + 2]
                jobs.append(DhtJob(getattr(s, f + 'z').ptr, None, bz.ptr, None, t['I0'].ptr, None, None,
                                   _lib.DHT_SCALAR))
                jobs.append(DhtJob(getattr(s, f + 'p').ptr, getattr(s, f + 'm').ptr, br.ptr, bt.ptr,
                                   t['Ip'].ptr, t['Im'].ptr, None, _lib.DHT_PM_TO_RT))
                ffts += [(bz, getattr(g, f + 'z')), (br, getattr(g, f + 'r')), (bt, getattr(g, f + 't'))]
        self._dht_batch(jobs)
        self._fft_many(ffts, 2)

    def fused_partial2interp_EB(self):
        """After the guard-cell exchange of exchange_and_damp_EB (main.py:741-747) the interpolation
        arrays hold E, B in (z, kr) space.  The inverse Hankel transform commutes with the z-FFT, so
        the real-space fields follow directly from those arrays -- iDHT only, (p,m)->(r,t) in the
        epilogue -- instead of spect2interp's iDHT + inverse FFT of the spectral arrays
        (main.py:768-769): 6*Nm FFTs less per step.  Call AFTER partial_interp2spect."""
        T = self._fused_tables(getattr(self, '_fused_key', True))
        ctx = _lib.context()
        jobs, swaps = [], []
        for m in range(self.Nm):
            g, tr, t = self.interp[m], self.trans[m], T[m]
            for f, o in (('E', 0), ('B', 3)):
                bz, br, bt = t['buf'][o], t['buf'][o + 1], t['buf'][o + 2]
                jobs.append(DhtJob(getattr(g, f + 'z').ptr, None, bz.ptr, None, tr.dht0.d_invM.ptr, None, None,
                                   _lib.DHT_SCALAR))
                jobs.append(DhtJob(getattr(g, f + 'r').ptr, getattr(g, f + 't').ptr, br.ptr, bt.ptr,
                                   tr.dhtp.d_invM.ptr, tr.dhtm.d_invM.ptr, None, _lib.DHT_PM_TO_RT))
                swaps += [(f + 'z', o), (f + 'r', o + 1), (f + 't', o + 2)]
            for name, o in swaps[-6:]:
                old = getattr(g, name)
                setattr(g, name, t['buf'][o])        # pointer swap: the result becomes the field array
                t['buf'][o] = old
        self._dht_batch(jobs)

    def _dht_batch(self, jobs):
        """b2_dht_batch on a job list of any length: one launch carries at most MAX_DHT_JOBS jobs (of any
        mix of kinds), longer lists are split."""
        ctx = _lib.context()
        for i in range(0, len(jobs), _lib.MAX_DHT_JOBS):
            chunk = jobs[i:i + _lib.MAX_DHT_JOBS]
            arr = (DhtJob * len(chunk))(*chunk)
            call.b2_dht_batch(ctx.handle, len(chunk), arr, self.Nz, self.Nr, None)

    def _pml_buffers(self):
        if getattr(self, '_pml_buf', None) is None:
            self._pml_buf = [[DeviceArray((self.Nz, self.Nr), np.complex128) for _ in range(4)] for m in range(self.Nm)]
        return self._pml_buf

    def fused_spect2interp_EB_pml(self):
        """spect2interp of 'E', 'B', 'E_pml', 'B_pml' (main.py:732-737) as batched inverse Hankel launches of all modes
        (1/Nz folded into the matrices) followed by one multi-lane call of unscaled inverse FFTs."""
        self.join_side()
        T = self._fused_tables(getattr(self, '_fused_key', True))
        P = self._pml_buffers()
        jobs, ffts = [], []
        for m in range(self.Nm):
            g, s, t = self.interp[m], self.spect[m], T[m]
            for f, o in (('E', 0), ('B', 3)):
                bz, br, bt = t['buf'][o], t['buf'][o + 1], t['buf'][o + 2]
                pr, pt = P[m][0 if f == 'E' else 2], P[m][1 if f == 'E' else 3]
                jobs.append(DhtJob(getattr(s, f + 'z').ptr, None, bz.ptr, None, t['I0'].ptr, None, None, _lib.DHT_SCALAR))
                jobs.append(DhtJob(getattr(s, f + 'p').ptr, getattr(s, f + 'm').ptr, br.ptr, bt.ptr,
                                   t['Ip'].ptr, t['Im'].ptr, None, _lib.DHT_PM_TO_RT))
                jobs.append(DhtJob(getattr(s, f + 'p_pml').ptr, getattr(s, f + 'm_pml').ptr, pr.ptr, pt.ptr,
                                   t['Ip'].ptr, t['Im'].ptr, None, _lib.DHT_PM_TO_RT))
                ffts += [(bz, getattr(g, f + 'z')), (br, getattr(g, f + 'r')), (bt, getattr(g, f + 't')),
                         (pr, getattr(g, f + 'r_pml')), (pt, getattr(g, f + 't_pml'))]
        self._dht_batch(jobs)
        self._fft_many(ffts, 2)

    def fused_interp2spect_EB_pml(self):
        """interp2spect of 'E', 'B', 'E_pml', 'B_pml' (main.py:757-761): one multi-lane call of forward FFTs, then
        batched forward Hankel launches with the (r,t)->(p,m) combination in their prologue."""
        self.join_side()
        T = self._fused_tables(getattr(self, '_fused_key', True))
        P = self._pml_buffers()
        jobs, ffts = [], []
        for m in range(self.Nm):
            g, s, t, tr = self.interp[m], self.spect[m], T[m], self.trans[m]
            for f, o in (('E', 0), ('B', 3)):
                bz, br, bt = t['buf'][o], t['buf'][o + 1], t['buf'][o + 2]
                pr, pt = P[m][0 if f == 'E' else 2], P[m][1 if f == 'E' else 3]
                ffts += [(getattr(g, f + 'z'), bz), (getattr(g, f + 'r'), br), (getattr(g, f + 't'), bt),
                         (getattr(g, f + 'r_pml'), pr), (getattr(g, f + 't_pml'), pt)]
                jobs.append(DhtJob(bz.ptr, None, getattr(s, f + 'z').ptr, None, tr.dht0.d_M.ptr, None, None,
                                   _lib.DHT_SCALAR))
                jobs.append(DhtJob(br.ptr, bt.ptr, getattr(s, f + 'p').ptr, getattr(s, f + 'm').ptr,
                                   tr.dhtp.d_M.ptr, tr.dhtm.d_M.ptr, None, _lib.DHT_RT_TO_PM))
                jobs.append(DhtJob(pr.ptr, pt.ptr, getattr(s, f + 'p_pml').ptr, getattr(s, f + 'm_pml').ptr,
                                   tr.dhtp.d_M.ptr, tr.dhtm.d_M.ptr, None, _lib.DHT_RT_TO_PM))
        self._fft_many(ffts, 0)
        self._dht_batch(jobs)

    # ---- interpolation-grid ops ----
    def erase(self, fieldtype):
        for m in range(self.Nm):
            self.interp[m].erase(fieldtype)

    def sum_reduce_deposition_array(self, fieldtype):
        """CPU-thread reduction of the reference (fields.py:566-594): nothing to do on GPU."""
        return

    def filter_spect(self, fieldtype):
        self.join_side()
        for m in range(self.Nm):
            self.spect[m].filter(fieldtype)

    def divide_by_volume(self, fieldtype):
        for m in range(self.Nm):
            self.interp[m].divide_by_volume(fieldtype)

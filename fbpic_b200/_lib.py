"""
ctypes binding of libfbpic_b200.so (include/fbpic_b200.h) and the `DeviceArray`
wrapper that replaces cupy arrays in the operator surface
(reference seam: fbpic/utils/cuda.py:101-182, 339-541).

There is NO host fallback: importing the operators works without a GPU (so the
host-side logic can be tested), but any compute call raises if the library or a
CUDA device is missing.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, 'libfbpic_b200.so')

c_void_p, c_int, c_double, c_int64, c_size_t = \
    ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64, ctypes.c_size_t
P = c_void_p


class SpectralMode(ctypes.Structure):
    """Mirror of `b2_spectral_mode` (include/fbpic_b200.h)."""
    _fields_ = [(n, c_void_p) for n in
                ('Ep', 'Em', 'Ez', 'Bp', 'Bm', 'Bz', 'Jp', 'Jm', 'Jz', 'rho_prev', 'rho_next',
                 'kz', 'kr', 'inv_k2', 'C', 'S_w', 'j_coef', 'rho_prev_coef', 'rho_next_coef',
                 'T_eb', 'T_cc', 'T_rho', 'j_corr_coef')] + \
               [('mu_0', c_double), ('epsilon_0', c_double)]


class DhtJob(ctypes.Structure):
    """Mirror of `b2_dht_job` (include/fbpic_b200.h)."""
    _fields_ = [('in1', c_void_p), ('in2', c_void_p), ('out1', c_void_p), ('out2', c_void_p),
                ('M1', c_void_p), ('M2', c_void_p), ('rowscale', c_void_p), ('kind', c_int)]


DHT_SCALAR, DHT_RT_TO_PM, DHT_PM_TO_RT = 0, 1, 2
MAX_DHT_JOBS = 16           # DHT_MAX_JOBS (csrc/b2_dht.cu): jobs (of any kind) carried by one batched Hankel launch
MAX_ARRAYS = 32             # B2_MAX_ARRAYS (include/fbpic_b200.h): pointers carried by one multi-array launch


# name -> argtypes; every entry returns int status unless listed in _RESTYPES
_SIGNATURES = {
    'b2_device_count': [ctypes.POINTER(c_int)],
    'b2_ctx_create': [c_int, ctypes.POINTER(P)],
    'b2_ctx_destroy': [P],
    'b2_ctx_stream': [P],
    'b2_error_string': [],
    'b2_version': [],
    'b2_malloc': [ctypes.POINTER(P), c_size_t],
    'b2_free': [P],
    'b2_host_alloc': [ctypes.POINTER(P), c_size_t],
    'b2_host_free': [P],
    'b2_memcpy_h2d': [P, P, c_size_t, P],
    'b2_memcpy_d2h': [P, P, c_size_t, P],
    'b2_memcpy_d2d': [P, P, c_size_t, P],
    'b2_memset': [P, c_int, c_size_t, P],
    'b2_stream_sync': [P],
    'b2_stream_create': [ctypes.POINTER(P)],
    'b2_stream_destroy': [P],
    'b2_stream_wait_event': [P, P],
    'b2_device_sync': [],
    'b2_event_create': [ctypes.POINTER(P)],
    'b2_event_destroy': [P],
    'b2_event_record': [P, P],
    'b2_event_elapsed_ms': [P, P, ctypes.POINTER(ctypes.c_float)],
    'b2_launch_count': [],
    'b2_profile_enable': [c_int],
    'b2_profile_reset': [],
    'b2_profile_slots': [],
    'b2_profile_name': [c_int],
    'b2_profile_read': [c_int, ctypes.POINTER(c_double), ctypes.POINTER(ctypes.c_uint64)],
    'b2_graph_begin': [P],
    'b2_graph_end': [P, ctypes.POINTER(P)],
    'b2_graph_launch': [P, P],
    'b2_graph_destroy': [P],
    'b2_cell_index': [P, c_int64, P, P, P, c_double, c_double, c_int, c_double, c_double, c_int, P, P],
    'b2_sort_cells': [P, c_int64, P, P, P, c_int, c_int, P],
    'b2_permute': [P, c_int64, P, c_int, P, P, P],
    'b2_gather': [P, c_int64, P, P, P, c_double, c_double, c_double, c_int, c_double, c_double, c_int,
                  c_int, P, c_int, P, P, P, P, P, P, P],
    'b2_push_p': [P, c_int64, P, P, P, P, P, P, P, P, P, P, c_double, c_double, c_double, P],
    'b2_push_x': [P, c_int64, P, P, P, P, P, P, P, c_double, c_double, c_double, c_double, P],
    'b2_gather_push': [P, c_int64, P, P, P, P, P, P, P, c_double, c_double, c_double, c_int, c_double,
                       c_double, c_int, c_int, P, c_int, c_double, c_double, c_double, c_double, P, c_double, P],
    'b2_push_x_key': [P, c_int64, P, P, P, P, P, P, P, c_double, c_int, c_double, c_double,
                      c_double, c_double, c_int, c_double, c_double, c_int, P, P],
    'b2_shift_periodic': [P, c_int64, P, c_double, c_double, P],
    'b2_add_scalar': [P, c_int64, P, c_double, P],
    'b2_exchange_classify': [P, c_int64, P, c_double, c_double, ctypes.POINTER(c_int64), P],
    'b2_exchange_scatter': [P, c_int64, P, c_double, c_double, c_int, P, P, P, P, P],
    'b2_deposit_rho': [P, c_int64, P, P, P, P, c_double, c_double, c_double, c_int, c_double, c_double,
                       c_int, c_int, P, P, P, P, c_int, P],
    'b2_deposit_J': [P, c_int64, P, P, P, P, c_double, P, P, P, P, c_double, c_double, c_int, c_double,
                     c_double, c_int, c_int, P, P, P, P, c_int, P],
    'b2_deposit_permute': [P, c_int, c_int64, P, P, c_double, c_double, c_double, c_int, c_double, c_double,
                           c_int, c_int, P, P, P, P, c_int, P],
    'b2_deposit_rho_displaced': [P, c_int64, P, P, P, P, c_double, c_double, c_double, c_int, c_double, c_double,
                                 c_int, c_int, P, P, P, P, P],
    'b2_push_deposit_rho': [P, c_int64, P, P, P, P, P, P, P, P, c_double, c_int, c_double, c_double, c_double,
                            c_double, c_double, c_int, c_double, c_double, c_int, c_int, P, P, P, c_int, P],
    'b2_scale_rows_by_r': [P, c_int, P, P, c_int, c_int, P],
    'b2_fft_z': [P, P, P, c_int, c_int, c_int, P],
    'b2_fft_z_multi': [P, c_int, P, P, c_int, c_int, c_int, P],
    'b2_dht': [P, P, P, P, P, c_int, c_int, P],
    'b2_dht_rt_to_pm': [P, P, P, P, P, P, P, P, c_int, c_int, P],
    'b2_dht_pm_to_rt': [P, P, P, P, P, P, P, P, c_int, c_int, P],
    'b2_dht_batch': [P, c_int, ctypes.POINTER(DhtJob), c_int, c_int, P],
    'b2_dht_flops': [],
    'b2_fft_has_plan': [c_int, c_int],
    'b2_rt_to_pm': [P, P, P, c_int, c_int, P],
    'b2_pm_to_rt': [P, P, P, c_int, c_int, P],
    'b2_filter': [P, c_int, P, P, P, c_int, c_int, P],
    'b2_correct_currents': [P, ctypes.POINTER(SpectralMode), c_int, c_double, c_int, c_int, P],
    'b2_push_eb': [P, ctypes.POINTER(SpectralMode), c_int, c_double, c_double, c_int, c_int, c_int, P],
    'b2_correct_push': [P, ctypes.POINTER(SpectralMode), c_int, c_double, c_double, c_int, c_int, c_int, P],
    'b2_damp_z': [P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P],
    'b2_shift_spect': [P, c_int, P, P, c_int, c_int, c_int, P],
    'b2_add_rows': [P, P, P, c_int, c_int, P],
    'b2_halo_stage': [P, c_int, c_int, P, c_int, c_int, c_int, P, P],
    'b2_nccl_unique_id': [P],
    'b2_nccl_init': [P, P, c_int, c_int],
    'b2_nccl_destroy': [P],
    'b2_nccl_group_start': [],
    'b2_nccl_group_end': [],
    'b2_comm_begin': [P],
    'b2_comm_end': [P],
    'b2_nccl_send': [P, P, c_size_t, c_int, P],
    'b2_nccl_recv': [P, P, c_size_t, c_int, P],
    'b2_nccl_allreduce_max_f64': [P, P, c_size_t, P],
    'b2_push_eb_pml': [P, P, P, P, P, P, P, P, P, P, P, c_int, c_int, P],
    'b2_damp_pml': [P, P, P, P, P, P, P, P, c_int, c_int, c_int, P],
    'b2_correct_currents_cross': [P, ctypes.POINTER(SpectralMode), P, P, c_int, c_double, c_int, c_int, P],
    'b2_correct_divE': [P, ctypes.POINTER(SpectralMode), c_int, c_int, P],
    'b2_push_p_after_plane': [P, c_int64, P, c_double, P, P, P, P, P, P, P, P, P, P, c_double, c_double, c_double, P],
    'b2_antenna_particles': [P, c_int64, P, P, P, P, P, P, P, c_double, P, P, P, P, P, P],
    'b2_axpy': [P, c_int64, c_double, P, P, P],
    'b2_push_p_ioniz': [P, c_int64, P, P, P, P, P, P, P, P, P, P, P, c_double, c_double, P],
    'b2_w_times_level': [P, c_int64, P, P, P, P],
    'b2_ionize': [P, c_int64, P, c_int, P, P, P, P, P, P, P, P, P, P, P, P, P, ctypes.c_uint64, c_int64, P, P,
                  ctypes.POINTER(c_int64), P],
    'b2_compton_count': [P, c_int64, P, P, P, P, P, P, P, P, ctypes.c_uint64, P, P, ctypes.POINTER(c_int64), P],
    'b2_compton_scatter': [P, c_int64, P, P, P, P, P, P, P, P, P, P, ctypes.c_uint64, P, P, P],
    'b2_select_crossing': [P, c_int64, P, P, P, c_double, c_double, c_double, c_double, c_int64, P, P,
                           ctypes.POINTER(c_int64), P],
    'b2_extract_slice': [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_double, P, P],
    'b2_external_field_compile': [ctypes.c_char_p, ctypes.POINTER(P)],
    'b2_external_field_cubin_size': [P, ctypes.POINTER(c_size_t)],
    'b2_external_field_apply': [P, P, c_int64, P, P, P, P, c_double, c_double, c_double, c_double, c_double, P],
    'b2_external_field_free': [P],
}
_RESTYPES = {'b2_dht_flops': c_double, 'b2_profile_name': ctypes.c_char_p, 'b2_profile_slots': c_int,
             'b2_error_string': ctypes.c_char_p, 'b2_version': ctypes.c_char_p,
             'b2_ctx_stream': c_void_p, 'b2_launch_count': ctypes.c_uint64}
EXPORTED = sorted(_SIGNATURES)

_lib = None


class B200Error(RuntimeError):
    pass


def load():
    """Load libfbpic_b200.so (declares argtypes).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B200Error('libfbpic_b200.so is missing: run `python -m fbpic_b200.build` '
                            '(there is no CPU fallback)')
        lib = ctypes.CDLL(SO_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, args in _SIGNATURES.items():
            f = getattr(lib, name)
            f.argtypes = args
            f.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error('libfbpic_b200: %s' % load().b2_error_string().decode())


class _Call(object):
    """`call.b2_xxx(...)`: invoke an entry point and raise on a non-zero status."""

    def __getattr__(self, name):
        f = getattr(load(), name)
        if name in _RESTYPES:
            return f

        def wrapped(*args):
            check(f(*args))
        wrapped.__name__ = name
        setattr(self, name, wrapped)
        return wrapped


call = _Call()


def device_count():
    n = c_int(0)
    try:
        rc = load().b2_device_count(ctypes.byref(n))
    except (B200Error, OSError):
        return 0
    return n.value if rc == 0 else 0


def cuda_available():
    return device_count() > 0


# ---------------------------------------------------------------------------
# context (one per process / GPU)
# ---------------------------------------------------------------------------
class Context(object):
    def __init__(self, device=0):
        self.handle = P()
        call.b2_ctx_create(device, ctypes.byref(self.handle))
        self.device = device
        self.stream = load().b2_ctx_stream(self.handle)

    def sync(self):
        call.b2_stream_sync(self.stream)


_ctx = None


def context(device=None):
    """The process-wide context; the device defaults to LOCAL_RANK (one process per GPU)."""
    global _ctx
    if _ctx is None:
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0')) % max(device_count(), 1)
        if not cuda_available():
            raise B200Error('no CUDA device: fbpic_b200 has no CPU fallback')
        _ctx = Context(device)
    return _ctx


# ---------------------------------------------------------------------------
# device arrays
# ---------------------------------------------------------------------------
class _Allocation(object):
    def __init__(self, nbytes):
        context()          # binds this process to its GPU (LOCAL_RANK) BEFORE the first cudaMalloc
        self.ptr = P()
        call.b2_malloc(ctypes.byref(self.ptr), nbytes)
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                load().b2_free(self.ptr)
        except Exception:
            pass


class _PinnedOwner(object):
    """Owns one cudaMallocHost block; returns it to the free list when the NumPy array that
    wraps it is garbage-collected."""

    def __init__(self, ptr, cap):
        self.ptr, self.cap = ptr, cap

    def __del__(self):
        try:
            _PINNED_FREE.setdefault(self.cap, []).append(self.ptr)
        except Exception:
            pass


_PINNED_FREE = {}


def pinned_empty(shape, dtype):
    """NumPy array in page-locked host memory (D2H/H2D copies run at full PCIe/NVLink-C2C speed);
    blocks are recycled through a size-keyed free list."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    # (sizes above 1 MB are rounded up to 1/8 of their power of two: a species whose particle count drifts from call
    #  to call -- moving window, injection -- still finds its buffers in the free list)
    if n < (1 << 20):
        cap = max(4096, 1 << (max(n, 1) - 1).bit_length())
    else:
        g = 1 << max(20, n.bit_length() - 4)
        cap = (n + g - 1) // g * g
    free = _PINNED_FREE.get(cap)
    if free:
        ptr = free.pop()
    else:
        p = P()
        call.b2_host_alloc(ctypes.byref(p), cap)
        ptr = p.value
    buf = (ctypes.c_char * cap).from_address(ptr)
    buf._owner = _PinnedOwner(ptr, cap)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# bytes moved by DeviceArray.set / .get since import (bench.py reports the per-call deltas)
TRANSFERRED = {'h2d': 0, 'd2h': 0}


class DeviceArray(object):
    """A C-contiguous array in HBM (or a view into one).  Mirrors the little of the
    cupy.ndarray interface that the reference's operator surface relies on:
    `.shape`, `.dtype`, `.get()`, `.fill()`, slicing along the first axis."""

    def __init__(self, shape, dtype, base=None, offset=0):
        self.shape = tuple(int(s) for s in (shape if hasattr(shape, '__len__') else (shape,)))
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        if base is None:
            base = _Allocation(max(self.nbytes, 16))
            offset = 0
        self.base = base
        self.offset = offset
        self.ptr = (base.ptr.value or 0) + offset

    # -- construction helpers
    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        d = cls(a.shape, a.dtype)
        d.set(a)
        return d

    @classmethod
    def zeros(cls, shape, dtype):
        d = cls(shape, dtype)
        d.fill(0)
        return d

    def view(self, shape, dtype=None, byte_offset=0):
        v = DeviceArray(shape, dtype or self.dtype, base=self.base, offset=self.offset + byte_offset)
        if v.offset + v.nbytes > self.base.nbytes:
            raise B200Error('DeviceArray.view: %d bytes at offset %d exceed the %d-byte allocation'
                            % (v.nbytes, v.offset, self.base.nbytes))
        return v

    @property
    def capacity(self):
        """Number of elements of this dtype that fit between this view's start and the end of the
        underlying allocation (per-particle arrays are allocated with headroom)."""
        return (self.base.nbytes - self.offset) // self.dtype.itemsize

    def __getitem__(self, key):
        """Row-range view: a[i0:i1] along the first axis (contiguous)."""
        if not isinstance(key, slice):
            raise TypeError('DeviceArray supports only first-axis slices')
        i0, i1, step = key.indices(self.shape[0])
        assert step == 1
        row = self.dtype.itemsize * int(np.prod(self.shape[1:])) if len(self.shape) > 1 else self.dtype.itemsize
        return self.view((max(i1 - i0, 0),) + self.shape[1:], byte_offset=i0 * row)

    def __len__(self):
        return self.shape[0]

    # -- transfers
    def set(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        assert a.size == self.size, (a.shape, self.shape)
        ctx = context()
        call.b2_memcpy_h2d(self.ptr, a.ctypes.data, self.nbytes, ctx.stream)
        call.b2_stream_sync(ctx.stream)
        TRANSFERRED['h2d'] += self.nbytes

    def get(self, pinned=False):
        out = pinned_empty(self.shape, self.dtype) if (pinned and self.nbytes >= (1 << 16)) \
            else np.empty(self.shape, dtype=self.dtype)
        ctx = context()
        call.b2_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, ctx.stream)
        call.b2_stream_sync(ctx.stream)
        TRANSFERRED['d2h'] += self.nbytes
        return out

    def fill(self, value):
        assert value == 0, 'only zero fill is supported'
        call.b2_memset(self.ptr, 0, self.nbytes, context().stream)

    def copy_from(self, other):
        assert other.nbytes == self.nbytes
        call.b2_memcpy_d2d(self.ptr, other.ptr, self.nbytes, context().stream)

    def copy(self):
        d = DeviceArray(self.shape, self.dtype)
        d.copy_from(self)
        return d


def ptr_array(arrays):
    """Host array of device pointers (void*[n]) for the multi-array entry points."""
    return (c_void_p * len(arrays))(*[a.ptr if isinstance(a, DeviceArray) else a for a in arrays])


def to_device(a):
    return a if isinstance(a, DeviceArray) else DeviceArray.from_numpy(a)


def to_host(a):
    """Device -> host; large arrays land in recycled page-locked buffers."""
    return a.get(pinned=True) if isinstance(a, DeviceArray) else a

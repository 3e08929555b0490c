"""
tests/fake_device.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A stand-in for libfbpic_b200.so that keeps "device" memory in host RAM and executes every C-ABI entry
point with the oracle (oracle/) or NumPy.  Its only purpose is to exercise the HOST-SIDE logic of the
operator surface (`Simulation.step` call order, PML / cross-deposition / antenna flows, sort-state
bookkeeping, particle exchange on one rank) in the GPU-less build container: a pytest fixture swaps it
in for `fbpic_b200._lib._lib`.  The product has no switch, environment variable or import that reaches
this file; without the fixture every compute call still fails loudly when there is no GPU
(tests/test_abi.py::test_no_cpu_fallback).  Kernels are NOT validated here -- that is what the `-m gpu`
tests (through the real library) and tests/hostemu (kernel source on the CPU) are for.

Semantics follow include/fbpic_b200.h entry by entry.
"""
import ctypes
import os
import subprocess
import numpy as np
import scipy.fft as sfft

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tests', 'hostemu')


def build_emu():
    """Compile (if stale) and load tests/hostemu/libemu_ext.so: the kernel source of b2_ext_kernels.cuh for the CPU."""
    so = os.path.join(EMU_DIR, 'libemu_ext.so')
    srcs = [os.path.join(EMU_DIR, 'emu_ext.cpp'), os.path.join(EMU_DIR, 'cuda_shim.h'),
            os.path.join(ROOT, 'fbpic_b200', 'csrc', 'b2_ext_kernels.cuh')]
    if (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        # the namespace is renamed and references are bound inside the library (-Bsymbolic): libfbpic_b200.so
        # is loaded RTLD_GLOBAL and exports host launch stubs with the same mangled kernel names
        subprocess.check_call(['g++', '-O1', '-ffp-contract=off', '-Db2ext=b2ext_hostemu', '-shared', '-fPIC',
                               '-Wl,-Bsymbolic', '-o', so, srcs[0]])
    return ctypes.CDLL(so)


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if hasattr(p, 'value'):
        return p.value or 0
    return ctypes.addressof(p)


def _arr(p, n, dtype=np.float64):
    """NumPy view of n elements at address p."""
    a = _addr(p)
    if n == 0 or a == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * nbytes).from_address(a)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


def _grid(p, Nz, Nr):
    return _arr(p, Nz * Nr, np.complex128).reshape(Nz, Nr)


def _ptrs(p, n):
    return [(_addr(v) if v is not None else 0) for v in list(p)[:n]] if p is not None else [0] * n


class FakeLib(object):
    def __init__(self):
        self._mem = {}
        self._perm = None
        self._part = None
        self.launches = 0
        self.calls = []
        self.emu = build_emu()

    def __getattr__(self, name):
        raise AttributeError('fake device: %s is not emulated' % name)

    # ---------------------------------------------------------------- runtime
    def b2_device_count(self, p):
        p._obj.value = 1
        return 0

    def b2_ctx_create(self, device, p):
        p._obj.value = 0xB200
        return 0

    def b2_ctx_destroy(self, ctx):
        return 0

    def b2_ctx_stream(self, ctx):
        return 1

    def b2_error_string(self):
        return b'fake device'

    def b2_version(self):
        return b'fake'

    def b2_launch_count(self):
        return self.launches

    def b2_malloc(self, p, nbytes):
        buf = np.zeros(int(nbytes) + int(os.environ.get('B2_FAKE_PAD', '64')), dtype=np.uint8)
        a = buf.ctypes.data
        self._mem[a] = buf
        p._obj.value = a
        return 0

    def b2_free(self, ptr):
        self._mem.pop(_addr(ptr), None)
        return 0

    b2_host_alloc = b2_malloc
    b2_host_free = b2_free

    def _copy(self, dst, src, nbytes, stream):
        if nbytes:
            ctypes.memmove(_addr(dst), _addr(src), int(nbytes))
        return 0

    b2_memcpy_h2d = b2_memcpy_d2h = b2_memcpy_d2d = _copy

    def b2_memset(self, ptr, value, nbytes, stream):
        if nbytes:
            ctypes.memset(_addr(ptr), value, int(nbytes))
        return 0

    def b2_stream_sync(self, stream):
        return 0

    def b2_fft_has_plan(self, Nz, max_radix):
        return 1

    # a second stream: the stand-in executes every call at once, in program order
    def b2_stream_create(self, p):
        p._obj.value = 0x5717
        return 0

    def b2_stream_destroy(self, stream):
        return 0

    def b2_stream_wait_event(self, stream, e):
        return 0

    # events and the per-kernel profiler: host clock, no per-kernel records
    def b2_event_create(self, p):
        p._obj.value = len(self.calls) + 1
        self.calls.append(0.)
        return 0

    def b2_event_destroy(self, e):
        return 0

    def b2_event_record(self, e, stream):
        import time
        self.calls[_addr(e) - 1] = time.perf_counter()
        return 0

    def b2_event_elapsed_ms(self, a, b, ms):
        ms._obj.value = 1e3 * (self.calls[_addr(b) - 1] - self.calls[_addr(a) - 1])
        return 0

    def b2_profile_enable(self, on):
        return 0

    def b2_profile_reset(self):
        return 0

    def b2_profile_slots(self):
        return 0

    def b2_profile_name(self, slot):
        return b''

    def b2_profile_read(self, slot, total_ms, count):
        total_ms._obj.value, count._obj.value = 0., 0
        return 0

    def b2_device_sync(self):
        return 0

    # ---------------------------------------------------------------- particles
    def b2_cell_index(self, ctx, n, x, y, z, invdz, zmin, Nz, invdr, rmin, Nr, cell_idx, stream):
        _arr(cell_idx, n, np.int32)[:] = orc.cell_index(_arr(x, n), _arr(y, n), _arr(z, n), invdz, zmin, Nz,
                                                        invdr, rmin, Nr)
        return 0

    def b2_sort_cells(self, ctx, n, cell_idx, sorted_idx, prefix_sum, Nz, Nr, stream):
        keys = _arr(cell_idx, n, np.int32)
        perm, prefix = orc.sort_contract(keys, Nz, Nr) if n else (np.zeros(0, np.int64),
                                                                  np.zeros(Nz * (Nr + 1), np.int32))
        self._perm = perm.copy()
        self._perm_prefix = _addr(prefix_sum)          # as the library: the cached permutation belongs to this array
        if _addr(sorted_idx):
            _arr(sorted_idx, n, np.int64)[:] = perm
            keys[:] = keys[perm]
        _arr(prefix_sum, Nz * (Nr + 1), np.int32)[:] = prefix
        return 0

    def b2_permute(self, ctx, n, sorted_idx, n_arrays, src, dst, stream):
        self._check_na(n_arrays)
        perm = _arr(sorted_idx, n, np.int64) if _addr(sorted_idx) else self._perm
        assert perm is not None and len(perm) == n
        span = int(perm.max()) + 1 if n else 0        # dst[i] = src[idx[i]]: a gather, the source may be longer
        for s, d in zip(_ptrs(src, n_arrays), _ptrs(dst, n_arrays)):
            _arr(d, n)[:] = _arr(s, span)[perm]
        return 0

    @staticmethod
    def _check_na(na):
        # the real library refuses more pointers than one multi-array launch carries (B2_MAX_ARRAYS, fbpic_b200.h)
        assert na <= 32, 'too many arrays (B2_MAX_ARRAYS = 32)'

    def _grids(self, grids, count, Nz, Nr):
        return [_grid(p, Nz, Nr) for p in _ptrs(grids, count)]

    def b2_gather(self, ctx, n, x, y, z, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, cubic,
                  Ex, Ey, Ez, Bx, By, Bz, stream):
        G = self._grids(grids, 6 * Nm, Nz, Nr)
        orc.gather(_arr(x, n), _arr(y, n), _arr(z, n), rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr,
                   [tuple(G[6 * m:6 * m + 6]) for m in range(Nm)], bool(cubic),
                   _arr(Ex, n), _arr(Ey, n), _arr(Ez, n), _arr(Bx, n), _arr(By, n), _arr(Bz, n))
        return 0

    def b2_push_p(self, ctx, n, ux, uy, uz, ig, Ex, Ey, Ez, Bx, By, Bz, q, m, dt, stream):
        orc.push_p(_arr(ux, n), _arr(uy, n), _arr(uz, n), _arr(ig, n), _arr(Ex, n), _arr(Ey, n), _arr(Ez, n),
                   _arr(Bx, n), _arr(By, n), _arr(Bz, n), q, m, dt)
        return 0

    def b2_push_x(self, ctx, n, x, y, z, ux, uy, uz, ig, dt, xp, yp, zp, stream):
        orc.push_x(_arr(x, n), _arr(y, n), _arr(z, n), _arr(ux, n), _arr(uy, n), _arr(uz, n), _arr(ig, n),
                   dt, xp, yp, zp)
        return 0

    def b2_gather_push(self, ctx, n, x, y, z, ux, uy, uz, ig, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr,
                       Nm, grids, cubic, q, m, dt_p, dt_x, cell_idx, key_zmin, stream):
        F = [np.zeros(n) for _ in range(6)]
        fp = [f.ctypes.data for f in F]
        self.b2_gather(ctx, n, x, y, z, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, cubic,
                       *fp, stream)
        self.b2_push_p(ctx, n, ux, uy, uz, ig, *fp, q, m, dt_p, stream)
        self.b2_push_x(ctx, n, x, y, z, ux, uy, uz, ig, dt_x, 1., 1., 1., stream)
        if _addr(cell_idx):
            self.b2_cell_index(ctx, n, x, y, z, invdz, key_zmin, Nz, invdr, rmin, Nr, cell_idx, stream)
        return 0

    def b2_push_x_key(self, ctx, n, x, y, z, ux, uy, uz, ig, dt, wrap, wzmin, wzmax, invdz, key_zmin, Nz,
                      invdr, rmin, Nr, cell_idx, stream):
        self.b2_push_x(ctx, n, x, y, z, ux, uy, uz, ig, dt, 1., 1., 1., stream)
        if wrap:
            orc.shift_periodic(_arr(z, n), wzmin, wzmax)
        if _addr(cell_idx):
            self.b2_cell_index(ctx, n, x, y, z, invdz, key_zmin, Nz, invdr, rmin, Nr, cell_idx, stream)
        return 0

    def b2_shift_periodic(self, ctx, n, z, zmin, zmax, stream):
        orc.shift_periodic(_arr(z, n), zmin, zmax)
        return 0

    def b2_add_scalar(self, ctx, n, v, value, stream):
        _arr(v, n)[:] += value
        return 0

    def b2_exchange_classify(self, ctx, n, z, zlo, zhi, counts, stream):
        zz = _arr(z, n)
        left, right = zz < zlo, zz > zhi
        self._part = (n, left.copy(), right.copy())
        counts[0], counts[1], counts[2] = int(n - left.sum() - right.sum()), int(left.sum()), int(right.sum())
        return 0

    def b2_exchange_scatter(self, ctx, n, z, zlo, zhi, n_arrays, src, stay, left, right, stream):
        self._check_na(n_arrays)
        assert self._part is not None and self._part[0] == n
        _, l, r = self._part
        s = ~(l | r)
        ls = _ptrs(left, n_arrays) if left is not None else [0] * n_arrays
        rs = _ptrs(right, n_arrays) if right is not None else [0] * n_arrays
        for k, (a, st) in enumerate(zip(_ptrs(src, n_arrays), _ptrs(stay, n_arrays))):
            v = _arr(a, n)
            _arr(st, int(s.sum()))[:] = v[s]
            if ls[k]:
                _arr(ls[k], int(l.sum()))[:] = v[l]
            if rs[k]:
                _arr(rs[k], int(r.sum()))[:] = v[r]
        return 0

    def _deposit(self, what, n, x, y, z, w, q, ux, uy, uz, ig, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                 r0, rh, cubic):
        if n <= 0:
            return 0
        ncomp = 3 if what == 'J' else 1
        G = self._grids(grids, ncomp * Nm, Nz, Nr)
        xx = _arr(x, n)
        raw = orc.deposit(what, xx, _arr(y, n), _arr(z, n), _arr(w, n), q,
                          _arr(ux, n) if what == 'J' else xx, _arr(uy, n) if what == 'J' else xx,
                          _arr(uz, n) if what == 'J' else xx, _arr(ig, n) if what == 'J' else xx,
                          invdz, zmin, Nz, invdr, rmin, Nr, Nm, bool(cubic),
                          _arr(r0, Nr + 1).copy(), _arr(rh, Nr + 1).copy(), 2)
        for m in range(Nm):
            for k in range(ncomp):
                G[ncomp * m + k] += raw[k, m]
        return 0

    def b2_deposit_rho(self, ctx, n, x, y, z, w, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, prefix, r0, rh,
                       cubic, stream):
        return self._deposit('rho', n, x, y, z, w, q, 0, 0, 0, 0, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                             r0, rh, cubic)

    def b2_deposit_J(self, ctx, n, x, y, z, w, q, ux, uy, uz, ig, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                     prefix, r0, rh, cubic, stream):
        return self._deposit('J', n, x, y, z, w, q, ux, uy, uz, ig, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                             r0, rh, cubic)

    def b2_deposit_permute(self, ctx, what, n, src8, dst8, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                           prefix, r0, rh, cubic, stream):
        if n <= 0:
            return 0
        assert self._perm is not None and len(self._perm) == n and _addr(prefix) == self._perm_prefix, \
            'b2_deposit_permute: no matching b2_sort_cells result'
        s, d = _ptrs(src8, 8), _ptrs(dst8, 8)
        for a, b in zip(s, d):
            _arr(b, n)[:] = _arr(a, n)[self._perm]
        x, y, z, w, ux, uy, uz, ig = d
        return self._deposit('J' if what else 'rho', n, x, y, z, w, q, ux, uy, uz, ig, invdz, zmin, Nz, invdr,
                             rmin, Nr, Nm, grids, r0, rh, cubic)

    def b2_deposit_rho_displaced(self, ctx, n, x, y, z, w, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, prefix,
                                 r0, rh, stream):
        return self._deposit('rho', n, x, y, z, w, q, 0, 0, 0, 0, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                             r0, rh, 0)

    def b2_push_deposit_rho(self, ctx, n, x, y, z, w, ux, uy, uz, ig, dt, wrap, wzmin, wzmax, q, invdz, zmin, Nz,
                            invdr, rmin, Nr, Nm, grids, r0, rh, cubic, stream):
        self.b2_push_x_key(ctx, n, x, y, z, ux, uy, uz, ig, dt, wrap, wzmin, wzmax, invdz, zmin, Nz, invdr, rmin,
                           Nr, None, stream)
        return self._deposit('rho', n, x, y, z, w, q, 0, 0, 0, 0, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                             r0, rh, cubic)

    # ---------------------------------------------------------------- grids
    def b2_scale_rows_by_r(self, ctx, na, arrays, v, Nz, Nr, stream):
        self._check_na(na)
        s = _arr(v, Nr)
        for a in self._grids(arrays, na, Nz, Nr):
            a *= s[None, :]
        return 0

    def b2_filter(self, ctx, na, arrays, fz, fr, Nz, Nr, stream):
        self._check_na(na)
        f = _arr(fz, Nz)[:, None] * _arr(fr, Nr)[None, :]
        for a in self._grids(arrays, na, Nz, Nr):
            a *= f
        return 0

    def b2_fft_z(self, ctx, a_in, a_out, Nz, Nr, inverse, stream):
        src = _grid(a_in, Nz, Nr)
        if inverse == 0:
            res = sfft.fft(src, axis=0)
        else:
            res = sfft.ifft(src, axis=0)
            if inverse == 2:
                res = res * Nz
        _grid(a_out, Nz, Nr)[:, :] = res
        return 0

    def b2_fft_z_multi(self, ctx, na, a_in, a_out, Nz, Nr, inverse, stream):
        for i, o in zip(_ptrs(a_in, na), _ptrs(a_out, na)):
            self.b2_fft_z(ctx, i, o, Nz, Nr, inverse, stream)
        return 0

    @staticmethod
    def _rs(rowscale, Nz):
        return _arr(rowscale, Nz)[:, None] if _addr(rowscale) else 1.

    def b2_dht(self, ctx, a_in, a_out, M, rowscale, Nz, Nr, stream):
        res = self._rs(rowscale, Nz) * orc.OracleTransformer.dht(_grid(a_in, Nz, Nr), _arr(M, Nr * Nr).reshape(Nr, Nr))
        _grid(a_out, Nz, Nr)[:, :] = res
        return 0

    def b2_dht_rt_to_pm(self, ctx, r, t, out_p, out_m, Mp, Mm, rowscale, Nz, Nr, stream):
        vr, vt = _grid(r, Nz, Nr).copy(), _grid(t, Nz, Nr).copy()
        rs = self._rs(rowscale, Nz)
        dht = orc.OracleTransformer.dht
        _grid(out_p, Nz, Nr)[:, :] = rs * dht(0.5 * (vr - 1.j * vt), _arr(Mp, Nr * Nr).reshape(Nr, Nr))
        _grid(out_m, Nz, Nr)[:, :] = rs * dht(0.5 * (vr + 1.j * vt), _arr(Mm, Nr * Nr).reshape(Nr, Nr))
        return 0

    def b2_dht_pm_to_rt(self, ctx, p, m, out_r, out_t, iMp, iMm, rowscale, Nz, Nr, stream):
        dht = orc.OracleTransformer.dht
        P = dht(_grid(p, Nz, Nr), _arr(iMp, Nr * Nr).reshape(Nr, Nr))
        Q = dht(_grid(m, Nz, Nr), _arr(iMm, Nr * Nr).reshape(Nr, Nr))
        rs = self._rs(rowscale, Nz)
        _grid(out_r, Nz, Nr)[:, :] = rs * (P + Q)
        _grid(out_t, Nz, Nr)[:, :] = rs * (1.j * (P - Q))
        return 0

    def b2_dht_batch(self, ctx, njobs, jobs, Nz, Nr, stream):
        assert njobs <= 16, 'b2_dht_batch: more than DHT_MAX_JOBS = 16 jobs in one launch'
        for k in range(njobs):
            j = jobs[k]
            if j.kind == 0:
                self.b2_dht(ctx, j.in1, j.out1, j.M1, j.rowscale, Nz, Nr, stream)
            elif j.kind == 1:
                self.b2_dht_rt_to_pm(ctx, j.in1, j.in2, j.out1, j.out2, j.M1, j.M2, j.rowscale, Nz, Nr, stream)
            else:
                self.b2_dht_pm_to_rt(ctx, j.in1, j.in2, j.out1, j.out2, j.M1, j.M2, j.rowscale, Nz, Nr, stream)
        return 0

    def b2_dht_flops(self):
        return 0.

    # spectral kernels: the oracle's statements on the arrays of the mode struct
    def _spectral(self, mode, comoving, dt, V, use_true_rho, Nz, Nr, correct, push):
        M = mode._obj
        sim = orc.OracleSim.__new__(orc.OracleSim)
        sim.Nm, sim.Nz, sim.Nr, sim.dt = 1, Nz, Nr, dt
        sim.v_comoving = (V if V is not None else 1.) if comoving else None
        sim.use_pml = False
        sim.current_correction = 'curl-free'
        sim.kz, sim.kr = _arr(M.kz, Nz), [_arr(M.kr, Nr)]
        sim.inv_k2 = [_arr(M.inv_k2, Nz * Nr).reshape(Nz, Nr)]
        ct = np.complex128 if comoving else np.float64
        coef = dict(C=_arr(M.C, Nz * Nr).reshape(Nz, Nr), S_w=_arr(M.S_w, Nz * Nr).reshape(Nz, Nr))
        for k in ('j_coef', 'rho_prev_coef', 'rho_next_coef'):
            coef[k] = _arr(getattr(M, k), Nz * Nr, ct).reshape(Nz, Nr)
        if comoving:
            for k in ('T_eb', 'T_cc', 'T_rho', 'j_corr_coef'):
                coef[k] = _arr(getattr(M, k), Nz * Nr, np.complex128).reshape(Nz, Nr)
        sim.coef = [coef]
        sim.spect = [{k: _grid(getattr(M, k), Nz, Nr) for k in orc.OracleSim.FIELDS_S}]
        if correct:
            sim.correct_currents()
        if push:
            sim.push_eb(bool(use_true_rho))
        return 0

    def b2_correct_currents(self, ctx, mode, comoving, inv_dt, Nz, Nr, stream):
        return self._spectral(mode, comoving, 1. / inv_dt, None, 0, Nz, Nr, True, False)

    def b2_push_eb(self, ctx, mode, comoving, dt, V, use_true_rho, Nz, Nr, stream):
        return self._spectral(mode, comoving, dt, V, use_true_rho, Nz, Nr, False, True)

    def b2_correct_push(self, ctx, mode, comoving, dt, V, use_true_rho, Nz, Nr, stream):
        return self._spectral(mode, comoving, dt, V, use_true_rho, Nz, Nr, True, True)

    def b2_damp_z(self, ctx, na, arrays, damp, nd, left, right, Nz, Nr, stream):
        self._check_na(na)
        d = _arr(damp, nd)
        for a in self._grids(arrays, na, Nz, Nr):
            if left:
                a[:nd] *= d[:, None]
            if right:
                a[Nz - nd:] *= d[::-1, None]
        return 0

    def b2_shift_spect(self, ctx, na, arrays, shift, n_move, Nz, Nr, stream):
        self._check_na(na)
        sft = _arr(shift, Nz, np.complex128)
        pw = np.ones(Nz, dtype=np.complex128)
        for _ in range(abs(n_move)):
            pw = pw * sft
        if n_move < 0:
            pw = pw.conj()
        for a in self._grids(arrays, na, Nz, Nr):
            a *= pw[:, None]
        return 0

    def b2_add_rows(self, ctx, dst, src, nrows, Nr, stream):
        _grid(dst, nrows, Nr)[:, :] += _grid(src, nrows, Nr)
        return 0

    # ---------------------------------------------------------------- z-slab exchange: gloo stands in for NCCL
    def b2_halo_stage(self, ctx, mode, na, arrays, row0, nrow, Nr, packed, stream):
        self._check_na(na)
        if na <= 0 or nrow <= 0:
            return 0
        pk = _arr(packed, na * nrow * Nr, np.complex128).reshape(na, nrow, Nr)
        for k, a in enumerate(_ptrs(arrays, na)):
            rows = _arr(a + 16 * row0 * Nr, nrow * Nr, np.complex128).reshape(nrow, Nr)
            if mode == 0:
                pk[k] = rows
            elif mode == 1:
                rows[:, :] = pk[k]
            else:
                rows[:, :] += pk[k]
        return 0

    def b2_nccl_unique_id(self, ident):
        ctypes.memset(_addr(ident), 0, 128)
        return 0

    def b2_nccl_init(self, ctx, ident, rank, size):
        self._rank, self._size, self._group = rank, size, None
        return 0

    def b2_nccl_destroy(self, ctx):
        return 0

    def b2_nccl_group_start(self):
        self._group = []
        return 0

    def b2_comm_begin(self, ctx):
        return self.b2_nccl_group_start()

    def _p2p(self, kind, buf, nbytes, peer):
        op = (kind, _addr(buf), int(nbytes), int(peer))
        if getattr(self, '_group', None) is not None:
            self._group.append(op)
        else:
            self._run_ops([op])
        return 0

    def b2_nccl_send(self, ctx, buf, nbytes, peer, stream):
        return self._p2p('send', buf, nbytes, peer)

    def b2_nccl_recv(self, ctx, buf, nbytes, peer, stream):
        return self._p2p('recv', buf, nbytes, peer)

    def _run_ops(self, ops):
        # NCCL pairs the k-th send to a peer with the k-th receive posted for that peer: same rule, as tags
        import torch
        import torch.distributed as dist
        reqs, keep, count = [], [], {}
        for kind, addr, nbytes, peer in ops:
            if nbytes == 0:
                continue
            k = count.get((kind, peer), 0)
            count[(kind, peer)] = k + 1
            t = torch.from_numpy(_arr(addr, nbytes, np.uint8))
            if kind == 'send':
                t = t.clone()
                reqs.append(dist.isend(t, peer, tag=k))
            else:
                reqs.append(dist.irecv(t, peer, tag=k))
            keep.append(t)
        for r in reqs:
            r.wait()

    def b2_nccl_group_end(self):
        ops, self._group = self._group or [], None
        self._run_ops(ops)
        return 0

    def b2_comm_end(self, ctx):
        return self.b2_nccl_group_end()

    # ---------------------------------------------------------------- solver variants: the real kernel
    # source of fbpic_b200/csrc/b2_ext_kernels.cuh, compiled for the host by tests/hostemu
    def b2_push_eb_pml(self, ctx, Ep, Em, Bp, Bm, Ez, Bz, C, S_w, T_eb, kr, Nz, Nr, stream):
        V = ctypes.c_void_p
        return self.emu.emu_push_eb_pml(V(_addr(Ep)), V(_addr(Em)), V(_addr(Bp)), V(_addr(Bm)), V(_addr(Ez)),
                                        V(_addr(Bz)), V(_addr(C)), V(_addr(S_w)), V(_addr(T_eb) or None),
                                        V(_addr(kr)), Nz, Nr)

    def b2_damp_pml(self, ctx, Et, Et_pml, Ez, Bt, Bt_pml, Bz, damp, n_pml, Nz, Nr, stream):
        V = ctypes.c_void_p
        return self.emu.emu_damp_pml(V(_addr(Et)), V(_addr(Et_pml)), V(_addr(Ez)), V(_addr(Bt)), V(_addr(Bt_pml)),
                                     V(_addr(Bz)), V(_addr(damp)), n_pml, Nz, Nr)

    def b2_correct_currents_cross(self, ctx, mode, rho_next_z, rho_next_xy, comoving, inv_dt, Nz, Nr, stream):
        M, V = mode._obj, ctypes.c_void_p
        return self.emu.emu_correct_currents_cross(
            V(M.rho_prev), V(M.rho_next), V(_addr(rho_next_z)), V(_addr(rho_next_xy)), V(M.Jp), V(M.Jm), V(M.Jz),
            V(M.kz), V(M.kr), V(M.T_cc), V(M.j_corr_coef), V(M.T_eb), int(comoving), ctypes.c_double(inv_dt), Nz, Nr)

    def b2_correct_divE(self, ctx, mode, Nz, Nr, stream):
        M, V = mode._obj, ctypes.c_void_p
        return self.emu.emu_correct_divE(V(M.Ep), V(M.Em), V(M.Ez), V(M.rho_prev), V(M.kz), V(M.kr), V(M.inv_k2),
                                         ctypes.c_double(1. / M.epsilon_0), Nz, Nr)

    def b2_push_p_after_plane(self, ctx, n, z, z_plane, ux, uy, uz, ig, Ex, Ey, Ez, Bx, By, Bz, q, m, dt, stream):
        V, D = ctypes.c_void_p, ctypes.c_double
        c = 299792458.
        return self.emu.emu_push_p_after_plane(ctypes.c_longlong(n), V(_addr(z)), D(z_plane),
                                               *[V(_addr(p)) for p in (ux, uy, uz, ig, Ex, Ey, Ez, Bx, By, Bz)],
                                               D(q * dt / (m * c)), D(0.5 * q * dt / m))

    def b2_antenna_particles(self, ctx, n, bx, by, ex, ey, vx, vy, vz, sign, x, y, ux, uy, uz, stream):
        V = ctypes.c_void_p
        return self.emu.emu_antenna_particles(ctypes.c_longlong(n), *[V(_addr(p)) for p in (bx, by, ex, ey, vx, vy, vz)],
                                              ctypes.c_double(sign), *[V(_addr(p)) for p in (x, y, ux, uy, uz)])

    # external fields: the CUDA C body produced by fbpic_b200.lpa_utils.external_fields, compiled for the host
    def b2_external_field_compile(self, body, handle):
        import hashlib
        import tempfile
        body = body.decode() if isinstance(body, bytes) else body
        src = HOST_EXT_FIELD_TEMPLATE % body
        d = os.path.join(tempfile.gettempdir(), 'b2_fake_extfield')
        os.makedirs(d, exist_ok=True)
        base = os.path.join(d, hashlib.sha1(src.encode()).hexdigest())
        if not os.path.exists(base + '.so'):
            with open(base + '.cpp', 'w') as f:
                f.write(src)
            subprocess.check_call(['g++', '-O1', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC',
                                   '-o', base + '.so', base + '.cpp'])
        self._ext = getattr(self, '_ext', {})
        key = len(self._ext) + 1
        self._ext[key] = ctypes.CDLL(base + '.so')
        handle._obj.value = key
        return 0

    def b2_external_field_apply(self, ctx, handle, n, F, x, y, z, t, amplitude, length_scale, gamma_b, beta_b,
                                stream):
        lib = self._ext[_addr(handle)]
        V, D = ctypes.c_void_p, ctypes.c_double
        lib.apply(ctypes.c_longlong(n), V(_addr(F)), V(_addr(x)), V(_addr(y)), V(_addr(z)), D(t), D(amplitude),
                  D(length_scale), D(gamma_b), D(beta_b))
        return 0

    def b2_external_field_free(self, handle):
        return 0

    def b2_axpy(self, ctx, n, a, x, y, stream):
        return self.emu.emu_axpy(ctypes.c_longlong(n), ctypes.c_double(a), ctypes.c_void_p(_addr(x)),
                                 ctypes.c_void_p(_addr(y)))

    def b2_push_p_ioniz(self, ctx, n, level, ux, uy, uz, ig, Ex, Ey, Ez, Bx, By, Bz, m, dt, stream):
        V, D = ctypes.c_void_p, ctypes.c_double
        e, c = 1.602176634e-19, 299792458.
        return self.emu.emu_push_p_ioniz(ctypes.c_longlong(n), V(_addr(level)),
                                         *[V(_addr(p)) for p in (ux, uy, uz, ig, Ex, Ey, Ez, Bx, By, Bz)],
                                         D(e * dt / (m * c)), D(0.5 * e * dt / m))

    def b2_w_times_level(self, ctx, n, w, level, out, stream):
        V = ctypes.c_void_p
        return self.emu.emu_w_times_level(ctypes.c_longlong(n), V(_addr(w)), V(_addr(level)), V(_addr(out)))

    def b2_ionize(self, ctx, n, level, level_max, pre, pw, ex, ux, uy, uz, Ex, Ey, Ez, Bx, By, Bz, draws, seed, cap,
                  events, d_count, h_count, stream):
        if not _addr(d_count):
            return -3
        V, D = ctypes.c_void_p, ctypes.c_double
        rc = self.emu.emu_ionize(ctypes.c_longlong(n), V(_addr(level)), level_max, V(_addr(pre)), V(_addr(pw)),
                                 V(_addr(ex)), *[V(_addr(p)) for p in (ux, uy, uz, Ex, Ey, Ez, Bx, By, Bz)],
                                 D(299792458.), V(_addr(draws)), ctypes.c_ulonglong(seed), ctypes.c_longlong(cap),
                                 V(_addr(events)), V(_addr(d_count)))
        h_count._obj.value = int(_arr(d_count, 1, np.int64)[0])
        return rc

    def b2_compton_count(self, ctx, n, x, y, z, ux, uy, uz, ig, params20, seed, nscatter, d_total, h_total, stream):
        V = ctypes.c_void_p
        rc = self.emu.emu_compton_count(ctypes.c_longlong(n), *[V(_addr(p)) for p in (x, y, z, ux, uy, uz, ig)],
                                        V(_addr(params20)), ctypes.c_ulonglong(seed), V(_addr(nscatter)),
                                        V(_addr(d_total)))
        h_total._obj.value = int(_arr(d_total, 1, np.int64)[0])
        return rc

    def b2_compton_scatter(self, ctx, n, nscatter, x, y, z, ux, uy, uz, ig, w, params20, seed, photon8, cursor, stream):
        V = ctypes.c_void_p
        ptrs = (ctypes.c_void_p * 8)(*_ptrs(photon8, 8))
        return self.emu.emu_compton_scatter(ctypes.c_longlong(n), V(_addr(nscatter)),
                                            *[V(_addr(p)) for p in (x, y, z, ux, uy, uz, ig, w)], V(_addr(params20)),
                                            ctypes.c_ulonglong(seed), ptrs, V(_addr(cursor)))

    def b2_select_crossing(self, ctx, n, z, uz, ig, c_light, dt, z_curr, z_prev, cap, idx, d_count, h_count, stream):
        if not _addr(d_count):
            return -3
        D, V = ctypes.c_double, ctypes.c_void_p
        rc = self.emu.emu_select_crossing(ctypes.c_longlong(n), V(_addr(z)), V(_addr(uz)), V(_addr(ig)), D(c_light),
                                          D(dt), D(z_curr), D(z_prev), ctypes.c_longlong(cap), V(_addr(idx)),
                                          V(_addr(d_count)))
        h_count._obj.value = int(_arr(d_count, 1, np.int64)[0])
        return rc

    def b2_extract_slice(self, ctx, fields10, m, Nm, Nz, Nr, Nr_out, iz, Sz, slice_, stream):
        if not (0 <= m < Nm and 0 < Nr_out <= Nr and 0 <= iz and iz + 1 < Nz):
            return -3
        ptrs = (ctypes.c_void_p * 10)(*_ptrs(fields10, 10))
        return self.emu.emu_extract_slice(ptrs, m, Nm, Nz, Nr, Nr_out, iz, ctypes.c_double(Sz),
                                          ctypes.c_void_p(_addr(slice_)))


HOST_EXT_FIELD_TEMPLATE = r'''
#include <cmath>
using namespace std;
static void one(long long i_, double *F_, const double *x_, const double *y_, const double *z_, double t_,
                double amplitude, double length_scale, double gamma_b_, double beta_b_) {
    const double c_ = 299792458.0;
    const double F = F_[i_], x = x_[i_], y = y_[i_];
    const double z = gamma_b_ * (z_[i_] + beta_b_ * c_ * t_);
    const double t = gamma_b_ * (t_ + beta_b_ * (1. / c_) * z_[i_]);
    (void)F; (void)x; (void)y; (void)z; (void)t; (void)amplitude; (void)length_scale;
%s
}
extern "C" void apply(long long n, double *F_, const double *x_, const double *y_, const double *z_, double t_,
                      double amplitude, double length_scale, double gamma_b_, double beta_b_) {
    for (long long i = 0; i < n; ++i) one(i, F_, x_, y_, z_, t_, amplitude, length_scale, gamma_b_, beta_b_);
}
'''

class CheckedFake(object):
    """What the product sees: every call is first put through the ctypes conversion rules of the real binding
    (`fbpic_b200._lib._SIGNATURES`: number of arguments, integers where the C prototype has integers, pointers or
    None where it has pointers), then forwarded to the fake.  A call that ctypes would reject fails here too."""

    def __init__(self, fake):
        self._fake = fake

    def __getattr__(self, name):
        from fbpic_b200 import _lib
        target = getattr(self._fake, name)
        argtypes = _lib._SIGNATURES.get(name)
        if argtypes is None:
            return target

        def checked(*args):
            if len(args) != len(argtypes):
                raise TypeError('%s: %d arguments given, the C prototype has %d' % (name, len(args), len(argtypes)))
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    if t in (ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_double, ctypes.c_void_p,
                             ctypes.c_char_p):
                        t.from_param(a)
                    elif a is not None and not isinstance(a, (ctypes.Array, ctypes.Structure)) \
                            and not hasattr(a, '_obj'):
                        t.from_param(a)
                except (TypeError, ctypes.ArgumentError) as exc:
                    raise TypeError('%s: argument %d (%r) is not convertible to %s: %s'
                                    % (name, i, a, t.__name__, exc))
            return target(*args)
        return checked


_KEEP = []      # fakes (and the host blocks they own) stay alive for the whole pytest process


def install_global():
    """Install the fake for the lifetime of the process (multi-rank worker scripts of tests/workers run with
    `--fake-device` under torchrun + gloo in the GPU-less container)."""
    from fbpic_b200 import _lib
    fake = FakeLib()
    _KEEP.append(fake)
    _lib._lib, _lib._ctx = CheckedFake(fake), None
    _lib.call.__dict__.clear()
    return fake


def install(monkeypatch):
    """Swap the fake in for the loaded library (pytest monkeypatch: undone at the end of the test).
    `fbpic_b200._lib.call` caches bound entry points: the cache is emptied here and must be emptied
    again by the caller after the test (see the `fake_device` fixture of tests/test_host_flow.py)."""
    from fbpic_b200 import _lib
    fake = FakeLib()
    _KEEP.append(fake)
    monkeypatch.setattr(_lib, '_lib', CheckedFake(fake))
    monkeypatch.setattr(_lib, '_ctx', None)
    monkeypatch.setattr(_lib, '_PINNED_FREE', {})
    _lib.call.__dict__.clear()
    return fake

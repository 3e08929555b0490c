"""GPU tests of the C-ABI entry points of b2_ext.cu one by one (radial PML push / damping, cross-deposition
correction, antenna helpers, NVRTC external field), on random data against the NumPy statements of the reference
formulas -- the device-side twin of tests/test_hostemu_ext.py (same formulas, same shapes: non-multiples of the
block size, k = 0 lines).  Runs before the whole-step tests of the widening so that a failure is localised."""
import ctypes
import numpy as np
import pytest
from scipy.constants import c

from conftest import assert_close

pytestmark = pytest.mark.gpu


def _cplx(rng, shape):
    return (rng.normal(size=shape) + 1.j * rng.normal(size=shape)).astype(np.complex128)


def _dev(*arrays):
    from fbpic_b200._lib import DeviceArray
    return [DeviceArray.from_numpy(a) for a in arrays]


@pytest.mark.parametrize('comoving', [False, True])
@pytest.mark.parametrize('shape', [(5, 7), (9, 130), (257, 64)])
def test_push_eb_pml(comoving, shape):
    from fbpic_b200 import _lib
    rng = np.random.default_rng(3)
    Nz, Nr = shape
    Ep, Em, Bp, Bm, Ez, Bz = [_cplx(rng, shape) for _ in range(6)]
    C, S_w = rng.normal(size=shape), rng.normal(size=shape) * 1e-9
    T = _cplx(rng, shape)
    kr = rng.normal(size=Nr) * 1e5
    Tn = T if comoving else 1.
    ref = [Tn * C * Ep + c**2 * Tn * S_w * (-1.j * 0.5 * kr[None, :] * Bz),        # numba_methods.py:189-214, 358-383
           Tn * C * Em + c**2 * Tn * S_w * (-1.j * 0.5 * kr[None, :] * Bz),
           Tn * C * Bp - Tn * S_w * (-1.j * 0.5 * kr[None, :] * Ez),
           Tn * C * Bm - Tn * S_w * (-1.j * 0.5 * kr[None, :] * Ez)]
    d = _dev(Ep, Em, Bp, Bm, Ez, Bz, C, S_w, T, kr)
    _lib.call.b2_push_eb_pml(_lib.context().handle, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, d[4].ptr, d[5].ptr,
                             d[6].ptr, d[7].ptr, d[8].ptr if comoving else None, d[9].ptr, Nz, Nr, None)
    for got, want, name in zip(d[:4], ref, ('Ep_pml', 'Em_pml', 'Bp_pml', 'Bm_pml')):
        assert_close(got.get(), want, 1e-14, name)
    assert np.array_equal(d[4].get(), Ez) and np.array_equal(d[5].get(), Bz)


@pytest.mark.parametrize('shape,n_pml', [((6, 9), 4), ((33, 70), 33), ((3, 5), 5), ((300, 96), 32)])
def test_damp_pml(shape, n_pml):
    from fbpic_b200 import _lib
    rng = np.random.default_rng(4)
    Nz, Nr = shape
    arrs = [_cplx(rng, shape) for _ in range(6)]
    damp = np.exp(-4. * 0.7 * (np.arange(n_pml) / n_pml)**2)
    wEt, wEtp, wEz, wBt, wBtp, wBz = [a.copy() for a in arrs]
    dd = damp[None, :]                               # pml_damping.py:66-83
    wEt[:, -n_pml:] -= wEtp[:, -n_pml:]
    wBt[:, -n_pml:] -= wBtp[:, -n_pml:]
    wEtp[:, -n_pml:] *= dd
    wBtp[:, -n_pml:] *= dd
    wEt[:, -n_pml:] += wEtp[:, -n_pml:]
    wBt[:, -n_pml:] += wBtp[:, -n_pml:]
    wBz[:, -n_pml:] *= dd
    wEz[:, -n_pml:] *= dd
    d = _dev(*arrs, damp)
    _lib.call.b2_damp_pml(_lib.context().handle, *[a.ptr for a in d[:6]], d[6].ptr, n_pml, Nz, Nr, None)
    for got, w, name in zip(d[:6], (wEt, wEtp, wEz, wBt, wBtp, wBz), ('Et', 'Et_pml', 'Ez', 'Bt', 'Bt_pml', 'Bz')):
        assert_close(got.get(), w, 1e-15, name)


@pytest.mark.parametrize('comoving', [False, True])
def test_correct_currents_cross(comoving):
    from fbpic_b200 import _lib
    from fbpic_b200._lib import SpectralMode
    rng = np.random.default_rng(5)
    Nz, Nr = 10, 67
    rp, rn, rz, rxy, Jp, Jm, Jz = [_cplx(rng, (Nz, Nr)) for _ in range(7)]
    kz1, kr1 = rng.normal(size=Nz) * 1e5, np.abs(rng.normal(size=Nr)) * 1e5
    kz1[0] = 0.
    kr1[3] = 0.
    Tcc, jcc, Teb = [_cplx(rng, (Nz, Nr)) for _ in range(3)]
    inv_dt = 3.e14
    kz, kr = np.broadcast_to(kz1[:, None], (Nz, Nr)), np.broadcast_to(kr1[None, :], (Nz, Nr))
    if comoving:      # numba_methods.py:243-275
        Dz = 1.j * kz * Jz + 0.5 * Tcc * jcc * (rn - Teb * rxy + rz - Teb * rp)
        Dxy = kr * (Jp - Jm) + 0.5 * Tcc * jcc * (rn + Teb * rxy - rz - Teb * rp)
    else:             # numba_methods.py:88-116
        Dz = 1.j * kz * Jz + 0.5 * inv_dt * (rn - rxy + rz - rp)
        Dxy = kr * (Jp - Jm) + 0.5 * inv_dt * (rn - rz + rxy - rp)
    wJp, wJm, wJz = Jp.copy(), Jm.copy(), Jz.copy()
    nzr, nzz = kr != 0, kz != 0
    wJp[nzr] += -0.5 * Dxy[nzr] / kr[nzr]
    wJm[nzr] += 0.5 * Dxy[nzr] / kr[nzr]
    wJz[nzz] += 1.j * Dz[nzz] / kz[nzz]
    d = dict(zip(('rho_prev', 'rho_next', 'rz', 'rxy', 'Jp', 'Jm', 'Jz', 'kz', 'kr', 'T_cc', 'j_corr_coef', 'T_eb'),
                 _dev(rp, rn, rz, rxy, Jp, Jm, Jz, kz1, kr1, Tcc, jcc, Teb)))
    s = SpectralMode()
    for k in ('rho_prev', 'rho_next', 'Jp', 'Jm', 'Jz', 'kz', 'kr', 'T_cc', 'j_corr_coef', 'T_eb'):
        setattr(s, k, d[k].ptr)
    _lib.call.b2_correct_currents_cross(_lib.context().handle, ctypes.byref(s), d['rz'].ptr, d['rxy'].ptr,
                                        int(comoving), inv_dt, Nz, Nr, None)
    for k, w in (('Jp', wJp), ('Jm', wJm), ('Jz', wJz)):
        assert_close(d[k].get(), w, 1e-14, k)
    assert np.array_equal(d['Jz'].get()[0], wJz[0]) and np.array_equal(d['Jp'].get()[:, 3], wJp[:, 3])


def test_correct_divE():
    from scipy.constants import epsilon_0, mu_0
    from fbpic_b200 import _lib
    from fbpic_b200._lib import SpectralMode
    rng = np.random.default_rng(9)
    Nz, Nr = 11, 70
    Ep, Em, Ez, rho = [_cplx(rng, (Nz, Nr)) for _ in range(4)]
    kz1, kr1 = rng.normal(size=Nz) * 1e5, np.abs(rng.normal(size=Nr)) * 1e5
    kz, kr = kz1[:, None], kr1[None, :]
    inv_k2 = np.ascontiguousarray(1. / (kz**2 + kr**2))
    F = -inv_k2 * (-rho / epsilon_0 + 1.j * kz * Ez + kr * (Ep - Em))        # spectral_grid.py:299-314
    want = dict(Ep=Ep + 0.5 * kr * F, Em=Em - 0.5 * kr * F, Ez=Ez - 1.j * kz * F)
    d = dict(zip(('Ep', 'Em', 'Ez', 'rho_prev', 'kz', 'kr', 'inv_k2'), _dev(Ep, Em, Ez, rho, kz1, kr1, inv_k2)))
    s = SpectralMode()
    for k, v in d.items():
        setattr(s, k, v.ptr)
    s.mu_0, s.epsilon_0 = mu_0, epsilon_0
    _lib.call.b2_correct_divE(_lib.context().handle, ctypes.byref(s), Nz, Nr, None)
    for k, w in want.items():
        assert_close(d[k].get(), w, 1e-14, k)


def test_antenna_helpers():
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray
    rng = np.random.default_rng(6)
    n = 700
    host = [rng.normal(size=n) for _ in range(7)]
    bx, by, ex, ey, vx, vy, vz = host
    d = _dev(*host)
    out = [DeviceArray(n, np.float64) for _ in range(5)]
    for sign in (1., -1.):
        _lib.call.b2_antenna_particles(_lib.context().handle, n, *[a.ptr for a in d], sign, *[a.ptr for a in out], None)
        x, y, ux, uy, uz = [a.get() for a in out]
        assert np.array_equal(x, bx + sign * ex) and np.array_equal(y, by + sign * ey)     # antenna_injection.py:360-361
        assert_close(ux, sign * vx / c, 1e-15, 'ux')
        assert_close(uy, sign * vy / c, 1e-15, 'uy')
        assert_close(uz, vz / c, 1e-15, 'uz')
    _lib.call.b2_axpy(_lib.context().handle, n, 0.37, d[4].ptr, d[1].ptr, None)
    assert np.array_equal(d[1].get(), by + 0.37 * vx)


def test_push_p_after_plane():
    """b2_push_p_after_plane against the oracle's Vay push masked to z > z_plane (push_p_after_plane_numba)."""
    from oracle import oracle as orc
    from scipy.constants import e, m_e
    from fbpic_b200 import _lib
    rng = np.random.default_rng(21)
    n = 900
    z = rng.uniform(-1., 1., n)
    z[:3] = 0.2
    u0 = [rng.normal(size=n) * 3. for _ in range(3)]
    ig0 = 1. / np.sqrt(1. + u0[0]**2 + u0[1]**2 + u0[2]**2)
    E = [rng.normal(size=n) * 1.e11 for _ in range(3)]
    B = [rng.normal(size=n) * 300. for _ in range(3)]
    q, m, dt = -e, m_e, 3.e-16
    want = [a.copy() for a in u0] + [ig0.copy()]
    orc.push_p(*want, *E, *B, q, m, dt)
    keep = z <= 0.2
    for w, a in zip(want, u0 + [ig0]):
        w[keep] = a[keep]
    dz, = _dev(z)
    du = _dev(*u0, ig0)
    df = _dev(*E, *B)
    _lib.call.b2_push_p_after_plane(_lib.context().handle, n, dz.ptr, 0.2, *[a.ptr for a in du],
                                    *[a.ptr for a in df], q, m, dt, None)
    for g, w, a, name in zip(du, want, u0 + [ig0], ('ux', 'uy', 'uz', 'inv_gamma')):
        g = g.get()
        assert np.array_equal(g[keep], a[keep]), name
        assert_close(g, w, 1e-14, name)


@pytest.mark.parametrize('Nm', [1, 3])
def test_extract_slice(Nm):
    """b2_extract_slice (lab-frame diagnostics) against extract_slice_cpu of the reference, bit for bit; a slice
    outside of the grid is refused."""
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray, B200Error
    from test_hostemu_ext import _slice_reference
    rng = np.random.default_rng(31)
    Nz, Nr, Nr_out, iz, Sz = 9, 150, 141, 6, 0.3125
    grids = [[_cplx(rng, (Nz, Nr)) for _ in range(10)] for _ in range(Nm)]
    out = DeviceArray(10 * (2 * Nm - 1) * Nr_out, np.float64)
    for m in range(Nm):
        d = _dev(*grids[m])
        _lib.call.b2_extract_slice(_lib.context().handle, _lib.ptr_array(d), m, Nm, Nz, Nr, Nr_out, iz, Sz, out.ptr, None)
    got = out.get().reshape(10, 2 * Nm - 1, Nr_out)
    assert np.array_equal(got, _slice_reference(grids, Nr_out, iz, Sz))
    with pytest.raises(B200Error):
        _lib.call.b2_extract_slice(_lib.context().handle, _lib.ptr_array(d), 0, Nm, Nz, Nr, Nr_out, Nz - 1, Sz, out.ptr,
                                   None)


def test_select_crossing():
    """b2_select_crossing (lab-frame particle diagnostics): the particles of the NumPy selection of the reference, in
    any order; with a buffer that is too small the returned count is still exact."""
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray
    from test_hostemu_ext import _crossing_case
    n, z, uz, ig, c_light, dt, z_curr, z_prev, want = _crossing_case()
    dz, duz, dig = _dev(z, uz, ig)
    count = DeviceArray(1, np.int64)
    for cap in (n, 7):
        idx = DeviceArray(cap, np.int64)
        found = ctypes.c_int64(-1)
        _lib.call.b2_select_crossing(_lib.context().handle, n, dz.ptr, duz.ptr, dig.ptr, c_light, dt, z_curr, z_prev,
                                     cap, idx.ptr, count.ptr, ctypes.byref(found), None)
        assert found.value == len(want)
        got = idx.get()
        if cap >= len(want):
            assert np.array_equal(np.sort(got[:len(want)]), want)
        else:
            assert np.all(np.isin(got, want))
    # the gather of the selected particles: b2_permute with a short index list
    idx = DeviceArray.from_numpy(want.astype(np.int64))
    out = DeviceArray(2 * len(want), np.float64)
    _lib.call.b2_permute(_lib.context().handle, len(want), idx.ptr, 2, _lib.ptr_array([dz, duz]),
                         _lib.ptr_array([out.ptr, out.ptr + 8 * len(want)]), None)
    got = out.get().reshape(2, -1)
    assert np.array_equal(got[0], z[want]) and np.array_equal(got[1], uz[want])

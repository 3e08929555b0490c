"""CPU tests of the kernel SOURCE of fbpic_b200/csrc/b2_ext_kernels.cuh (radial PML, cross-deposition,
laser-antenna helpers): tests/hostemu compiles the very same __global__ bodies with g++ (a launch becomes
nested loops over blockIdx/threadIdx, same launch geometry as b2_ext.cu) and the results are compared with
the NumPy statements of the reference formulas.  This checks index arithmetic and formulas in the GPU-less
build container; the `-m gpu` tests check the compiled sm_100a kernels through the C ABI."""
import ctypes
import numpy as np
import pytest
from scipy.constants import c

from conftest import assert_close


@pytest.fixture(scope='module')
def emu():
    from fake_device import build_emu
    return build_emu()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _cplx(rng, shape):
    return (rng.normal(size=shape) + 1.j * rng.normal(size=shape)).astype(np.complex128)


@pytest.mark.parametrize('comoving', [False, True])
@pytest.mark.parametrize('shape', [(5, 7), (9, 130)])
def test_emu_push_eb_pml(emu, comoving, shape):
    rng = np.random.default_rng(3)
    Nz, Nr = shape
    Ep, Em, Bp, Bm, Ez, Bz = [_cplx(rng, shape) for _ in range(6)]
    C, S_w = rng.normal(size=shape), rng.normal(size=shape) * 1e-9
    T = _cplx(rng, shape) if comoving else None
    kr = rng.normal(size=Nr) * 1e5
    Tn = T if comoving else 1.
    # fbpic/fields/numba_methods.py:189-214 (standard), :358-383 (comoving)
    ref = [Tn * C * Ep + c**2 * Tn * S_w * (-1.j * 0.5 * kr[None, :] * Bz),
           Tn * C * Em + c**2 * Tn * S_w * (-1.j * 0.5 * kr[None, :] * Bz),
           Tn * C * Bp - Tn * S_w * (-1.j * 0.5 * kr[None, :] * Ez),
           Tn * C * Bm - Tn * S_w * (-1.j * 0.5 * kr[None, :] * Ez)]
    Ez0, Bz0 = Ez.copy(), Bz.copy()
    emu.emu_push_eb_pml(_p(Ep), _p(Em), _p(Bp), _p(Bm), _p(Ez), _p(Bz), _p(C), _p(S_w),
                        _p(T) if comoving else None, _p(kr), Nz, Nr)
    for got, want, name in zip((Ep, Em, Bp, Bm), ref, ('Ep_pml', 'Em_pml', 'Bp_pml', 'Bm_pml')):
        assert_close(got, want, 1e-14, name)
    assert np.array_equal(Ez, Ez0) and np.array_equal(Bz, Bz0)


@pytest.mark.parametrize('shape,n_pml', [((6, 9), 4), ((33, 70), 33), ((3, 5), 5)])
def test_emu_damp_pml(emu, shape, n_pml):
    rng = np.random.default_rng(4)
    Nz, Nr = shape
    Et, Etp, Ez, Bt, Btp, Bz = [_cplx(rng, shape) for _ in range(6)]
    damp = np.exp(-4. * 0.7 * (np.arange(n_pml) / n_pml)**2)
    want = [a.copy() for a in (Et, Etp, Ez, Bt, Btp, Bz)]
    wEt, wEtp, wEz, wBt, wBtp, wBz = want
    d = damp[None, :]                               # pml_damping.py:66-83
    wEt[:, -n_pml:] -= wEtp[:, -n_pml:]
    wBt[:, -n_pml:] -= wBtp[:, -n_pml:]
    wEtp[:, -n_pml:] *= d
    wBtp[:, -n_pml:] *= d
    wEt[:, -n_pml:] += wEtp[:, -n_pml:]
    wBt[:, -n_pml:] += wBtp[:, -n_pml:]
    wBz[:, -n_pml:] *= d
    wEz[:, -n_pml:] *= d
    emu.emu_damp_pml(_p(Et), _p(Etp), _p(Ez), _p(Bt), _p(Btp), _p(Bz), _p(damp), n_pml, Nz, Nr)
    for got, w, name in zip((Et, Etp, Ez, Bt, Btp, Bz), want, ('Et', 'Et_pml', 'Ez', 'Bt', 'Bt_pml', 'Bz')):
        assert np.array_equal(got, w), name


@pytest.mark.parametrize('comoving', [False, True])
def test_emu_correct_currents_cross(emu, comoving):
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
    emu.emu_correct_currents_cross(_p(rp), _p(rn), _p(rz), _p(rxy), _p(Jp), _p(Jm), _p(Jz), _p(kz1), _p(kr1),
                                   _p(Tcc), _p(jcc), _p(Teb), int(comoving), ctypes.c_double(inv_dt), Nz, Nr)
    for got, w, name in zip((Jp, Jm, Jz), (wJp, wJm, wJz), ('Jp', 'Jm', 'Jz')):
        assert_close(got, w, 1e-14, name)
    assert np.array_equal(Jz[0], wJz[0]) and np.array_equal(Jp[:, 3], wJp[:, 3])      # untouched where k == 0


def test_emu_antenna_helpers(emu):
    rng = np.random.default_rng(6)
    n = 700
    bx, by, ex, ey, vx, vy, vz = [rng.normal(size=n) for _ in range(7)]
    for sign in (1., -1.):
        x, y, ux, uy, uz = [np.full(n, np.nan) for _ in range(5)]
        emu.emu_antenna_particles(ctypes.c_longlong(n), _p(bx), _p(by), _p(ex), _p(ey), _p(vx), _p(vy), _p(vz),
                                  ctypes.c_double(sign), _p(x), _p(y), _p(ux), _p(uy), _p(uz))
        # antenna_injection.py:360-361, 373-374, 411-413
        assert np.array_equal(x, bx + sign * ex) and np.array_equal(y, by + sign * ey)
        assert_close(ux, sign * vx / c, 1e-15, 'ux')
        assert_close(uy, sign * vy / c, 1e-15, 'uy')
        assert_close(uz, vz / c, 1e-15, 'uz')
    y0 = by.copy()
    emu.emu_axpy(ctypes.c_longlong(n), ctypes.c_double(0.37), _p(vx), _p(by))
    assert np.array_equal(by, y0 + 0.37 * vx)


def test_emu_correct_divE(emu):
    from scipy.constants import epsilon_0
    rng = np.random.default_rng(9)
    Nz, Nr = 11, 70
    Ep, Em, Ez, rho = [_cplx(rng, (Nz, Nr)) for _ in range(4)]
    kz1, kr1 = rng.normal(size=Nz) * 1e5, np.abs(rng.normal(size=Nr)) * 1e5
    kz, kr = kz1[:, None], kr1[None, :]
    inv_k2 = 1. / (kz**2 + kr**2)
    F = -inv_k2 * (-rho / epsilon_0 + 1.j * kz * Ez + kr * (Ep - Em))        # spectral_grid.py:299-314
    want = [Ep + 0.5 * kr * F, Em - 0.5 * kr * F, Ez - 1.j * kz * F]
    emu.emu_correct_divE(_p(Ep), _p(Em), _p(Ez), _p(rho), _p(kz1), _p(kr1), _p(np.ascontiguousarray(inv_k2)),
                         ctypes.c_double(1. / epsilon_0), Nz, Nr)
    for got, w, name in zip((Ep, Em, Ez), want, ('Ep', 'Em', 'Ez')):
        assert_close(got, w, 1e-14, name)


def test_emu_push_p_after_plane(emu):
    """push/numba_methods.py push_p_after_plane_numba: the Vay push where z > z_plane, untouched momenta elsewhere."""
    from oracle import oracle as orc
    from scipy.constants import e, m_e
    rng = np.random.default_rng(21)
    n = 900
    z = rng.uniform(-1., 1., n)
    z[:3] = 0.2                                                  # exactly on the plane: not pushed (strict >)
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
    got = [a.copy() for a in u0] + [ig0.copy()]
    emu.emu_push_p_after_plane(ctypes.c_longlong(n), _p(z), ctypes.c_double(0.2), *[_p(a) for a in got],
                               *[_p(a) for a in E + B], ctypes.c_double(q * dt / (m * c)),
                               ctypes.c_double(0.5 * q * dt / m))
    for g, w, a, name in zip(got, want, u0 + [ig0], ('ux', 'uy', 'uz', 'inv_gamma')):
        assert np.array_equal(g[keep], a[keep]), name
        assert_close(g, w, 1e-14, name)
    assert 0 < keep.sum() < n


def _slice_reference(grids, Nr_out, iz, Sz):
    """extract_slice_cpu + get_dataset (boosted_field_diag.py:604-684): Sz * F[iz] + (1 - Sz) * F[iz + 1], modes m > 0
    doubled, real and imaginary parts as separate rows"""
    Nm = len(grids)
    out = np.empty((10, 2 * Nm - 1, Nr_out))
    for k in range(10):
        for m in range(Nm):
            a = grids[m][k] * (2. if m else 1.)
            row = Sz * a[iz, :Nr_out] + (1. - Sz) * a[iz + 1, :Nr_out]
            if m == 0:
                out[k, 0] = row.real
            else:
                out[k, 2 * m - 1], out[k, 2 * m] = row.real, row.imag
    return out


@pytest.mark.parametrize('Nm', [1, 3])
def test_emu_extract_slice(emu, Nm):
    rng = np.random.default_rng(31)
    Nz, Nr, Nr_out, iz, Sz = 9, 150, 141, 6, 0.3125
    grids = [[_cplx(rng, (Nz, Nr)) for _ in range(10)] for _ in range(Nm)]
    got = np.full((10, 2 * Nm - 1, Nr_out), np.nan)
    for m in range(Nm):
        ptrs = (ctypes.c_void_p * 10)(*[g.ctypes.data for g in grids[m]])
        emu.emu_extract_slice(ptrs, m, Nm, Nz, Nr, Nr_out, iz, ctypes.c_double(Sz), _p(got))
    want = _slice_reference(grids, Nr_out, iz, Sz)
    assert np.array_equal(got, want)


def _crossing_case():
    rng = np.random.default_rng(41)
    n = 3000
    z = rng.uniform(0., 1., n)
    uz = rng.normal(size=n) * 2.
    ig = 1. / np.sqrt(1. + uz**2 + rng.uniform(0., 1., n))
    c_light, dt, z_curr, z_prev = 3., 0.01, 0.48, 0.52
    z_old = z - uz * ig * c_light * dt
    want = np.flatnonzero(((z >= z_curr) & (z_old <= z_prev)) | ((z <= z_curr) & (z_old >= z_prev)))
    return n, z, uz, ig, c_light, dt, z_curr, z_prev, want


def test_emu_select_crossing(emu):
    """get_particle_slice (boosted_particle_diag.py:598-629): same particles as the NumPy selection; the count stays
    exact when the index buffer is too small."""
    n, z, uz, ig, c_light, dt, z_curr, z_prev, want = _crossing_case()
    assert 20 < len(want) < n // 4
    D = ctypes.c_double
    for cap in (n, 7):
        idx = np.full(max(cap, 1), -1, dtype=np.int64)
        count = np.zeros(1, dtype=np.uint64)
        emu.emu_select_crossing(ctypes.c_longlong(n), _p(z), _p(uz), _p(ig), D(c_light), D(dt), D(z_curr), D(z_prev),
                                ctypes.c_longlong(cap), _p(idx), _p(count))
        assert int(count[0]) == len(want)
        if cap >= len(want):
            assert np.array_equal(np.sort(idx[:len(want)]), want)
        else:
            assert np.all(np.isin(idx[:cap], want))


def _ionize_case():
    """inputs and probabilities of the reference's ADK functions for nitrogen (tests/golden/ionization.npz)"""
    from conftest import load_golden
    from fbpic_b200.ionization import Ionizer
    import types
    g = load_golden('ionization')
    ion = types.SimpleNamespace(level_max=None)
    Ionizer.initialize_ADK_parameters(ion, 'N', float(g['dt']))
    tables = [np.ascontiguousarray(a) for a in (ion.adk_prefactor, ion.adk_power, ion.adk_exp_prefactor)]
    return g, tables


def test_emu_ionize(emu):
    """k_ionize against get_E_amplitude / get_ionization_probability of the reference (inline_functions.py:9-45):
    with given draws, exactly the ions with draw < p (and level < level_max) move up one level and are reported with
    their former level; with the built-in generator the ionized fraction matches the mean probability."""
    g, tables = _ionize_case()
    n, level_max = len(g['level']), 6
    rng = np.random.default_rng(7)
    draws = rng.uniform(size=n)
    p = g['probability']
    clear = np.abs(draws - p) > 1e-9                      # away from the rounding of p
    assert clear.all()
    want = np.flatnonzero((draws < p) & (g['level'] < level_max))
    assert 30 < len(want) < n - 30
    arrs = [np.ascontiguousarray(a) for a in (*g['u'], *g['E'], *g['B'])]
    D = ctypes.c_double
    level = g['level'].astype(np.uint64)
    events = np.full(2 * n, -1, dtype=np.int64)
    count = np.zeros(1, dtype=np.uint64)
    emu.emu_ionize(ctypes.c_longlong(n), _p(level), level_max, *[_p(t) for t in tables], *[_p(a) for a in arrs],
                   D(c), _p(draws), ctypes.c_ulonglong(0), ctypes.c_longlong(n), _p(events), _p(count))
    k = int(count[0])
    ev = events[:2 * k].reshape(k, 2)
    ev = ev[np.argsort(ev[:, 0])]
    assert np.array_equal(ev[:, 0], want) and np.array_equal(ev[:, 1], g['level'][want])
    expect = g['level'].copy()
    expect[want] += 1
    assert np.array_equal(level, expect.astype(np.uint64))
    # built-in counter-based generator: deterministic for a seed, different between seeds, right on average
    big = 200
    tiled = [np.tile(a, big) for a in arrs]
    fractions = []
    for seed in (11, 11, 12):
        level = np.tile(g['level'], big).astype(np.uint64)
        events = np.zeros(2 * n * big, dtype=np.int64)
        emu.emu_ionize(ctypes.c_longlong(n * big), _p(level), level_max, *[_p(t) for t in tables], *[_p(a) for a in tiled],
                       D(c), None, ctypes.c_ulonglong(seed), ctypes.c_longlong(n * big), _p(events), _p(count))
        fractions.append(int(count[0]))
    mean = np.where(g['level'] < level_max, p, 0.).sum() * big
    sigma = np.sqrt(np.where(g['level'] < level_max, p * (1 - p), 0.).sum() * big)
    assert fractions[0] == fractions[1] != fractions[2]
    assert abs(fractions[0] - mean) < 5 * sigma and abs(fractions[2] - mean) < 5 * sigma


def test_emu_push_p_ioniz_and_weights(emu):
    """k_push_p_ioniz: Vay push with the charge level * e, neutral particles untouched (push_p_ioniz_numba);
    k_w_times_level."""
    from oracle import oracle as orc
    from scipy.constants import e, m_p
    rng = np.random.default_rng(23)
    n = 500
    level = rng.integers(0, 4, n).astype(np.uint64)
    u0 = [rng.normal(size=n) for _ in range(3)]
    ig0 = 1. / np.sqrt(1. + u0[0]**2 + u0[1]**2 + u0[2]**2)
    E = [rng.normal(size=n) * 1.e12 for _ in range(3)]
    B = [rng.normal(size=n) * 3000. for _ in range(3)]
    m, dt = 14. * m_p, 2.e-16
    want = [a.copy() for a in u0] + [ig0.copy()]
    for lv in range(1, 4):
        sel = level == lv
        part = [a[sel].copy() for a in want]
        orc.push_p(*part, *[a[sel].copy() for a in E + B], lv * e, m, dt)
        for w_, p_ in zip(want, part):
            w_[sel] = p_
    got = [a.copy() for a in u0] + [ig0.copy()]
    emu.emu_push_p_ioniz(ctypes.c_longlong(n), _p(level), *[_p(a) for a in got], *[_p(a) for a in E + B],
                         ctypes.c_double(e * dt / (m * c)), ctypes.c_double(0.5 * e * dt / m))
    for g_, w_, a, name in zip(got, want, u0 + [ig0], ('ux', 'uy', 'uz', 'inv_gamma')):
        assert np.array_equal(g_[level == 0], a[level == 0]), name
        assert_close(g_, w_, 1e-14, name)
    w = rng.uniform(1., 2., n)
    out = np.zeros(n)
    emu.emu_w_times_level(ctypes.c_longlong(n), _p(w), _p(level), _p(out))
    assert np.array_equal(out, w * level)


def test_emu_compton_count_and_scatter(emu):
    """k_compton_count against the NumPy statement of get_photon_density_gaussian / get_scattering_probability
    (compton/inline_functions.py:39-112): the number of photons per electron is p x ratio on average and reproducible
    for a seed; k_compton_scatter fills exactly that many photon slots, each with |p| = 1 / inv_gamma, at the position
    of an emitting electron, with a weight w / ratio."""
    from scipy.constants import h, m_e, physical_constants
    r_e = physical_constants['classical electron radius'][0]
    rng = np.random.default_rng(61)
    n = 40000
    x, y = rng.normal(size=n) * 5.e-6, rng.normal(size=n) * 5.e-6
    z = rng.normal(size=n) * 2.e-6
    ux, uy = rng.normal(size=n) * 0.1, rng.normal(size=n) * 0.1
    uz = 30. + rng.normal(size=n)
    ig = 1. / np.sqrt(1 + ux**2 + uy**2 + uz**2)
    w = rng.uniform(1., 2., n)
    lam, waist, ctau, z0, ratio, dt = h * c / 1.602176634e-19, 30.e-6, 6.e-4, 0., 40., 4.e-13
    p_ph = h / lam
    # peak photon density such that an electron at the centre emits about one photon macroparticle in 40 cycles
    n_peak = (1. / 40. / ratio) / (8. / 3 * np.pi * r_e**2 * 2. * c * dt)
    ct = 1.e-6
    P = np.array([ct, n_peak, 1. / waist**2, 1. / ctau**2, z0, 1., 0., p_ph, 0., 0., -p_ph, 0., 0., -1., dt, ratio,
                  1. / ratio, np.pi * r_e**2, 1. / (m_e * c), c])
    # NumPy statement of the reference formulas (lab frame: gamma_boost = 1)
    n_ph = n_peak * np.exp(-2 * (x**2 + y**2) / waist**2 - 2 * (z - z0 + ct)**2 / ctau**2)
    tf = 1. / ig + uz
    k = p_ph * tf / (m_e * c)
    sigma = np.pi * r_e**2 * (2 * (2 + k * (1 + k) * (8 + k)) / (k**2 * (1 + 2 * k)**2)
                              - (2 + k * (2 - k)) * np.log(1 + 2 * k) / k**3)
    prob = 1 - np.exp(-sigma * n_ph * tf * c * dt * ig)
    assert 0.005 < prob.mean() * ratio < 5.
    ns = [np.zeros(n, dtype=np.int32) for _ in range(3)]
    total = np.zeros(1, dtype=np.uint64)
    totals = []
    for seed, out in zip((5, 5, 6), ns):
        emu.emu_compton_count(ctypes.c_longlong(n), *[_p(a) for a in (x, y, z, ux, uy, uz, ig)], _p(P),
                              ctypes.c_ulonglong(seed), _p(out), _p(total))
        totals.append(int(total[0]))
        assert totals[-1] == out.sum()
    assert np.array_equal(ns[0], ns[1]) and not np.array_equal(ns[0], ns[2])
    mean, sig = (prob * ratio).sum(), np.sqrt(n / 12. + 1.)        # int(p r + u): uniform rounding noise
    assert abs(totals[0] - mean) < 6 * sig and abs(totals[2] - mean) < 6 * sig
    # the photons
    N = totals[0]
    ph = [np.full(N, np.nan) for _ in range(8)]
    ptrs = (ctypes.c_void_p * 8)(*[a.ctypes.data for a in ph])
    e_u = [ux.copy(), uy.copy(), uz.copy()]
    cursor = np.zeros(1, dtype=np.uint64)
    emu.emu_compton_scatter(ctypes.c_longlong(n), _p(ns[0]), _p(x), _p(y), _p(z), *[_p(a) for a in e_u], _p(ig), _p(w),
                            _p(P), ctypes.c_ulonglong(5), ptrs, _p(cursor))
    assert int(cursor[0]) == N and not np.isnan(np.stack(ph)).any()
    px, py, pz, inv_p = ph[3], ph[4], ph[5], ph[6]
    assert np.allclose(np.sqrt(px**2 + py**2 + pz**2) * inv_p, 1., rtol=1e-12)
    assert np.all(np.isin(ph[0], x[ns[0] > 0])) and np.all(np.isin(np.round(ph[7] * ratio, 9), np.round(w, 9)))
    # up-shifted by up to 4 gamma^2 and beamed forward; electrons that recoiled lost longitudinal momentum
    assert np.mean(pz > 0) > 0.99 and (1. / inv_p).max() < 4 * 35.**2 * p_ph
    changed = e_u[2] != uz
    assert 0 < changed.sum() <= (ns[0] > 0).sum() and np.all(e_u[2][changed] < uz[changed])

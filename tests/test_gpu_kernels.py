"""GPU parity tests (kernel level): the CUDA path, called through the operator surface /
C ABI, against (a) the golden vectors generated from the unmodified FBPIC reference and
(b) the oracle on larger seeded inputs.  Tolerances: indices bit-exact; fp64 results within
1e-13*(max|a|+max|b|) for deposition (the reference's own CPU/GPU tolerance,
tests/test_cpu_gpu_deposition.py:96-98), 1e-13 for gather / push."""
import ctypes
import os
import numpy as np
import pytest
from scipy.constants import c

from conftest import load_golden, assert_close
from fbpic_b200 import _lib
from fbpic_b200 import host_tables as ht

pytestmark = pytest.mark.gpu


def make_sim(g, shape, Nm):
    from fbpic_b200 import Simulation
    Nz, Nr = int(g['Nz']), int(g['Nr'])
    sim = Simulation(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']),
                     p_zmin=float(g['zmin']), p_zmax=float(g['zmax']), p_rmin=0, p_rmax=float(g['rmax']),
                     p_nz=1, p_nr=1, p_nt=4, n_e=1.e24, zmin=float(g['zmin']), particle_shape=shape)
    return sim


def set_particles(sp, P):
    n = len(P['x'])
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        setattr(sp, k, np.array(P[k], dtype=np.float64))
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(n))
    sp.Ntot = n
    sp.data_is_on_gpu = False
    sp.sorted = False


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
@pytest.mark.parametrize('Nm', [1, 2, 3])
def test_kernels_vs_reference_golden(shape, Nm):
    from oracle import oracle as orc
    g = load_golden('kernels_%s_Nm%d' % (shape, Nm))
    sim = make_sim(g, shape, Nm)
    sp = sim.ptcl[0]
    P = {k: g['p_' + k] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')}
    set_particles(sp, P)
    sp.q, sp.m = float(g['q']), float(g['m'])
    Nz, Nr = int(g['Nz']), int(g['Nr'])
    g0 = sim.fld.interp[0]
    # ---- sort: bit-exact keys, stable order, prefix sum
    sim.send_data_to_gpu()
    sp.sort_particles(sim.fld)
    cell = orc.cell_index(P['x'], P['y'], P['z'], g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr)
    idx, prefix = orc.sort_contract(cell, Nz, Nr)
    assert np.array_equal(sp.sorted_idx.get(), idx)
    assert np.array_equal(sp.prefix_sum.get(), prefix)
    assert np.array_equal(sp.cell_idx.get(), cell[idx])
    for k in P:
        assert np.array_equal(getattr(sp, k).get(), P[k][idx]), k
    sp.sorted = True
    # ---- deposit rho, J
    for ft, names in (('rho', ('rho',)), ('J', ('Jr', 'Jt', 'Jz'))):
        sim.fld.erase(ft)
        sp.deposit(sim.fld, ft)
        sim.fld.divide_by_volume(ft)
        for m in range(Nm):
            for nme in names:
                assert_close(getattr(sim.fld.interp[m], nme).get(), g['dep_%s_m%d' % (nme, m)], 1e-13,
                             'deposit %s m%d' % (nme, m))
    # ---- gather (original particle order)
    sim.receive_data_from_gpu()
    set_particles(sp, P)
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            setattr(sim.fld.interp[m], k, g['grid_%s_m%d' % (k, m)].copy())
    sim.send_data_to_gpu()
    sp.gather(sim.fld.interp, sim.comm)
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        assert_close(getattr(sp, k).get(), g['gath_' + k], 1e-13, 'gather ' + k)
    # ---- push_p + push_x, separate kernels
    sp.push_p(0.)
    sp.push_x(0.5 * sim.dt)
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma'):
        assert_close(getattr(sp, k).get(), g['push_' + k], 1e-14, 'push ' + k)
    # ---- fused gather+push gives the same
    sim.receive_data_from_gpu()
    set_particles(sp, P)
    sim.send_data_to_gpu()
    sp.gather_and_push(sim.fld.interp, sim.comm, 0.5 * sim.dt)
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma'):
        assert_close(getattr(sp, k).get(), g['push_' + k], 1e-13, 'fused push ' + k)


@pytest.mark.parametrize('shape,Nm', [('linear', 2), ('cubic', 2), ('linear', 4)])
def test_deposit_gather_vs_oracle_large(shape, Nm):
    """Seeded plasma of ~0.4 M particles on a 96x40 grid: CUDA vs oracle."""
    from fbpic_b200 import Simulation
    from oracle import oracle as orc
    np.random.seed(3)
    Nz, Nr, zmax, rmax = 96, 40, 30.e-6, 20.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4 * Nm, n_e=1.e24, particle_shape=shape)
    sp = sim.ptcl[0]
    n = sp.Ntot
    rng = np.random.default_rng(5)
    sp.ux, sp.uy, sp.uz = rng.normal(size=n) * 0.3, rng.normal(size=n) * 0.3, rng.normal(size=n)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    # perturb positions so that particles are unsorted and cross the axis / the outer edge
    sp.x += rng.normal(size=n) * 0.3e-6
    sp.y += rng.normal(size=n) * 0.3e-6
    sp.z = np.mod(sp.z + rng.normal(size=n) * 0.4e-6, zmax)
    P = {k: getattr(sp, k).copy() for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')}
    g0 = sim.fld.interp[0]
    cubic = shape == 'cubic'
    coef = 'ruyten_cubic_coef' if cubic else 'ruyten_linear_coef'
    sim.send_data_to_gpu()
    for ft, names in (('rho', ('rho',)), ('J', ('Jr', 'Jt', 'Jz'))):
        sim.fld.erase(ft)
        sp.deposit(sim.fld, ft)
        raw = orc.deposit(ft, P['x'], P['y'], P['z'], P['w'], sp.q, P['ux'], P['uy'], P['uz'], P['inv_gamma'],
                          g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr, Nm, cubic,
                          getattr(sim.fld.interp[0], coef), getattr(sim.fld.interp[1], coef))
        for m in range(Nm):
            for k, nme in enumerate(names):
                assert_close(getattr(sim.fld.interp[m], nme).get(), raw[k, m], 1e-13, '%s m%d' % (nme, m))
    cell = orc.cell_index(P['x'], P['y'], P['z'], g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr)
    idx, prefix = orc.sort_contract(cell, Nz, Nr)
    assert np.array_equal(sp.sorted_idx.get(), idx)
    assert np.array_equal(sp.prefix_sum.get(), prefix)
    # gather from smooth random fields
    sim.receive_data_from_gpu()
    grids = []
    for m in range(Nm):
        gm = []
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            a = (rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))) * (1e9 if k[0] == 'E' else 3.)
            setattr(sim.fld.interp[m], k, a)
            gm.append(a)
        grids.append(tuple(gm))
    sim.send_data_to_gpu()
    sp.gather(sim.fld.interp, sim.comm)
    Ps = {k: getattr(sp, k).get() for k in ('x', 'y', 'z')}
    F = {k: np.zeros(n) for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz')}
    orc.gather(Ps['x'], Ps['y'], Ps['z'], rmax, g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr, grids, cubic,
               F['Ex'], F['Ey'], F['Ez'], F['Bx'], F['By'], F['Bz'])
    for k in F:
        assert_close(getattr(sp, k).get(), F[k], 1e-13, 'gather ' + k)
    # fused gather + push on the SORTED particles (the tiled kernels, linear and cubic: field tiles in shared memory)
    # against the oracle's push with the oracle's gathered fields
    Pu = {k: getattr(sp, k).get() for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma')}
    sp.gather_and_push(sim.fld.interp, sim.comm, 0.5 * dt)
    orc.push_p(Pu['ux'], Pu['uy'], Pu['uz'], Pu['inv_gamma'], F['Ex'], F['Ey'], F['Ez'], F['Bx'], F['By'], F['Bz'],
               sp.q, sp.m, dt)
    orc.push_x(Pu['x'], Pu['y'], Pu['z'], Pu['ux'], Pu['uy'], Pu['uz'], Pu['inv_gamma'], 0.5 * dt)
    for k in ('ux', 'uy', 'uz', 'inv_gamma', 'x', 'y', 'z'):
        assert_close(getattr(sp, k).get(), Pu[k], 1e-13, 'fused gather+push ' + k)


@pytest.mark.parametrize('Nz,Nr', [(64, 48), (200, 64), (37, 50), (128, 256)])
def test_transforms_vs_numpy(Nz, Nr):
    """FFT + Hankel GEMM (DMMA) round trips and values against NumPy on random data."""
    from fbpic_b200.fields import SpectralTransformer
    from fbpic_b200._lib import DeviceArray
    rng = np.random.default_rng(Nz + Nr)
    rmax = 30.e-6
    for m in (0, 1, 2):
        tr = SpectralTransformer(Nz, Nr, m, rmax)
        f = rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
        h = rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
        d_f, d_h = DeviceArray.from_numpy(f), DeviceArray.from_numpy(h)
        d_s, d_p, d_m = [DeviceArray((Nz, Nr), np.complex128) for _ in range(3)]
        tr.interp2spect_scal(d_f, d_s)
        ref = np.fft.fft(f, axis=0) @ tr.dht0.M
        assert_close(d_s.get(), ref, 1e-13, 'fwd scal m%d' % m)
        tr.interp2spect_vect(d_f, d_h, d_p, d_m)
        fr, ft = np.fft.fft(f, axis=0), np.fft.fft(h, axis=0)
        assert_close(d_p.get(), (0.5 * (fr - 1.j * ft)) @ tr.dhtp.M, 1e-13, 'fwd p m%d' % m)
        assert_close(d_m.get(), (0.5 * (fr + 1.j * ft)) @ tr.dhtm.M, 1e-13, 'fwd m m%d' % m)
        d_r, d_t = DeviceArray((Nz, Nr), np.complex128), DeviceArray((Nz, Nr), np.complex128)
        tr.spect2interp_vect(d_f, d_h, d_r, d_t)
        pp, mm = f @ tr.dhtp.invM, h @ tr.dhtm.invM
        assert_close(d_r.get(), np.fft.ifft(pp + mm, axis=0), 1e-13, 'inv r m%d' % m)
        assert_close(d_t.get(), np.fft.ifft(1.j * (pp - mm), axis=0), 1e-13, 'inv t m%d' % m)
        tr.spect2interp_scal(d_s, d_r)
        assert_close(d_r.get(), np.fft.ifft(d_s.get() @ tr.dht0.invM, axis=0), 1e-13, 'inv scal m%d' % m)
        if m == 0:      # the order-0 matrices of mode 0 are exact inverses of each other
            assert_close(d_r.get(), f, 1e-11, 'round trip m%d' % m)


@pytest.mark.parametrize('shape,Nm', [('linear', 2), ('linear', 3), ('cubic', 2)])
def test_fused_deposition_paths_vs_oracle(shape, Nm):
    """deposit_fused(): sort + (permutation fused into the deposition kernel) for J, and the
    displaced rho deposition on particles that moved since the sort (incl. a few that moved by
    several cells and take the per-particle fallback), all against the oracle."""
    from fbpic_b200 import Simulation
    from oracle import oracle as orc
    np.random.seed(4)
    Nz, Nr, zmax, rmax = 80, 36, 24.e-6, 18.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4 * Nm, n_e=1.e24, particle_shape=shape)
    sp = sim.ptcl[0]
    n = sp.Ntot
    rng = np.random.default_rng(9)
    sp.ux, sp.uy, sp.uz = rng.normal(size=n) * 0.3, rng.normal(size=n) * 0.3, rng.normal(size=n)
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    sp.x += rng.normal(size=n) * 0.4e-6
    sp.y += rng.normal(size=n) * 0.4e-6
    sp.z = np.mod(sp.z + rng.normal(size=n) * 0.5e-6, zmax)
    g0 = sim.fld.interp[0]
    cubic = shape == 'cubic'
    coef = 'ruyten_cubic_coef' if cubic else 'ruyten_linear_coef'
    r0, rh = getattr(sim.fld.interp[0], coef), getattr(sim.fld.interp[1], coef)
    P = {k: getattr(sp, k).copy() for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')}

    def oracle_dep(what, Q):
        return orc.deposit(what, Q['x'], Q['y'], Q['z'], Q['w'], sp.q, Q['ux'], Q['uy'], Q['uz'], Q['inv_gamma'],
                           g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr, Nm, cubic, r0, rh)

    sim.send_data_to_gpu()
    # J: sort + fused permute/deposit
    sim.fld.erase('J')
    sp.deposit_fused(sim.fld, 'J')
    raw = oracle_dep('J', P)
    for m in range(Nm):
        for k, nme in enumerate(('Jr', 'Jt', 'Jz')):
            assert_close(getattr(sim.fld.interp[m], nme).get(), raw[k, m], 1e-13, 'fused %s m%d' % (nme, m))
    cell = orc.cell_index(P['x'], P['y'], P['z'], g0.invdz, g0.zmin, Nz, g0.invdr, 0., Nr)
    idx, prefix = orc.sort_contract(cell, Nz, Nr)
    assert np.array_equal(sp.prefix_sum.get(), prefix)
    for k in P:
        assert np.array_equal(getattr(sp, k).get(), P[k][idx]), k     # SoA permuted by the same kernel
    # move the particles (most by < 1 cell, a few by several cells), keep the array order
    Q = {k: P[k][idx].copy() for k in P}
    dz, dr = zmax / Nz, rmax / Nr
    Q['x'] += rng.uniform(-0.45, 0.45, n) * dr
    Q['y'] += rng.uniform(-0.45, 0.45, n) * dr
    Q['z'] += rng.uniform(-0.9, 0.9, n) * dz
    far = rng.choice(n, 200, replace=False)
    Q['z'][far] += rng.uniform(-6, 6, 200) * dz
    Q['x'][far] += rng.uniform(-4, 4, 200) * dr
    Q['z'] = np.mod(Q['z'], zmax)
    for k in ('x', 'y', 'z'):
        getattr(sp, k).set(Q[k])
    sp.sorted = False
    sim.fld.erase('rho')
    sp.deposit_fused(sim.fld, 'rho')
    raw = oracle_dep('rho', Q)
    for m in range(Nm):
        assert_close(sim.fld.interp[m].rho.get(), raw[0, m], 1e-13, 'displaced/fused rho m%d' % m)


@pytest.mark.parametrize('Nz,Nr', [(1000, 200), (2304, 72), (90, 16)])
def test_dht_batch_all_kinds_vs_numpy(Nz, Nr):
    """b2_dht_batch with a mixed job list (scalar, (r,t)->(p,m), (p,m)->(r,t), with and without a row scale) at
    sizes whose 16-row units do not divide evenly among the persistent CTAs of the TMA-fed kernel (partial
    tiles, several column strips, zero-padded K and N of the packed matrices); out = rowscale * (A @ M)."""
    import ctypes
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray, DhtJob, call
    rng = np.random.default_rng(Nz * 7 + Nr)
    cplx = lambda: rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
    mats = [rng.normal(size=(Nr, Nr)) for _ in range(4)]
    d_mats = [DeviceArray.from_numpy(M) for M in mats]
    rs = rng.uniform(0.5, 1.5, size=Nz)
    d_rs = DeviceArray.from_numpy(rs)
    ins = [cplx() for _ in range(8)]
    d_ins = [DeviceArray.from_numpy(a) for a in ins]
    d_outs = [DeviceArray((Nz, Nr), np.complex128) for _ in range(10)]
    jobs, want = [], []
    # 3 scalar jobs (one with row scale)
    for k, (mi, scale) in enumerate([(0, None), (1, d_rs), (2, None)]):
        jobs.append(DhtJob(d_ins[k].ptr, None, d_outs[k].ptr, None, d_mats[mi].ptr, None,
                           scale.ptr if scale is not None else None, _lib.DHT_SCALAR))
        want.append((k, (ins[k] @ mats[mi]) * (rs[:, None] if scale is not None else 1.)))
    # forward vector job with row scale: p = (r - i t)/2 @ M0, m = (r + i t)/2 @ M3
    jobs.append(DhtJob(d_ins[3].ptr, d_ins[4].ptr, d_outs[3].ptr, d_outs[4].ptr, d_mats[0].ptr, d_mats[3].ptr,
                       d_rs.ptr, _lib.DHT_RT_TO_PM))
    want.append((3, (0.5 * (ins[3] - 1.j * ins[4]) @ mats[0]) * rs[:, None]))
    want.append((4, (0.5 * (ins[3] + 1.j * ins[4]) @ mats[3]) * rs[:, None]))
    # two inverse vector jobs: r = P + Q, t = i (P - Q)
    for a, b, o, (m1, m2) in ((5, 6, 5, (1, 2)), (6, 7, 7, (3, 0))):
        jobs.append(DhtJob(d_ins[a].ptr, d_ins[b].ptr, d_outs[o].ptr, d_outs[o + 1].ptr, d_mats[m1].ptr,
                           d_mats[m2].ptr, None, _lib.DHT_PM_TO_RT))
        P, Q = ins[a] @ mats[m1], ins[b] @ mats[m2]
        want.append((o, P + Q))
        want.append((o + 1, 1.j * (P - Q)))
    arr = (DhtJob * len(jobs))(*jobs)
    call.b2_dht_batch(_lib.context().handle, len(jobs), arr, Nz, Nr, None)
    for o, ref in want:
        assert_close(d_outs[o].get(), ref, 1e-13, 'output %d' % o)
    # a matrix buffer that is overwritten must not be served from the packed-matrix cache
    M_new = rng.normal(size=(Nr, Nr))
    d_mats[0].set(M_new)
    call.b2_dht(_lib.context().handle, d_ins[0].ptr, d_outs[9].ptr, d_mats[0].ptr, None, Nz, Nr, None)
    assert_close(d_outs[9].get(), ins[0] @ M_new, 1e-13, 'after matrix update')


def test_gather_push_pipe_variant_matches_golden():
    """The persistent TMA-staged gather+push kernel (opt-in, B2_GATHER_IMPL=pipe; it hands the last n % 128 particles
    to the tiled kernel) against the same goldens and oracle cases as the default: the switch is read once per
    process, hence the subprocess."""
    import subprocess
    import sys
    env = dict(os.environ, B2_GATHER_IMPL='pipe')
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-m', 'gpu', '-q', '-x',
                          '-p', 'no:cacheprovider', '-k', 'golden and not variant or gather_vs_oracle'],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.parametrize('Nz,Nr,na', [(4224, 256, 13), (2240, 512, 3), (4096, 300, 2), (256, 16, 2), (4416, 130, 2)])
def test_fft_many_arrays_vs_numpy(Nz, Nr, na):
    """The batched z-FFT call against numpy.fft at the lengths of the bench configurations (4224 = z-slab of C3 with
    its guards, 2240 = C5, 4416 = C2 with damping cells, radix 23): more arrays than one launch group carries, a
    ragged last column tile (Nr = 300, 130), back-to-back calls on the same scratch, in-place round trip."""
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray, call, ptr_array
    rng = np.random.default_rng(Nz + 7 * Nr)
    ctx = _lib.context()
    arrs = [rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr)) for _ in range(na)]
    d_in = [DeviceArray.from_numpy(a) for a in arrs]
    d_out = [DeviceArray((Nz, Nr), np.complex128) for _ in range(na)]
    for rep in range(2):
        for inverse, ref in ((0, lambda a: np.fft.fft(a, axis=0)), (2, lambda a: np.fft.ifft(a, axis=0) * Nz)):
            call.b2_fft_z_multi(ctx.handle, na, ptr_array(d_in), ptr_array(d_out), Nz, Nr, inverse, None)
            for k in range(na):
                assert_close(d_out[k].get(), ref(arrs[k]), 1e-13, 'rep %d array %d, inverse=%d' % (rep, k, inverse))
    call.b2_fft_z_multi(ctx.handle, na, ptr_array(d_in), ptr_array(d_in), Nz, Nr, 0, None)
    call.b2_fft_z_multi(ctx.handle, na, ptr_array(d_in), ptr_array(d_in), Nz, Nr, 1, None)
    for k in range(na):
        assert_close(d_in[k].get(), arrs[k], 1e-13, 'in place round trip, array %d' % k)


@pytest.mark.parametrize('Nz,Nr', [(4096, 256), (4224, 64), (4320, 32), (4416, 20), (2048, 50), (200, 64),
                                   (1000, 33), (64, 48), (37, 50)])
def test_fft_z_vs_numpy(Nz, Nr):
    """z-FFT (two-pass transform of b2_fft.cu for the lengths it has a plan for -- powers of two and the
    guard-cell lengths 4224 = 64*66, 4320 = 60*72, 4416 = 64*69 -- cuFFT otherwise) against numpy.fft: forward,
    inverse scaled by 1/Nz, inverse unscaled; single call and a multi-array call with an odd number of arrays."""
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray, call, ptr_array
    rng = np.random.default_rng(Nz + 3 * Nr)
    ctx = _lib.context()
    arrs = [rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr)) for _ in range(5)]
    d_in = [DeviceArray.from_numpy(a) for a in arrs]
    d_out = [DeviceArray((Nz, Nr), np.complex128) for _ in range(5)]
    for inverse, ref in ((0, lambda a: np.fft.fft(a, axis=0)), (1, lambda a: np.fft.ifft(a, axis=0)),
                         (2, lambda a: np.fft.ifft(a, axis=0) * Nz)):
        call.b2_fft_z(ctx.handle, d_in[0].ptr, d_out[0].ptr, Nz, Nr, inverse, None)
        assert_close(d_out[0].get(), ref(arrs[0]), 1e-13, 'single, inverse=%d' % inverse)
        call.b2_fft_z_multi(ctx.handle, 5, ptr_array(d_in), ptr_array(d_out), Nz, Nr, inverse, None)
        for k in range(5):
            assert_close(d_out[k].get(), ref(arrs[k]), 1e-13, 'multi %d, inverse=%d' % (k, inverse))
    # in place
    call.b2_fft_z(ctx.handle, d_in[1].ptr, d_in[1].ptr, Nz, Nr, 0, None)
    assert_close(d_in[1].get(), np.fft.fft(arrs[1], axis=0), 1e-13, 'in place')

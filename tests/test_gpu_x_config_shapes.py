"""GPU parity at the SHAPES of the BASELINE.json configurations (the radial size, mode count, particles per cell and
solver of each; a z-slab of a few hundred cells so that the oracle finishes in seconds):
  C2  LWFA, Nr=256, Nm=2, 2x2x4 ppc, standard PSATD, laser-like seed field          (+ transforms at 4096 x 256)
  C4  high-mode case, Nr=512, Nm=4, 2x2x16 ppc (the 512 x 512 Hankel GEMM)            (+ transforms at 2048 x 512)
  C5  boosted-frame Galilean PSATD, Nr=256, Nm=2, 2x2x8 ppc, electrons + ions flowing at gamma = 10
CUDA step() against the oracle on the same seeded inputs; tolerances as in test_gpu_step.py."""
import os
import numpy as np
import pytest
from scipy.constants import c

from conftest import assert_close

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get('B2_SHAPE_SCALE', '1'))        # < 1: reduced sizes for the CPU host-flow run


def _n(v):
    return max(int(round(v * SCALE)), 8)


def _run_case(Nz, Nr, Nm, p_nt, fused, v_comoving=None, ions=False, n_order=-1, nsteps=3, sort_period=1):
    from fbpic_b200 import Simulation
    from oracle import oracle as orc
    np.random.seed(2)
    dz, dr = 0.05e-6, 0.4e-6
    zmax, rmax = Nz * dz, Nr * dr
    dt = dz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2,
                     p_nt=p_nt, n_e=4.e24, n_order=n_order, n_guard=(None if n_order == -1 else 16),
                     v_comoving=v_comoving, use_galilean=(v_comoving is not None), initialize_ions=ions, fused=fused,
                     sort_period=sort_period)
    w0 = 0.2 * rmax
    k0 = 2 * np.pi / zmax * 3
    for sp in sim.ptcl:
        g = np.exp(-(sp.x**2 + sp.y**2) / w0**2)
        sp.uz = 0.05 * np.sin(k0 * sp.z) * g * (1 + sp.x / w0 + (sp.x**2 - sp.y**2) / w0**2) * (-1. if sp.q > 0 else 1.)
        sp.ux = 0.02 * np.cos(k0 * sp.z) * g * sp.y / w0
        if v_comoving is not None:
            sp.uz += -np.sqrt(1. / (1 - (v_comoving / c)**2) - 1)
        sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    # a laser-like seed field in mode 1 (C2 is laser-driven)
    g1 = sim.fld.interp[1]
    zz, rr = np.meshgrid(g1.z, g1.r, indexing='ij')
    prof = 1.e11 * np.exp(-(zz - 0.5 * zmax)**2 / (0.15 * zmax)**2) * np.exp(-rr**2 / w0**2) * np.cos(8 * k0 * zz)
    g1.Er[:, :], g1.Et[:, :] = 0.5 * prof, -0.5j * prof
    g1.Br[:, :], g1.Bt[:, :] = 0.5j * prof / c, 0.5 * prof / c
    ref = orc.OracleSim(Nz, zmax, Nr, rmax, Nm, dt, n_order=n_order, v_comoving=v_comoving,
                        use_galilean=(v_comoving is not None))
    for sp in sim.ptcl:
        ref.add_species(sp.q, sp.m, sp.x, sp.y, sp.z, sp.ux, sp.uy, sp.uz, sp.inv_gamma, sp.w)
    for k in ('Er', 'Et', 'Br', 'Bt'):
        ref.interp[1][k][:, :] = getattr(g1, k)
    sim.step(nsteps)
    ref.step(nsteps)
    assert abs(sim.fld.interp[0].zmin - ref.zmin) <= 1e-12 * zmax
    for grp, names in (('E', ('Er', 'Et', 'Ez')), ('B', ('Br', 'Bt', 'Bz')), ('J', ('Jr', 'Jt', 'Jz')), ('rho', ('rho',))):
        scale = max(np.abs(ref.interp[m][k]).max() for m in range(Nm) for k in names)
        for m in range(Nm):
            for k in names:
                assert_close(getattr(sim.fld.interp[m], k), ref.interp[m][k], 1e-9, '%s m%d' % (k, m), scale=scale)
    for i, sp in enumerate(sim.ptcl):
        r = ref.species[i]
        got = np.stack([getattr(sp, k) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w')])
        want = np.stack([r[k] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w')])
        go, wo = np.lexsort((got[2], got[1], got[0], got[6])), np.lexsort((want[2], want[1], want[0], want[6]))
        # one scale per vector: a component that stays at rounding level (uy of a plasma driven in x, z) is
        # compared on the scale of the momentum, not of its own noise
        s_x = 2 * max(np.abs(want[j]).max() for j in (0, 1))
        s_u = 2 * max(np.abs(want[j]).max() for j in (3, 4, 5))
        for j, k in enumerate(('x', 'y', 'z', 'ux', 'uy', 'uz')):
            assert_close(got[j][go], want[j][wo], 1e-10, 'species %d %s' % (i, k),
                         scale=(s_x if j < 2 else None if j == 2 else s_u))


@pytest.mark.parametrize('fused', [False, True])
def test_c2_shape(fused):
    _run_case(_n(256), _n(256), 2, 4, fused)


def test_c2_shape_as_benched():
    """C2 shape with the bench's own settings: fused step, sort_period = 4, 9 cycles (two re-sorts, the cycles in
    between deposit and gather on particles that have moved since their last sort)."""
    _run_case(_n(256), _n(256), 2, 4, True, nsteps=9, sort_period=4)


@pytest.mark.parametrize('fused', [False, True])
def test_c4_shape(fused):
    _run_case(_n(64), _n(512), 4, 16, fused)


@pytest.mark.parametrize('fused', [False, True])
def test_c5_shape(fused):
    _run_case(_n(256), _n(256), 2, 8, fused, v_comoving=-c * np.sqrt(1. - 1. / 10.**2), ions=True, n_order=32)


@pytest.mark.parametrize('Nz,Nr', [(4096, 256), (2048, 512)])
def test_transforms_at_config_sizes(Nz, Nr):
    """z-FFT + Hankel GEMM at the full grid sizes of C2 and C4 against NumPy (modes 0-2)."""
    import test_gpu_kernels
    test_gpu_kernels.test_transforms_vs_numpy(_n(Nz), _n(Nr))


def test_mode3_transforms_512():
    """the order-2/3/4 Hankel matrices of the fourth mode at Nr = 512 (C4)"""
    from fbpic_b200.fields import SpectralTransformer
    from fbpic_b200._lib import DeviceArray
    Nz, Nr = _n(256), _n(512)
    rng = np.random.default_rng(11)
    tr = SpectralTransformer(Nz, Nr, 3, 200.e-6)
    f = rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
    h = rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
    d_f, d_h = DeviceArray.from_numpy(f), DeviceArray.from_numpy(h)
    d_s, d_p, d_m = [DeviceArray((Nz, Nr), np.complex128) for _ in range(3)]
    tr.interp2spect_scal(d_f, d_s)
    assert_close(d_s.get(), np.fft.fft(f, axis=0) @ tr.dht0.M, 1e-13, 'fwd scal m3')
    tr.interp2spect_vect(d_f, d_h, d_p, d_m)
    fr, ft = np.fft.fft(f, axis=0), np.fft.fft(h, axis=0)
    assert_close(d_p.get(), (0.5 * (fr - 1.j * ft)) @ tr.dhtp.M, 1e-13, 'fwd p m3')
    assert_close(d_m.get(), (0.5 * (fr + 1.j * ft)) @ tr.dhtm.M, 1e-13, 'fwd m m3')

"""GPU parity tests of `ExternalField` (fbpic/lpa_utils/external_fields.py; SURVEY 2: user-defined fields applied
after the gather): the user's Python function is translated to CUDA C, JIT-compiled by NVRTC inside the
library and applied on the device; whole steps against golden outputs of the unmodified reference."""
import ctypes
import math
import numpy as np
import pytest
from scipy.constants import c

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu


def undulator_field(F, x, y, z, t, amplitude, length_scale):
    return F + amplitude * math.cos(2 * np.pi * z / length_scale)


def focusing_field(F, x, y, z, t, amplitude, length_scale):
    k = 2 * math.pi / length_scale
    if z > 2.e-6:
        g = math.exp(-(x**2 + y**2) / length_scale**2)
    else:
        g = 0.
    return F - amplitude * k * x * g * math.sin(k * (z - 299792458. * t))


def _dev(*arrays):
    from fbpic_b200._lib import DeviceArray
    return [DeviceArray.from_numpy(np.ascontiguousarray(a)) for a in arrays]


def test_external_field_jit():
    """NVRTC -> cubin -> cudaLibraryLoadData -> launch, lab frame and boosted (z, t) arguments."""
    from fbpic_b200 import _lib
    rng = np.random.default_rng(7)
    n = 1000
    F, x, y, z = rng.normal(size=n), rng.normal(size=n) * 1e-6, rng.normal(size=n) * 1e-6, rng.uniform(0, 2e-5, n)
    h = ctypes.c_void_p()
    _lib.call.b2_external_field_compile(b'    F_[i_] = F + amplitude * cos(z / length_scale) * x - t * 1.e9 * y;',
                                        ctypes.byref(h))
    for g, b in ((1., 0.), (3., np.sqrt(1 - 1 / 9.))):
        d = _dev(F, x, y, z)
        t, amp, L = 5.e-15, 2.5, 3.e-6
        _lib.call.b2_external_field_apply(_lib.context().handle, h, n, d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, t, amp, L,
                                          g, b, None)
        zl, tl = g * (z + b * c * t), g * (t + b * (1. / c) * z)
        assert_close(d[0].get(), F + amp * np.cos(zl / L) * x - tl * 1.e9 * y, 1e-14, 'external field g=%g' % g)
    _lib.call.b2_external_field_free(h)


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', ['lab', 'boost'])
def test_external_fields_step_vs_reference_golden(tag, fused):
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.external_fields import ExternalField
    g = load_golden('step_external_' + tag)
    gb = float(g['gamma_boost']) or None
    np.random.seed(3)
    zmax, rmax = float(g['zmax']), float(g['rmax'])
    sim = Simulation(int(g['Nz']), zmax, int(g['Nr']), rmax, int(g['Nm']), float(g['dt']),
                     p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2, p_nt=4, n_e=1.e23, n_order=-1,
                     gamma_boost=gb, boundaries={'z': 'periodic', 'r': 'reflective'}, fused=fused)
    sp = sim.ptcl[0]
    assert sp.Ntot == len(g['s0_in_x'])
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        assert_close(getattr(sp, k), g['s0_in_' + k], 1e-14, 'initial ' + k)
        setattr(sp, k, g['s0_in_' + k].copy())
    sim.external_fields = [
        ExternalField(undulator_field, 'By', 40., 5.e-6, gamma_boost=gb),
        ExternalField(focusing_field, 'Ex', 3.e10, 4.e-6, species=sp, gamma_boost=gb)]
    sim.step(int(g['nsteps']))
    names = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
    ref = np.stack([g['s0_out_' + k] for k in names])
    got = np.stack([getattr(sp, k) for k in names])
    assert got.shape == ref.shape
    # the GPU path reorders the particles; match them through their (unchanged) weight and initial lattice
    ro, go = np.lexsort((ref[2], ref[1], ref[0], ref[7])), np.lexsort((got[2], got[1], got[0], got[7]))
    for j, k in enumerate(names):
        assert_close(got[j][go], ref[j][ro], 1e-10, 'external %s %s' % (tag, k))
    Nm = sim.fld.Nm
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         'external %s %s m%d' % (tag, k, m), scale=sc)


def test_external_field_string_expression():
    """The CUDA C expression form, applied to a species on the device, against NumPy."""
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.external_fields import ExternalField
    from scipy.constants import c
    zmax, rmax = 10.e-6, 6.e-6
    np.random.seed(1)
    sim = Simulation(16, zmax, 8, rmax, 2, zmax / 16 / c, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4, n_e=1.e23)
    sp = sim.ptcl[0]
    x, y, z = sp.x.copy(), sp.y.copy(), sp.z.copy()
    sp.Ez = np.linspace(-1., 1., sp.Ntot)
    F0 = sp.Ez.copy()
    ext = ExternalField('F + amplitude * exp(-(x*x + y*y) / (length_scale*length_scale)) * sin(z / length_scale + 1e14 * t)',
                        'Ez', 7., 3.e-6)
    sp.send_particles_to_gpu()
    ext.apply_expression(sim.ptcl, 2.e-15)
    sp.receive_particles_from_gpu()
    want = F0 + 7. * np.exp(-(x * x + y * y) / (3.e-6)**2) * np.sin(z / 3.e-6 + 1e14 * 2.e-15)
    assert_close(sp.Ez, want, 1e-14, 'string expression')


# ------------------------------------------------------------------ the reference's tests/test_external_fields.py
def _laser_func(F, x, y, z, t, amplitude, length_scale):
    """the user function of the reference's test (tests/test_external_fields.py:140-144), verbatim semantics: a
    plane wave added to the gathered field; `c` and `np` are module-level names, `math` an imported module"""
    return (F + amplitude * math.cos(2 * np.pi * (z - c * t) / length_scale))


@pytest.mark.parametrize('gamma_boost', [None, 10])
def test_external_fields_as_written(gamma_boost):
    """tests/test_external_fields.py (test_external_fields_lab / _boost) as written: particles in an external plane
    wave (Ex, By through `ExternalField`) follow ux = a0 sin(k0' (z - ct)), uz = -gamma0 beta0 + gamma0 (1 - beta0)
    ux^2 / 2 over two laser periods (400 calls of step(1)), lab frame and boosted frame (gamma = 10), atol 5e-2."""
    from scipy.constants import e, m_e
    from fbpic_b200 import Simulation
    from fbpic_b200.lpa_utils.boosted_frame import BoostConverter
    from fbpic_b200.lpa_utils.external_fields import ExternalField
    Nz, Nr, Nm, zmin, zmax, rmax = 5, 10, 2, 0.e-6, 0.8e-6, 2.e-6
    a0, lambda0 = 1., 0.8e-6
    k0 = 2 * np.pi / lambda0
    dt, N_step = lambda0 / c / 200, 400
    boost = BoostConverter(gamma0=1.) if gamma_boost is None else BoostConverter(gamma_boost)
    if gamma_boost is not None:
        dt = dt * (1. + boost.beta0) / boost.gamma0
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, initialize_ions=False, zmin=zmin,
                     boundaries={'z': 'periodic', 'r': 'reflective'}, gamma_boost=gamma_boost)
    sim.ptcl = []
    sim.add_new_species(-e, m_e, n=1., p_rmax=rmax / Nr, p_nz=1, p_nr=1, p_nt=1)
    sim.external_fields = [ExternalField(_laser_func, 'Ex', a0 * m_e * c**2 * k0 / e, lambda0, gamma_boost=gamma_boost),
                           ExternalField(_laser_func, 'By', a0 * m_e * c * k0 / e, lambda0, gamma_boost=gamma_boost)]
    sp = sim.ptcl[0]
    Nptcl = sp.Ntot
    z, ux, uz = (np.zeros((N_step, Nptcl)) for _ in range(3))
    k0p = k0 * boost.gamma0 * (1. - boost.beta0)
    sp.ux = a0 * np.sin(k0p * sp.z)
    sp.uz[:] = -boost.gamma0 * boost.beta0 + boost.gamma0 * (1 - boost.beta0) * 0.5 * sp.ux**2
    for i in range(N_step):
        z[i, :], ux[i, :], uz[i, :] = sp.z[:], sp.ux[:], sp.uz[:]
        sim.step(1)
    t = sim.dt * np.arange(N_step)
    ux_analytical = a0 * np.sin(k0p * (z - c * t[:, None]))
    uz_analytical = -boost.gamma0 * boost.beta0 + boost.gamma0 * (1 - boost.beta0) * 0.5 * ux_analytical**2
    assert np.allclose(ux, ux_analytical, atol=5.e-2)
    assert np.allclose(uz, uz_analytical, atol=5.e-2)
    assert np.abs(ux).max() > 0.9

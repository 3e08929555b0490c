"""
oracle/gen_golden.py -- generates tests/golden/*.npz from the UNMODIFIED FBPIC
reference (CPU/numba path).  Test infrastructure; runs only in the build
container where /root/reference exists:

    NUMBA_THREADING_LAYER=omp OPENBLAS_NUM_THREADS=1 NUMBA_NUM_THREADS=4 \
        python oracle/gen_golden.py

The reference is imported through oracle/ref_shim (a scipy-backed stand-in for
the missing `pyfftw`).  Every fixture stores the inputs and
the reference's outputs so that the oracle (tests, not gpu) and the CUDA path
(tests, gpu) are checked against the same numbers.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('FBPIC_REFERENCE', '/root/reference')
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), REF, ROOT]
OUT = os.path.join(ROOT, 'tests', 'golden')

from scipy.constants import c, e, m_e, m_p          # noqa: E402
from fbpic.main import Simulation                    # noqa: E402
from fbpic.fields.spectral_transform.hankel import DHT   # noqa: E402
from fbpic.fields.psatd_coefs import PsatdCoeffs     # noqa: E402
from fbpic.fields.utility_methods import get_modified_k, get_stencil_reach  # noqa: E402


def save(name, **arrs):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrs)
    print('wrote %s (%.1f kB)' % (path, os.path.getsize(path) / 1e3))


def ptcl_arrays(sp):
    return {k: getattr(sp, k).copy() for k in
            ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')}


def field_arrays(sim, kinds=('E', 'B', 'J', 'rho')):
    out = {}
    for m in range(sim.fld.Nm):
        g = sim.fld.interp[m]
        for k in kinds:
            names = ['rho'] if k == 'rho' else [k + 'r', k + 't', k + 'z']
            for n in names:
                out['%s_m%d' % (n, m)] = getattr(g, n).copy()
    return out


# ---------------------------------------------------------------------------
# 1. host tables
# ---------------------------------------------------------------------------
def gen_tables():
    Nr, Nz, rmax, dz = 12, 16, 15.e-6, 0.4e-6
    dt = dz / c
    out = dict(Nr=Nr, Nz=Nz, rmax=rmax, dz=dz, dt=dt)
    for m in range(3):
        for p in (m - 1, m, m + 1):
            d = DHT(p, m, Nr, Nz, rmax)
            out['M_p%d_m%d' % (p + 1, m)] = d.M
            out['invM_p%d_m%d' % (p + 1, m)] = d.invM
            out['nu_m%d' % m] = d.nu
    sim = Simulation(Nz, Nz * dz, Nr, rmax, 3, dt, zmin=0., n_order=8, n_guard=8, verbose_level=0)
    for m in range(3):
        g = sim.fld.interp[m]
        out['invvol_m%d' % m] = g.invvol
        out['ruyten_linear_m%d' % m] = g.ruyten_linear_coef
        out['ruyten_cubic_m%d' % m] = g.ruyten_cubic_coef
        s = sim.fld.spect[m]
        out['filter_z_m%d' % m] = s.filter_array_z
        out['filter_r_m%d' % m] = s.filter_array_r
        out['inv_k2_m%d' % m] = s.inv_k2
        out['kz_m%d' % m] = s.kz[:, 0]
        out['kr_m%d' % m] = s.kr[0, :]
    kz_true = 2 * np.pi * np.fft.fftfreq(Nz, dz)
    for n_order in (8, 16):
        out['kzmod_%d' % n_order] = get_modified_k(kz_true, n_order, dz)
    variants = {'std': (None, False), 'gal': (-0.97 * c, True), 'com': (-0.97 * c, False)}
    s1 = sim.fld.spect[1]
    for tag, (V, gal) in variants.items():
        ps = PsatdCoeffs(s1.kz, s1.kr, 1, dt, Nz, Nr, V=V, use_galilean=gal)
        for k in ('C', 'S_w', 'j_coef', 'rho_prev_coef', 'rho_next_coef') + \
                (() if V is None else ('T_eb', 'T_cc', 'T_rho', 'j_corr_coef')):
            out['psatd_%s_%s' % (tag, k)] = getattr(ps, k)
    out['reach'] = np.array([get_stencil_reach(256, dz, c * dt, 16, None, False),
                             get_stencil_reach(256, dz, c * dt, 32, None, False),
                             get_stencil_reach(256, dz, c * dt, 16, -0.97 * c, True)])
    save('tables', **out)


# ---------------------------------------------------------------------------
# 2. kernel-level fixtures: cell keys, deposit, gather, push
# ---------------------------------------------------------------------------
def random_species(sim, n, rng, rmax, zmin, zmax):
    """Overwrite species 0 with n random particles that exercise the edge cases:
    on-axis, r beyond the box, first/last z cells."""
    sp = sim.ptcl[0]
    r = rmax * 1.04 * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    z = zmin + (zmax - zmin) * rng.random(n)
    r[:4] = [0., 1.e-9 * rmax, 0.49 * rmax / sim.fld.Nr, 0.51 * rmax / sim.fld.Nr]
    z[4:8] = [zmin, zmin + 1e-3 * (zmax - zmin) / sim.fld.Nz,
              zmax - 1e-3 * (zmax - zmin) / sim.fld.Nz, zmin + 0.5 * (zmax - zmin) / sim.fld.Nz]
    sp.x, sp.y, sp.z = r * np.cos(th), r * np.sin(th), z
    sp.x[0], sp.y[0] = 0., 0.
    sp.ux, sp.uy, sp.uz = [rng.normal(size=n) * s for s in (0.7, 0.5, 2.0)]
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    sp.w = rng.random(n) * 1.e6 + 1.e5
    sp.Ntot = n
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        setattr(sp, k, np.zeros(n))
    return sp


def gen_kernels(shape, Nm):
    rng = np.random.default_rng(1234 + Nm + (7 if shape == 'cubic' else 0))
    Nz, Nr, rmax, zmin, zmax = 12, 10, 10.e-6, -2.e-6, 5.2e-6
    dt = (zmax - zmin) / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=zmin, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=1, p_nr=1, p_nt=4, n_e=1.e24, zmin=zmin, particle_shape=shape,
                     verbose_level=0)
    n = 600
    sp = random_species(sim, n, rng, rmax, zmin, zmax)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, rmax=rmax, zmin=zmin, zmax=zmax, dt=dt, q=sp.q, m=sp.m)
    out.update({'p_' + k: v for k, v in ptcl_arrays(sp).items()})
    # deposit rho and J (reference call pattern: tests/test_uniform_rho_deposition.py:61-66)
    for ft in ('rho', 'J'):
        sim.fld.erase(ft)
        sp.deposit(sim.fld, ft)
        sim.fld.sum_reduce_deposition_array(ft)
        sim.fld.divide_by_volume(ft)
    out.update({'dep_' + k: v for k, v in field_arrays(sim, ('J', 'rho')).items()})
    # gather from random E, B grids
    for m in range(Nm):
        g = sim.fld.interp[m]
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            a = rng.normal(size=(Nz, Nr)) + 1.j * rng.normal(size=(Nz, Nr))
            if m == 0:
                a = a.real + 0.j
            scale = 1.e10 if k[0] == 'E' else 30.
            getattr(g, k)[:, :] = a * scale
    out.update({'grid_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    sp.gather(sim.fld.interp, sim.comm)
    for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
        out['gath_' + k] = getattr(sp, k).copy()
    # push
    sp.push_p(0.)
    sp.push_x(0.5 * dt)
    out.update({'push_' + k: v for k, v in ptcl_arrays(sp).items()})
    save('kernels_%s_Nm%d' % (shape, Nm), **out)


# ---------------------------------------------------------------------------
# 3. whole-step fixtures: a small periodic plasma wave
#    (parameters scaled down from tests/test_periodic_plasma_wave.py:134-165)
# ---------------------------------------------------------------------------
def impart_momenta(sp, epsilon, k0, w0, wp):
    """Linear plasma-wave initial momenta (restated from
    tests/test_periodic_plasma_wave.py:300-311: mode-0 part only)."""
    r2 = sp.x**2 + sp.y**2
    sp.uz[:] = -epsilon * wp / (c * k0) * np.exp(-r2 / w0**2) * np.sin(k0 * sp.z) * k0 * c / wp
    sp.ux[:] = epsilon * 2 * sp.x / (k0 * w0**2) * np.exp(-r2 / w0**2) * np.cos(k0 * sp.z)
    sp.uy[:] = epsilon * 2 * sp.y / (k0 * w0**2) * np.exp(-r2 / w0**2) * np.cos(k0 * sp.z)
    sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)


def gen_step(tag, shape, Nm, n_order, v_comoving, use_galilean, nsteps=3, ions=False, open_z=False):
    np.random.seed(0)
    Nz, Nr, zmax, rmax = 24, 12, 12.e-6, 8.e-6
    dt = zmax / Nz / c
    n_e = 2.e24
    pz0, pz1 = (0.25 * zmax, 0.75 * zmax) if open_z else (0, zmax)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=pz0, p_zmax=pz1, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4 * max(Nm - 1, 1), n_e=n_e, n_order=n_order,
                     particle_shape=shape, v_comoving=v_comoving, use_galilean=use_galilean,
                     initialize_ions=ions, verbose_level=0,
                     n_guard=(16 if open_z else (None if n_order == -1 else 8)),
                     n_damp={'z': 16, 'r': 32},
                     boundaries={'z': ('open' if open_z else 'periodic'), 'r': 'reflective'})
    k0 = 2 * np.pi / zmax * 2
    wp = np.sqrt(n_e * e**2 / (m_e * 8.8541878128e-12))
    impart_momenta(sim.ptcl[0], 0.05, k0, 3.e-6, wp)
    if v_comoving is not None:
        # flowing plasma, as in boosted-frame runs
        g = 1. / np.sqrt(1 - (v_comoving / c)**2)
        for sp in sim.ptcl:
            sp.uz += -np.sqrt(g**2 - 1)
            sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, n_order=n_order, nsteps=nsteps,
               v_comoving=(0. if v_comoving is None else v_comoving),
               has_v=(v_comoving is not None), use_galilean=use_galilean, n_species=len(sim.ptcl),
               open_z=open_z, Nz_local=sim.fld.interp[0].Nz)
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_in_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
        out['s%d_q' % i], out['s%d_m' % i] = sp.q, sp.m
    sim.step(nsteps, show_progress=False)
    for i, sp in enumerate(sim.ptcl):
        out.update({'s%d_out_%s' % (i, k): v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_' + tag, **out)


def window_dens(z, r):
    """density ramp used by the moving-window fixture (restated identically in tests/test_gpu_step.py)"""
    return np.clip((z - 5.e-6) / 3.e-6, 0., 1.)


def gen_window(nsteps=26):
    """Open-z box with a moving window (v = c) and continuous plasma injection."""
    np.random.seed(5)
    Nz, Nr, Nm, zmax, rmax = 40, 12, 2, 20.e-6, 8.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=5.e-6, p_zmax=30.e-6, p_rmin=0, p_rmax=6.e-6,
                     p_nz=2, p_nr=2, p_nt=4, n_e=1.e24, dens_func=window_dens, n_order=-1,
                     n_guard=16, n_damp={'z': 16, 'r': 32}, verbose_level=0,
                     boundaries={'z': 'open', 'r': 'reflective'})
    sim.set_moving_window(v=c)
    g1 = sim.fld.interp[1]
    zz, rr = np.meshgrid(g1.z, g1.r, indexing='ij')
    prof = 2.e11 * np.exp(-(zz - 12.e-6)**2 / (3.e-6)**2) * np.exp(-rr**2 / (3.e-6)**2) * np.cos(2 * np.pi * (zz - 12.e-6) / 2.e-6)
    g1.Er[:, :], g1.Et[:, :] = 0.5 * prof, -0.5j * prof
    g1.Br[:, :], g1.Bt[:, :] = 0.5j * prof / c, 0.5 * prof / c
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, nsteps=nsteps, Nz_local=sim.fld.interp[0].Nz)
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'in_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    np.random.seed(7)
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_moving_window', **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if '--only-window' in sys.argv:
        gen_window()
        sys.exit(0)
    if '--only-open' in sys.argv:
        gen_step('linear_open', 'linear', 2, -1, None, False, nsteps=5, open_z=True)
        sys.exit(0)
    gen_tables()
    for shape in ('linear', 'cubic'):
        for Nm in (1, 2, 3):
            gen_kernels(shape, Nm)
    gen_step('linear_std', 'linear', 2, -1, None, False)
    gen_step('cubic_std', 'cubic', 2, -1, None, False)
    gen_step('linear_Nm3_order8', 'linear', 3, 8, None, False)
    gen_step('linear_galilean', 'linear', 2, 16, -0.995 * c, True, ions=True)
    gen_step('linear_comoving', 'linear', 2, 16, -0.995 * c, False, ions=True)
    gen_step('linear_open', 'linear', 2, -1, None, False, nsteps=5, open_z=True)
    gen_window()

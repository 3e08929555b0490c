"""
oracle/gen_golden_ext.py -- fixtures for the solver variants around the hot loop:
radial PML, cross-deposition, laser antenna, external fields, boosted-frame set-up.  Like
oracle/gen_golden.py it runs the UNMODIFIED FBPIC reference (CPU/numba path, through
oracle/ref_shim) and only in the build container:

    NUMBA_THREADING_LAYER=omp OPENBLAS_NUM_THREADS=1 NUMBA_NUM_THREADS=4 \
        python oracle/gen_golden_ext.py [--only NAME]

Test infrastructure.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import save, ptcl_arrays, field_arrays, impart_momenta   # noqa: E402  (sets sys.path)
from scipy.constants import c, e, m_e                                      # noqa: E402
from fbpic.main import Simulation                                          # noqa: E402

PML_NAMES = ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')


def diverging_mode1_field(sim, z0, w0, amp, lam=1.6e-6, ctau=2.5e-6):
    """A short, tightly focused pulse in mode 1 (x-polarised-laser structure, restated from
    gen_golden.gen_window): it diffracts into the radial boundary within a few steps."""
    g1 = sim.fld.interp[1]
    zz, rr = np.meshgrid(g1.z, g1.r, indexing='ij')
    prof = amp * np.exp(-(zz - z0)**2 / ctau**2) * np.exp(-rr**2 / w0**2) * np.cos(2 * np.pi * (zz - z0) / lam)
    g1.Er[:, :], g1.Et[:, :] = 0.5 * prof, -0.5j * prof
    g1.Br[:, :], g1.Bt[:, :] = 0.5j * prof / c, 0.5 * prof / c


def gen_pml(tag, open_z, v_comoving=None, use_galilean=False, nsteps=6, window=False):
    np.random.seed(11)
    Nz, Nr, Nm, zmax, rmax = 32, 10, 2, 16.e-6, 6.e-6
    dt = zmax / Nz / c
    n_order = 8 if (v_comoving is not None) else -1
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=(4.e-6 if open_z else 0.), p_zmax=(12.e-6 if open_z else zmax),
                     p_rmin=0, p_rmax=rmax, p_nz=2, p_nr=2, p_nt=4, n_e=5.e23, n_order=n_order,
                     v_comoving=v_comoving, use_galilean=use_galilean, verbose_level=0,
                     n_guard=(12 if open_z else None), n_damp={'z': 12, 'r': 6},
                     boundaries={'z': ('open' if open_z else 'periodic'), 'r': 'open'})
    if window:
        sim.set_moving_window(v=c)
    assert sim.use_pml
    diverging_mode1_field(sim, 0.5 * zmax, 1.8e-6, 3.e11)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, n_order=n_order, nsteps=nsteps, open_z=open_z,
               nr_damp=6, nz_damp=12, n_guard=sim.comm.n_guard, window=window,
               v_comoving=(0. if v_comoving is None else v_comoving), has_v=(v_comoving is not None),
               use_galilean=use_galilean, Nz_local=sim.fld.interp[0].Nz, Nr_local=sim.fld.interp[0].Nr)
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['s0_q'], out['s0_m'] = sp.q, sp.m
    out.update({'in_' + k: v for k, v in field_arrays(sim, ('E', 'B')).items()})
    np.random.seed(12)
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    for m in range(Nm):
        for k in PML_NAMES:
            out['out_%s_m%d' % (k, m)] = getattr(sim.fld.interp[m], k).copy()
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_pml_' + tag, **out)


def gen_cross(tag, v_comoving, use_galilean, nsteps=3):
    """Periodic plasma wave with current_correction='cross-deposition' (main.py:512-514, 672-717)."""
    np.random.seed(0)
    Nz, Nr, Nm, zmax, rmax = 24, 12, 2, 12.e-6, 8.e-6
    dt = zmax / Nz / c
    n_e = 2.e24
    n_order = -1 if v_comoving is None else 16
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=4, n_e=n_e, n_order=n_order, n_guard=(None if n_order == -1 else 8),
                     v_comoving=v_comoving, use_galilean=use_galilean, verbose_level=0,
                     current_correction='cross-deposition',
                     boundaries={'z': 'periodic', 'r': 'reflective'})
    k0 = 2 * np.pi / zmax * 2
    wp = np.sqrt(n_e * e**2 / (m_e * 8.8541878128e-12))
    impart_momenta(sim.ptcl[0], 0.05, k0, 3.e-6, wp)
    if v_comoving is not None:
        g = 1. / np.sqrt(1 - (v_comoving / c)**2)
        for sp in sim.ptcl:
            sp.uz += -np.sqrt(g**2 - 1)
            sp.inv_gamma[:] = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    out = dict(Nz=Nz, Nr=Nr, Nm=Nm, zmax=zmax, rmax=rmax, dt=dt, n_order=n_order, nsteps=nsteps,
               v_comoving=(0. if v_comoving is None else v_comoving), has_v=(v_comoving is not None),
               use_galilean=use_galilean, n_species=1, open_z=False, Nz_local=Nz)
    sp = sim.ptcl[0]
    out.update({'s0_in_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out['s0_q'], out['s0_m'] = sp.q, sp.m
    sim.step(nsteps, show_progress=False)
    out.update({'s0_out_%s' % k: v for k, v in ptcl_arrays(sp).items()})
    out.update({'out_' + k: v for k, v in field_arrays(sim).items()})
    out['zmin_end'] = sim.fld.interp[0].zmin
    save('step_cross_' + tag, **out)


GENERATORS = {
    'pml_periodic': lambda: gen_pml('periodic', False),
    'pml_open': lambda: gen_pml('open', True),
    'pml_galilean': lambda: gen_pml('galilean', True, v_comoving=0.999 * c, use_galilean=True),
    'pml_window': lambda: gen_pml('window', True, window=True, nsteps=10),
    'cross_std': lambda: gen_cross('std', None, False),
    'cross_galilean': lambda: gen_cross('galilean', -0.995 * c, True),
}

if __name__ == '__main__':
    only = sys.argv[sys.argv.index('--only') + 1] if '--only' in sys.argv else None
    for name, fn in GENERATORS.items():
        if only is None or only == name:
            fn()

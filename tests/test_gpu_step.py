"""GPU parity tests (whole PIC cycle): Simulation.step on the B200 against the golden
outputs of the unmodified reference's CPU path (same inputs), fused and unfused."""
import numpy as np
import pytest

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu

TAGS = ['linear_std', 'cubic_std', 'linear_Nm3_order8', 'linear_galilean', 'linear_comoving', 'linear_open']


def build_sim(g, tag, fused):
    from fbpic_b200 import Simulation
    from scipy.constants import e, m_e, m_p
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    V = float(g['v_comoving']) if bool(g['has_v']) else None
    n_order = int(g['n_order'])
    open_z = bool(g['open_z']) if 'open_z' in g else False
    sim = Simulation(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']),
                     n_order=n_order, v_comoving=V, use_galilean=bool(g['use_galilean']),
                     particle_shape=('cubic' if 'cubic' in tag else 'linear'),
                     n_guard=(16 if open_z else (None if n_order == -1 else 8)), n_damp={'z': 16, 'r': 32},
                     boundaries={'z': ('open' if open_z else 'periodic'), 'r': 'reflective'}, fused=fused)
    if open_z:      # open z: guard + damp + injection cells around the physical box, sin^2 damping of E, B
        assert sim.fld.interp[0].Nz == int(g['Nz_local'])
    for i in range(int(g['n_species'])):
        sp = sim.add_new_species(q=float(g['s%d_q' % i]), m=float(g['s%d_m' % i]))
        for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
            setattr(sp, k, g['s%d_in_%s' % (i, k)].copy())
        sp.Ntot = len(sp.x)
        for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
            setattr(sp, k, np.zeros(sp.Ntot))
    return sim


@pytest.mark.parametrize('fused', [False, True, 3])
@pytest.mark.parametrize('tag', TAGS)
def test_step_vs_reference_golden(tag, fused):
    """fused: False = one kernel per reference operator; True = fused kernels; 3 = fused with the
    particle arrays re-sorted only every 3rd step (sort_period=3)."""
    g = load_golden('step_' + tag)
    Nm = int(g['Nm'])
    sim = build_sim(g, tag, bool(fused))
    if fused == 3:
        sim.sort_period = 3
        for sp in sim.ptcl:
            sp.sort_period = 3
    sim.step(int(g['nsteps']))
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * abs(float(g['zmax']))
    for i, sp in enumerate(sim.ptcl):
        # the GPU path reorders particles (cell sort): compare as sorted multisets keyed by w and z
        ref = np.stack([g['s%d_out_%s' % (i, k)] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')])
        got = np.stack([getattr(sp, k) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')])
        assert got.shape == ref.shape
        ro, go = np.lexsort((ref[2], ref[0], ref[7])), np.lexsort((got[2], got[0], got[7]))
        for j, k in enumerate(('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma')):
            assert_close(got[j][go], ref[j][ro], 1e-10, '%s species %d %s' % (tag, i, k))
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         '%s %s m%d' % (tag, k, m), scale=sc)


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
def test_step_Nm4_vs_oracle(shape):
    """High-mode case (BASELINE configs[3]: Nm=4): 4 PIC cycles, CUDA vs the oracle."""
    from fbpic_b200 import Simulation
    from oracle import oracle as orc
    from scipy.constants import c
    np.random.seed(2)
    Nz, Nr, Nm, zmax, rmax = 40, 20, 4, 16.e-6, 10.e-6
    dt = zmax / Nz / c
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax,
                     p_nz=2, p_nr=2, p_nt=16, n_e=2.e24, particle_shape=shape)
    sp = sim.ptcl[0]
    k0 = 2 * np.pi / zmax * 2
    g = np.exp(-(sp.x**2 + sp.y**2) / (3.e-6)**2)
    sp.uz = 0.05 * np.sin(k0 * sp.z) * g * (1 + sp.x / 3.e-6 + (sp.x**2 - sp.y**2) / (3.e-6)**2)
    sp.ux = 0.02 * np.cos(k0 * sp.z) * g * sp.y / 3.e-6
    sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
    ref = orc.OracleSim(Nz, zmax, Nr, rmax, Nm, dt, particle_shape=shape, nthreads=2)
    ref.add_species(sp.q, sp.m, sp.x, sp.y, sp.z, sp.ux, sp.uy, sp.uz, sp.inv_gamma, sp.w)
    sim.step(4)
    ref.step(4)
    for grp, names in (('E', ('Er', 'Et', 'Ez')), ('B', ('Br', 'Bt', 'Bz')), ('J', ('Jr', 'Jt', 'Jz')), ('rho', ('rho',))):
        scale = max(np.abs(ref.interp[m][k]).max() for m in range(Nm) for k in names)
        for m in range(Nm):
            for k in names:
                assert_close(getattr(sim.fld.interp[m], k), ref.interp[m][k], 1e-9, '%s m%d' % (k, m), scale=scale)


def _window_dens(z, r):
    return np.clip((z - 5.e-6) / 3.e-6, 0., 1.)


@pytest.mark.parametrize('fused', [False, True])
def test_moving_window_vs_reference_golden(fused):
    """Open-z box, moving window at c, continuous plasma injection, 26 cycles (the window moves and
    injects several times): grids shifted in spectral space, particles dropped at the left edge and
    injected at the right one -- against the unmodified reference (fbpic/boundaries/moving_window.py,
    particles/injection/continuous_injection.py) run on the same inputs."""
    from fbpic_b200 import Simulation
    from scipy.constants import c
    g = load_golden('step_moving_window')
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    np.random.seed(5)       # the azimuthal offsets of the particle loader come from np.random (gen_golden.py)
    sim = Simulation(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']),
                     p_zmin=5.e-6, p_zmax=30.e-6, p_rmin=0, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4, n_e=1.e24,
                     dens_func=_window_dens, n_order=-1, n_guard=16, n_damp={'z': 16, 'r': 32},
                     boundaries={'z': 'open', 'r': 'reflective'}, fused=fused)
    sim.set_moving_window(v=c)
    assert sim.fld.interp[0].Nz == int(g['Nz_local'])
    sp = sim.ptcl[0]
    # the initial plasma comes out of the same generator as the reference's
    assert sp.Ntot == len(g['s0_in_x'])
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):
        assert_close(getattr(sp, k), g['s0_in_' + k], 1e-14, 'initial ' + k)
        setattr(sp, k, g['s0_in_' + k].copy())
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            getattr(sim.fld.interp[m], k)[:, :] = g['in_%s_m%d' % (k, m)]
    np.random.seed(7)       # ... and so do those of the continuously injected plasma
    sim.step(int(g['nsteps']))
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * abs(float(g['zmax']))
    names = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
    ref = np.stack([g['s0_out_' + k] for k in names])
    got = np.stack([getattr(sp, k) for k in names])
    assert got.shape == ref.shape, 'particle count after drop + injection: %s vs %s' % (got.shape, ref.shape)
    ro, go = np.lexsort((ref[2], ref[1], ref[0], ref[7])), np.lexsort((got[2], got[1], got[0], got[7]))
    for j, k in enumerate(names):
        assert_close(got[j][go], ref[j][ro], 1e-9, 'window %s' % k)
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         'window %s m%d' % (k, m), scale=sc)

"""CPU tests: the oracle (oracle/) and the host tables are pinned against the
golden vectors generated from the unmodified FBPIC reference by
oracle/gen_golden.py."""
import numpy as np
import pytest
from scipy.constants import c

from conftest import load_golden, assert_close, group_scale
from fbpic_b200 import host_tables as ht
from oracle import oracle as orc


def test_tables_bit_exact():
    g = load_golden('tables')
    Nr, Nz, rmax, dz, dt = int(g['Nr']), int(g['Nz']), float(g['rmax']), float(g['dz']), float(g['dt'])
    kz_true = 2 * np.pi * np.fft.fftfreq(Nz, dz)
    dzs = (Nz * dz) / Nz            # the grid spacing as the Simulation computes it
    kz_true_s = 2 * np.pi * np.fft.fftfreq(Nz, dzs)
    rmax_s = Nr * (rmax / Nr)       # idem for rmax (boundary_communicator.get_rmax)
    for m in range(3):
        for p in (m - 1, m, m + 1):
            M, iM, nu = ht.hankel_matrices(p, m, Nr, rmax)
            assert np.array_equal(M, g['M_p%d_m%d' % (p + 1, m)])
            assert np.array_equal(iM, g['invM_p%d_m%d' % (p + 1, m)])
            assert np.array_equal(nu, g['nu_m%d' % m])
        vol = ht.cell_volumes(m, Nr, rmax_s, dzs)
        assert np.array_equal(1. / vol, g['invvol_m%d' % m])
        lin, cub = ht.ruyten_coefs(vol, rmax_s / Nr, dzs)
        assert np.array_equal(lin, g['ruyten_linear_m%d' % m])
        assert np.array_equal(cub, g['ruyten_cubic_m%d' % m])
        kr = 2 * np.pi * ht.hankel_matrices(m, m, Nr, rmax_s)[2]
        kz = ht.modified_kz(kz_true_s, 8, dzs)
        assert np.array_equal(kr, g['kr_m%d' % m]) and np.array_equal(kz, g['kz_m%d' % m])
        fz, fr = ht.binomial_filters(kz_true_s, kr, dzs, rmax_s / Nr)
        assert np.array_equal(fz, g['filter_z_m%d' % m]) and np.array_equal(fr, g['filter_r_m%d' % m])
        assert np.array_equal(ht.inverse_k2(kz, kr), g['inv_k2_m%d' % m])
    for n in (8, 16):
        assert np.array_equal(ht.modified_kz(kz_true, n, dz), g['kzmod_%d' % n])
    kz, kr = g['kz_m1'], g['kr_m1']
    for tag, (V, gal) in {'std': (None, False), 'gal': (-0.97 * c, True), 'com': (-0.97 * c, False)}.items():
        t = ht.psatd_coefficients(kz, kr, dt, V, gal)
        for k, v in t.items():
            assert np.array_equal(v, g['psatd_%s_%s' % (tag, k)]), (tag, k)
    reach = [ht.stencil_reach(256, dz, c * dt, 16, None, False),
             ht.stencil_reach(256, dz, c * dt, 32, None, False),
             ht.stencil_reach(256, dz, c * dt, 16, -0.97 * c, True)]
    assert list(g['reach']) == reach


@pytest.mark.parametrize('shape', ['linear', 'cubic'])
@pytest.mark.parametrize('Nm', [1, 2, 3])
def test_oracle_kernels_vs_reference(shape, Nm):
    g = load_golden('kernels_%s_Nm%d' % (shape, Nm))
    Nz, Nr = int(g['Nz']), int(g['Nr'])
    rmax, zmin, zmax, dt = float(g['rmax']), float(g['zmin']), float(g['zmax']), float(g['dt'])
    q, mass = float(g['q']), float(g['m'])
    invdz, invdr = Nz / (zmax - zmin), Nr / rmax
    # reference grid attributes: invdz = 1/dz with dz=(zmax-zmin)/Nz
    invdz, invdr = 1. / ((zmax - zmin) / Nz), 1. / (rmax / Nr)
    P = {k: g['p_' + k].copy() for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')}
    cubic = (shape == 'cubic')
    vols = [ht.cell_volumes(m, Nr, rmax, (zmax - zmin) / Nz) for m in range(Nm)]
    ruy = [ht.ruyten_coefs(v, rmax / Nr, (zmax - zmin) / Nz)[1 if cubic else 0] for v in vols]
    for what, names in (('rho', ('rho',)), ('J', ('Jr', 'Jt', 'Jz'))):
        for nthreads in (1, 3):
            raw = orc.deposit(what, P['x'], P['y'], P['z'], P['w'], q, P['ux'], P['uy'], P['uz'], P['inv_gamma'],
                              invdz, zmin, Nz, invdr, 0., Nr, Nm, cubic, ruy[0], ruy[1 if Nm > 1 else 0], nthreads)
            for m in range(Nm):
                for k, nme in enumerate(names):
                    assert_close(raw[k, m] / vols[m][None, :], g['dep_%s_m%d' % (nme, m)], 1e-13,
                                 'deposit %s m%d' % (nme, m))
    grids = [tuple(g['grid_%s_m%d' % (k, m)] for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz')) for m in range(Nm)]
    n = len(P['x'])
    F = {k: np.zeros(n) for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz')}
    orc.gather(P['x'], P['y'], P['z'], rmax, invdz, zmin, Nz, invdr, 0., Nr, grids, cubic,
               F['Ex'], F['Ey'], F['Ez'], F['Bx'], F['By'], F['Bz'])
    for k in F:
        assert_close(F[k], g['gath_' + k], 1e-13, 'gather ' + k)
    orc.push_p(P['ux'], P['uy'], P['uz'], P['inv_gamma'], *[g['gath_' + k] for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz')],
               q, mass, dt)
    orc.push_x(P['x'], P['y'], P['z'], P['ux'], P['uy'], P['uz'], P['inv_gamma'], 0.5 * dt)
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma'):
        assert_close(P[k], g['push_' + k], 1e-14, 'push ' + k)


def test_sort_contract():
    g = load_golden('kernels_linear_Nm2')
    Nz, Nr = int(g['Nz']), int(g['Nr'])
    zmin, zmax, rmax = float(g['zmin']), float(g['zmax']), float(g['rmax'])
    cell = orc.cell_index(g['p_x'], g['p_y'], g['p_z'], 1. / ((zmax - zmin) / Nz), zmin, Nz, 1. / (rmax / Nr), 0., Nr)
    assert cell.min() >= 0 and cell.max() < Nz * (Nr + 1)
    idx, prefix = orc.sort_contract(cell, Nz, Nr)
    assert np.all(np.diff(cell[idx]) >= 0) and prefix[-1] == len(cell)
    same = np.diff(cell[idx]) == 0
    assert np.all(np.diff(idx)[same] > 0)        # stability


@pytest.mark.parametrize('tag', ['linear_std', 'cubic_std', 'linear_Nm3_order8',
                                 'linear_galilean', 'linear_comoving'])
def test_oracle_step_vs_reference(tag):
    g = load_golden('step_' + tag)
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    V = float(g['v_comoving']) if bool(g['has_v']) else None
    sim = orc.OracleSim(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']),
                        n_order=int(g['n_order']), v_comoving=V, use_galilean=bool(g['use_galilean']),
                        particle_shape=('cubic' if 'cubic' in tag else 'linear'), nthreads=2)
    for i in range(int(g['n_species'])):
        sim.add_species(float(g['s%d_q' % i]), float(g['s%d_m' % i]),
                        *[g['s%d_in_%s' % (i, k)] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')])
    sim.step(int(g['nsteps']))
    assert abs(sim.zmin - float(g['zmin_end'])) <= 1e-12 * abs(float(g['zmax']))
    for i, sp in enumerate(sim.species):
        for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma'):
            assert_close(sp[k], g['s%d_out_%s' % (i, k)], 1e-11, '%s species %d %s' % (tag, i, k))
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(sim.interp[m][k], g['out_%s_m%d' % (k, m)], 1e-10, '%s %s m%d' % (tag, k, m), scale=sc)


def _oracle_from_step_golden(g, **kw):
    Nz, Nr, Nm = int(g['Nz']), int(g['Nr']), int(g['Nm'])
    V = float(g['v_comoving']) if bool(g['has_v']) else None
    sim = orc.OracleSim(Nz, float(g['zmax']), Nr, float(g['rmax']), Nm, float(g['dt']),
                        n_order=int(g['n_order']), v_comoving=V, use_galilean=bool(g['use_galilean']),
                        nthreads=2, **kw)
    sim.add_species(float(g['s0_q']), float(g['s0_m']),
                    *[g['s0_in_%s' % k] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')])
    return sim


def _check_step_outputs(sim, g, tag, names, ptol=1e-11, ftol=1e-10):
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma'):
        assert_close(sim.species[0][k], g['s0_out_%s' % k], ptol, '%s %s' % (tag, k))
    for m in range(sim.Nm):
        for k in names:
            grp = 'rho' if k == 'rho' else k[0]
            sc = group_scale(g, 'out_', grp, sim.Nm)
            assert_close(sim.interp[m][k], g['out_%s_m%d' % (k, m)], ftol, '%s %s m%d' % (tag, k, m), scale=sc)


@pytest.mark.parametrize('tag', ['std', 'galilean'])
def test_oracle_cross_deposition_vs_reference(tag):
    """current_correction='cross-deposition' (main.py:512-514, 672-717; numba_methods.py:88-116, 243-275)"""
    g = load_golden('step_cross_' + tag)
    sim = _oracle_from_step_golden(g, current_correction='cross-deposition')
    sim.step(int(g['nsteps']))
    assert abs(sim.zmin - float(g['zmin_end'])) <= 1e-12 * abs(float(g['zmax']))
    _check_step_outputs(sim, g, 'cross ' + tag, ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'))


def test_oracle_pml_vs_reference():
    """boundaries['r']='open': split PML components, their spectral push and the radial damping
    (main.py:410-415, 732-761; numba_methods.py:189-214; pml_damping.py:46-83), z periodic."""
    g = load_golden('step_pml_periodic')
    sim = _oracle_from_step_golden(g, nr_damp=int(g['nr_damp']))
    assert sim.Nr == int(g['Nr_local'])
    for m in range(sim.Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            sim.interp[m][k][:, :] = g['in_%s_m%d' % (k, m)]
    sim.step(int(g['nsteps']))
    _check_step_outputs(sim, g, 'pml', ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho',
                                        'Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml'))

"""GPU parity of the non-default options of `Simulation` / `step()` (use_true_rho with neutralising ions, standard
and Galilean; correct_currents=False; filter_currents=False; move_positions=False; move_momenta=False; correct_divE; a single azimuthal mode; cubic shapes
with three modes) against golden outputs of the unmodified reference (oracle/gen_golden_ext.py)."""
import numpy as np
import pytest
from scipy.constants import c

from conftest import load_golden, assert_close, group_scale

pytestmark = pytest.mark.gpu

OPTIONS = {
    'true_rho': dict(step=dict(use_true_rho=True), sim=dict()),
    'true_rho_galilean': dict(step=dict(use_true_rho=True),
                              sim=dict(v_comoving=-0.995 * c, use_galilean=True, n_order=16, n_guard=8)),
    'no_correction': dict(step=dict(correct_currents=False), sim=dict()),
    'correct_divE': dict(step=dict(correct_divE=True, correct_currents=False), sim=dict()),
    'no_filter': dict(step=dict(), sim=dict(filter_currents=False)),
    'no_push_x': dict(step=dict(move_positions=False), sim=dict()),
    'no_push_p': dict(step=dict(move_momenta=False), sim=dict()),
    'nm1': dict(step=dict(), sim=dict()),
    'nm1_cubic_galilean': dict(step=dict(), sim=dict(particle_shape='cubic', v_comoving=-0.995 * c, use_galilean=True,
                                                     n_order=16, n_guard=8)),
    'cubic_true_rho_nm3': dict(step=dict(use_true_rho=True), sim=dict(particle_shape='cubic')),
}
STATE = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')


@pytest.mark.parametrize('fused', [False, True])
@pytest.mark.parametrize('tag', sorted(OPTIONS))
def test_step_options_vs_reference_golden(tag, fused):
    from fbpic_b200 import Simulation
    g = load_golden('step_opt_' + tag)
    opt = OPTIONS[tag]
    Nm = int(g['Nm'])
    sim = Simulation(int(g['Nz']), float(g['zmax']), int(g['Nr']), float(g['rmax']), Nm, float(g['dt']),
                     boundaries={'z': 'periodic', 'r': 'reflective'}, fused=fused, **opt['sim'])
    for i in range(int(g['n_species'])):
        sp = sim.add_new_species(q=float(g['s%d_q' % i]), m=float(g['s%d_m' % i]))
        for k in STATE:
            setattr(sp, k, g['s%d_in_%s' % (i, k)].copy())
        sp.Ntot = len(sp.x)
        for k in ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz'):
            setattr(sp, k, np.zeros(sp.Ntot))
    sim.step(int(g['nsteps']), **opt['step'])
    assert abs(sim.fld.interp[0].zmin - float(g['zmin_end'])) <= 1e-12 * float(g['zmax'])
    for i, sp in enumerate(sim.ptcl):
        ref = np.stack([g['s%d_out_%s' % (i, k)] for k in STATE])
        got = np.stack([getattr(sp, k) for k in STATE])
        assert got.shape == ref.shape
        ro, go = np.lexsort((ref[2], ref[0], ref[7])), np.lexsort((got[2], got[0], got[7]))
        for j, k in enumerate(STATE[:7]):
            assert_close(got[j][go], ref[j][ro], 1e-10, '%s species %d %s' % (tag, i, k))
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9,
                         '%s %s m%d' % (tag, k, m), scale=sc)


@pytest.mark.parametrize('fused', [False, True])
def test_species_mix_vs_reference_golden(fused):
    """Four species built by `add_new_species` with the same seed as the reference: thermal drifting electrons,
    doubly charged heavy ions with another sampling, a tracer species and a neutral one -- the loader draws the same
    random numbers in the same order (checked on the initial arrays), then 4 cycles."""
    from fbpic_b200 import Simulation
    from scipy.constants import e, m_e, m_p
    g = load_golden('step_species_mix')
    Nm, zmax, rmax = int(g['Nm']), float(g['zmax']), float(g['rmax'])
    np.random.seed(8)
    sim = Simulation(int(g['Nz']), zmax, int(g['Nr']), rmax, Nm, float(g['dt']),
                     boundaries={'z': 'periodic', 'r': 'reflective'}, fused=fused)
    kw = dict(p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax)
    sim.add_new_species(q=-e, m=m_e, n=2.e24, p_nz=2, p_nr=2, p_nt=4, ux_th=0.02, uy_th=0.01, uz_th=0.05, uz_m=0.1, **kw)
    sim.add_new_species(q=2 * e, m=4 * m_p, n=1.e24, p_nz=1, p_nr=2, p_nt=6, **kw)
    sim.add_new_species(q=-e, m=m_e, n=1.e20, p_nz=1, p_nr=1, p_nt=4, is_tracer=True, ux_m=0.3, **kw)
    sim.add_new_species(q=0., m=m_e, n=1.e24, p_nz=1, p_nr=1, p_nt=4, uz_m=2., uy_th=0.1, **kw)
    assert len(sim.ptcl) == int(g['n_species'])
    for i, sp in enumerate(sim.ptcl):
        for k in STATE:
            assert_close(getattr(sp, k), g['s%d_in_%s' % (i, k)], 1e-14, 'initial species %d %s' % (i, k))
    sim.step(int(g['nsteps']))
    for i, sp in enumerate(sim.ptcl):
        ref = np.stack([g['s%d_out_%s' % (i, k)] for k in STATE])
        got = np.stack([getattr(sp, k) for k in STATE])
        assert got.shape == ref.shape
        ro, go = np.lexsort((ref[2], ref[1], ref[0], ref[7])), np.lexsort((got[2], got[1], got[0], got[7]))
        for j, k in enumerate(STATE[:7]):
            assert_close(got[j][go], ref[j][ro], 1e-10, 'species %d %s' % (i, k))
    for m in range(Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho'):
            sc = group_scale(g, 'out_', 'rho' if k == 'rho' else k[0], Nm)
            assert_close(getattr(sim.fld.interp[m], k), g['out_%s_m%d' % (k, m)], 1e-9, 'mix %s m%d' % (k, m), scale=sc)

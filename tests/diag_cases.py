"""Set-ups of the diagnostics fixtures, written ONCE against a namespace `ns` (Simulation and the diagnostics classes
of either side of the drop-in boundary): oracle/gen_golden_ext.py runs them with the unmodified reference (its h5py
calls land in the stand-in of oracle/ref_shim), the tests with fbpic_b200."""
import os
import numpy as np
from scipy.constants import c, e, m_e


def build_diag_sim(ns, **sim_kw):
    """Two species (tracked electrons with a momentum modulation, ions), open z, radial PML."""
    np.random.seed(23)
    Nz, Nr, Nm, zmax, rmax = 32, 12, 2, 16.e-6, 8.e-6
    dt = zmax / Nz / c
    sim = ns.Simulation(Nz, zmax, Nr, rmax, Nm, dt, p_zmin=2.e-6, p_zmax=14.e-6, p_rmin=0, p_rmax=6.e-6, p_nz=2,
                        p_nr=2, p_nt=4, n_e=2.e24, n_order=-1, n_guard=12, n_damp={'z': 10, 'r': 5},
                        particle_shape='linear', boundaries={'z': 'open', 'r': 'open'}, **sim_kw)
    elec = sim.ptcl[0]
    elec.uz[:] = 0.4 * np.sin(2 * np.pi * elec.z / 8.e-6)
    elec.ux[:] = 0.1 * np.cos(2 * np.pi * elec.z / 16.e-6)
    elec.inv_gamma[:] = 1. / np.sqrt(1 + elec.ux**2 + elec.uy**2 + elec.uz**2)
    ions = sim.add_new_species(q=e, m=1836. * m_e, n=2.e24, p_nz=1, p_nr=2, p_nt=4, p_zmin=2.e-6, p_zmax=14.e-6,
                               p_rmin=0, p_rmax=6.e-6)
    elec.track(sim.comm)
    return sim, elec, ions


def build_cpu_gpu_deposition(ns, particle_shape, root, **sim_kw):
    """The reference's own CPU-vs-GPU parity test, as written (tests/test_cpu_gpu_deposition.py:31-84): a Gaussian
    electron bunch (N = 2000, seed 0) deposits rho and J for 3 cycles; the comparison goes through the files of a
    FieldDiagnostic with period 1."""
    Nz, zmax, zmin, Nr, rmax, Nm = 100, 30.e-6, -10.e-6, 50, 20.e-6, 2
    dt = (zmax - zmin) / Nz / c
    sim = ns.Simulation(Nz, zmax, Nr, rmax, Nm, dt, zmin=zmin, particle_shape=particle_shape, **sim_kw)
    sim.ptcl = []
    np.random.seed(0)
    ns.add_elec_bunch_gaussian(sim, 20.e-6, 10.e-6, 10.e-6, 10, 0., 10.e-12, 2000)
    sim.diags = [ns.FieldDiagnostic(1, sim.fld, fieldtypes=['rho', 'J'], comm=sim.comm, write_dir=root)]
    return sim


DIAG_DIRS = ('all', 'sel', 'dens', 'chk')


def attach_diags(ns, sim, elec, ions, root):
    """Field, particle (all quantities / with a selection), per-species charge density diagnostics every 4 cycles and
    a checkpoint every 2 cycles; returns the directories (DIAG_DIRS order; the checkpoint one is that of rank 0)."""
    d = [os.path.join(root, k) for k in DIAG_DIRS]
    sim.diags = [ns.FieldDiagnostic(4, sim.fld, sim.comm, fieldtypes=['rho', 'E', 'B', 'J'], write_dir=d[0]),
                 ns.ParticleDiagnostic(4, {'electrons': elec, 'ions': ions}, sim.comm,
                                       particle_data=['position', 'momentum', 'weighting', 'gamma', 'E', 'B'],
                                       write_dir=d[0]),
                 ns.ParticleDiagnostic(4, {'electrons': elec}, sim.comm, select={'uz': [0.05, None], 'x': [None, 3.e-6]},
                                       write_dir=d[1]),
                 ns.ParticleChargeDensityDiagnostic(4, sim, {'electrons': elec, 'ions': ions}, write_dir=d[2])]
    ns.set_periodic_checkpoint(sim, 2, checkpoint_dir=d[3])
    d[3] = os.path.join(d[3], 'proc0')
    return d


DIAG_STEPS = 5


def build_lab_diag_sim(ns, **sim_kw):
    """Boosted-frame run (gamma = 4) of a laser pulse entering a plasma, with a moving window."""
    np.random.seed(29)
    gamma_boost = 4.
    boost = ns.BoostConverter(gamma_boost)
    Nz, Nr, Nm, zmax, zmin, rmax = 64, 12, 2, 0., -32.e-6, 12.e-6
    dt_lab = (zmax - zmin) / Nz / c
    sim = ns.Simulation(Nz, zmax, Nr, rmax, Nm, dt_lab, zmin=zmin, n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 4},
                        gamma_boost=gamma_boost, v_comoving=-0.9999 * c, use_galilean=False,
                        boundaries={'z': 'open', 'r': 'reflective'}, **sim_kw)
    elec = sim.add_new_species(q=-e, m=m_e, n=1.e24, p_zmin=-26.e-6, p_zmax=400.e-6, p_rmax=10.e-6, p_nz=1, p_nr=2,
                               p_nt=4)
    elec.track(sim.comm)
    ns.add_laser_pulse(sim, ns.GaussianLaser(1.5, 5.e-6, 12.e-15, -14.e-6, lambda0=3.2e-6), gamma_boost=gamma_boost)
    v_window, = boost.velocity([c])
    sim.set_moving_window(v=v_window)
    return sim, gamma_boost


LAB_DIAG_STEPS = 40


def attach_lab_diag(ns, sim, gamma_boost, root):
    """4 lab-frame snapshots every 20 fs in a window that follows the pulse: all of E, B, J, rho and the electrons,
    flushed every 16 cycles; a selection of the electrons in a second series, flushed every 8 cycles."""
    sim.diags = [ns.BackTransformedFieldDiagnostic(-32.e-6, 0., c, 20.e-15, 4, gamma_boost, 16, sim.fld, comm=sim.comm,
                                                   fieldtypes=['E', 'B', 'J', 'rho'], write_dir=root),
                 # the (tracked) plasma electrons in the same snapshot files, and a selection of them in other ones
                 ns.BackTransformedParticleDiagnostic(-32.e-6, 0., c, 20.e-15, 4, gamma_boost, 16, sim.fld,
                                                      species={'electrons': sim.ptcl[0]}, comm=sim.comm,
                                                      write_dir=root),
                 ns.BackTransformedParticleDiagnostic(-32.e-6, 0., c, 20.e-15, 4, gamma_boost, 8, sim.fld,
                                                      select={'uz': [0.02, None], 'x': [0., None]},
                                                      species={'electrons': sim.ptcl[0]}, comm=sim.comm,
                                                      write_dir=os.path.join(root, 'selected'))]
    return root


# ---------------------------------------------------------------------------------------------------------------
# comparison of written trees with the fixture harvested from the reference
# ---------------------------------------------------------------------------------------------------------------
SKIP_ATTRS = ('@software', '@date')


def golden_files(golden, tag):
    """{file stem: {path: value}} of one diagnostics directory of the fixture"""
    out = {}
    for key in list(golden):
        if key.startswith(tag + '/'):
            name, path = key[len(tag) + 1:].split(':', 1)
            out.setdefault(name.rsplit('.', 1)[0], {})[path] = golden[key]
    return out


def written_files(directory):
    from fbpic_b200.openpmd_store import read_tree
    d = os.path.join(directory, 'hdf5')
    return {name.rsplit('.', 1)[0]: read_tree(os.path.join(d, name)) for name in sorted(os.listdir(d))
            if name.endswith(('.npz', '.h5'))}


def _particle_order(tree, group):
    """canonical order of the particles of one species group: by id when tracked, else by position"""
    if group + '/id' in tree:
        return np.argsort(tree[group + '/id'], kind='stable')
    x, y, z = (np.round(tree[group + '/position/' + k] / 1.e-11) for k in 'xyz')
    return np.lexsort((y, x, z))


def compare_trees(got, ref, tol, what):
    """Same paths and attributes; datasets equal within tol of the largest value of their record."""
    extra = sorted(k for k in got if k not in ref and not k.startswith('/restart/'))
    missing = sorted(k for k in ref if k not in got)
    assert not extra and not missing, '%s: extra %s, missing %s' % (what, extra[:5], missing[:5])
    orders = {}
    for key in ref:
        if key.endswith(SKIP_ATTRS):
            continue
        g, r = np.asarray(got[key]), np.asarray(ref[key])
        if '@' in key:
            assert g.shape == r.shape, '%s %s: %s vs %s' % (what, key, g.shape, r.shape)
            if r.dtype.kind in 'SU':
                assert np.array_equal(g.astype('S'), r.astype('S')), '%s %s: %s vs %s' % (what, key, g, r)
            else:
                assert np.allclose(g, r, rtol=1e-12, atol=0), '%s %s: %s vs %s' % (what, key, g, r)
            continue
        assert g.shape == r.shape and g.dtype == r.dtype, '%s %s: %s %s vs %s %s' % (what, key, g.shape, g.dtype,
                                                                                    r.shape, r.dtype)
        if g.size == 0:
            continue
        if '/particles/' in key:
            group = '/'.join(key.split('/')[:5])
            if group not in orders:
                orders[group] = (_particle_order(got, group), _particle_order(ref, group))
            g, r = g[orders[group][0]], r[orders[group][1]]
            record = key.rsplit('/', 1)[0] if key.count('/') > 5 else key
        else:
            record = key.rsplit('/', 1)[0] if key.rsplit('/', 1)[1] in 'rtz' else key
        if r.dtype.kind in 'ui':
            assert np.array_equal(g, r), '%s %s' % (what, key)
            continue
        scale = max(np.abs(np.asarray(ref[k])).max() for k in ref
                    if '@' not in k and (k == record or k.startswith(record + '/')) and np.asarray(ref[k]).size)
        err = np.abs(g - r).max()
        assert err <= tol * scale, '%s %s: max err %.3e vs scale %.3e' % (what, key, err, scale)

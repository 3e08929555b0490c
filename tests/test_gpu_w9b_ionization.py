"""GPU tests of ADK ionization (fbpic_b200/ionization.py; fbpic/particles/elementary_process/ionization/): the
kernel-level pieces against the formulas of the reference, conservation and plumbing (sort, window, exchange, restart).
The reference's own acceptance test (tests/test_ionization.py) is in test_gpu_y2_ionization_laser.py: its laser is an
`ExternalField`, whose kernel is compiled at run time (those tests run last)."""
import numpy as np
import pytest
from scipy.constants import c, m_e, m_p, e

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ kernel level, through the C ABI
def _dev(*arrays):
    from fbpic_b200._lib import DeviceArray
    return [DeviceArray.from_numpy(np.ascontiguousarray(a)) for a in arrays]


def test_ionize_kernel_vs_reference_probabilities():
    """b2_ionize against get_E_amplitude / get_ionization_probability of the reference on 600 random particles
    (tests/golden/ionization.npz): with given draws exactly the ions with draw < p move up one level."""
    import ctypes
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray
    from test_hostemu_ext import _ionize_case
    g, tables = _ionize_case()
    n, level_max = len(g['level']), 6
    draws = np.random.default_rng(7).uniform(size=n)
    p = g['probability']
    want = np.flatnonzero((draws < p) & (g['level'] < level_max))
    d_level, = _dev(g['level'].astype(np.uint64))
    d_tab = _dev(*tables)
    d_arr = _dev(*g['u'], *g['E'], *g['B'])
    d_draws, = _dev(draws)
    events, count, found = DeviceArray(2 * n, np.int64), DeviceArray(1, np.int64), ctypes.c_int64(-1)
    _lib.call.b2_ionize(_lib.context().handle, n, d_level.ptr, level_max, *[t.ptr for t in d_tab],
                        *[a.ptr for a in d_arr], d_draws.ptr, 0, n, events.ptr, count.ptr, ctypes.byref(found), None)
    k = found.value
    ev = events.get()[:2 * k].reshape(k, 2)
    ev = ev[np.argsort(ev[:, 0])]
    assert np.array_equal(ev[:, 0], want) and np.array_equal(ev[:, 1], g['level'][want])
    expect = g['level'].copy()
    expect[want] += 1
    assert np.array_equal(d_level.get(), expect.astype(np.uint64))
    # the built-in generator: same seed, same events; right number on average
    counts = []
    for seed in (11, 11, 12):
        d_level.set(g['level'].astype(np.uint64))
        _lib.call.b2_ionize(_lib.context().handle, n, d_level.ptr, level_max, *[t.ptr for t in d_tab],
                            *[a.ptr for a in d_arr], None, seed, n, events.ptr, count.ptr, ctypes.byref(found), None)
        counts.append(found.value)
    mean = np.where(g['level'] < level_max, p, 0.).sum()
    sigma = np.sqrt(np.where(g['level'] < level_max, p * (1 - p), 0.).sum())
    assert counts[0] == counts[1] and abs(counts[0] - mean) < 5 * sigma and abs(counts[2] - mean) < 5 * sigma


def test_push_p_ioniz_and_weights():
    """b2_push_p_ioniz (charge = level * e, neutral particles untouched) against the oracle's Vay push;
    b2_w_times_level."""
    from oracle import oracle as orc
    from fbpic_b200 import _lib
    from fbpic_b200._lib import DeviceArray
    from conftest import assert_close
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
    d_level, = _dev(level)
    d_u = _dev(*u0, ig0)
    d_f = _dev(*E, *B)
    _lib.call.b2_push_p_ioniz(_lib.context().handle, n, d_level.ptr, *[a.ptr for a in d_u], *[a.ptr for a in d_f], m, dt,
                              None)
    for g_, w_, a, name in zip(d_u, want, u0 + [ig0], ('ux', 'uy', 'uz', 'inv_gamma')):
        g_ = g_.get()
        assert np.array_equal(g_[level == 0], a[level == 0]), name
        assert_close(g_, w_, 1e-14, name)
    w = rng.uniform(1., 2., n)
    d_w, = _dev(w)
    out = DeviceArray(n, np.float64)
    _lib.call.b2_w_times_level(_lib.context().handle, n, d_w.ptr, d_level.ptr, out.ptr, None)
    assert np.array_equal(out.get(), w * level)


def _standing_field(sim, amplitude):
    """a static longitudinal field in mode 0, strong enough for gradual ionization of nitrogen (no laser needed)"""
    g = sim.fld.interp[0]
    g.Ez[:, :] = amplitude * (1 + 0.3 * np.sin(2 * np.pi * g.z / 12.e-6))[:, None]


def test_ionization_events_free_one_electron_each():
    """Static ions in a field: after every call of step() the electron species has grown by exactly the number of
    ionization events (sum of the level increases), the new electrons sit on ions and carry their weight, and the
    deposition weight of the ions is w * level."""
    from fbpic_b200 import Simulation
    Nz, Nr, Nm, zmax, rmax = 48, 8, 2, 24.e-6, 8.e-6
    np.random.seed(2)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, zmax / Nz / c, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 4},
                     boundaries={'z': 'open', 'r': 'reflective'})
    kw = dict(p_zmin=9.e-6, p_zmax=15.e-6, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4, continuous_injection=False)
    elec = sim.add_new_species(q=-e, m=m_e)
    ions = sim.add_new_species(q=0, m=14. * m_p, n=1.e18, **kw)
    ions.make_ionizable('N', target_species=elec, level_start=0, level_max=5)
    assert ions.q == e and elec.Ntot == 0
    _standing_field(sim, 1.5e11)
    total = 0
    for _ in range(3):
        levels_before, n_before = ions.ionizer.ionization_level.sum(), elec.Ntot
        sim.step(3, correct_currents=False)
        lv = ions.ionizer.ionization_level
        new = elec.Ntot - n_before
        assert new == int(lv.sum() - levels_before) and lv.max() <= 5
        assert np.array_equal(ions.ionizer.w_times_level, np.array(ions.w) * lv)
        total += new
    assert total > 50 and len(np.unique(lv)) > 1
    assert set(np.round(np.array(elec.w) / ions.w[0], 9)) <= set(np.round(np.array(ions.w) / ions.w[0], 9))
    assert np.array(elec.z).min() > 4.e-6 and np.array(elec.z).max() < 20.e-6


def test_ionizable_species_through_window_sort_and_exchange():
    """The ionization levels follow the ions through cell sorts, the removal at the left edge of a moving window and
    the continuous injection at the right edge: an ion never loses charge, injected ions start at level_start, the
    per-particle arrays keep the length of the species and the deposition weight stays w * level."""
    from fbpic_b200 import Simulation
    Nz, Nr, Nm, zmax, rmax = 48, 8, 2, 24.e-6, 8.e-6
    np.random.seed(2)
    sim = Simulation(Nz, zmax, Nr, rmax, Nm, zmax / Nz / c, zmin=0., n_order=-1, n_guard=12, n_damp={'z': 12, 'r': 4},
                     boundaries={'z': 'open', 'r': 'reflective'})
    kw = dict(p_zmin=4.e-6, p_zmax=200.e-6, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4)
    elec = sim.add_new_species(q=-e, m=m_e, n=1.e18, **kw)
    ions = sim.add_new_species(q=0, m=14. * m_p, n=1.e18, **kw)
    ions.make_ionizable('N', target_species=elec, level_start=1)
    ions.track(sim.comm)
    sim.set_moving_window(v=c)
    _standing_field(sim, 2.2e11)
    seen = {}
    n_first = ions.Ntot
    for _ in range(4):
        sim.step(5, correct_currents=False)
        ids, lv = ions.tracker.id, ions.ionizer.ionization_level
        assert len(lv) == ions.Ntot == len(ids) and lv.min() >= 1 and lv.max() <= 7
        assert np.array_equal(ions.ionizer.w_times_level, np.array(ions.w) * lv)
        for pid, level in zip(ids.tolist(), lv.tolist()):
            assert level >= seen.get(pid, 1)
            seen[pid] = level
    assert len(seen) > n_first, 'no ion was injected'          # the window brought new ions in (and dropped others)
    assert max(seen.values()) > 1, 'nothing was ionized'


def test_grow_device_arrays_beyond_capacity():
    """`Particles.grow_device_arrays` (room for freed electrons): within the allocated head room the arrays are
    re-viewed, beyond it they are reallocated; the existing particles, their ids and fields are kept either way."""
    from fbpic_b200 import Simulation, GpuMemoryManager
    np.random.seed(1)
    zmax, rmax = 8.e-6, 4.e-6
    sim = Simulation(16, zmax, 8, rmax, 2, zmax / 16 / c, p_zmin=0, p_zmax=zmax, p_rmin=0, p_rmax=rmax, p_nz=1, p_nr=1,
                     p_nt=4, n_e=1.e20)
    sp = sim.ptcl[0]
    sp.track(sim.comm)
    n0 = sp.Ntot
    x0, ids0 = np.array(sp.x), np.array(sp.tracker.id)
    with GpuMemoryManager(sim):
        cap0 = sp._capacity
        sp.grow_device_arrays(n0 + 10)                       # fits
        assert sp._capacity == cap0 and sp.Ntot == n0 + 10
        sp.x.view((10,), byte_offset=8 * n0).set(np.arange(10.))
        sp.grow_device_arrays(cap0 + 1000)                   # does not fit
        assert sp._capacity > cap0 and sp.Ntot == cap0 + 1000 and sp.cell_idx.size == sp.Ntot
        for k in ('y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w'):          # fill the rest: the data goes back to the host
            getattr(sp, k).view((sp.Ntot - n0,), byte_offset=8 * n0).fill(0)
        sp.x.view((sp.Ntot - n0 - 10,), byte_offset=8 * (n0 + 10)).fill(0)
    assert np.array_equal(np.array(sp.x)[:n0], x0) and np.array_equal(np.array(sp.x)[n0:n0 + 10], np.arange(10.))
    ids = np.array(sp.tracker.id)
    assert np.array_equal(ids[:n0], ids0) and len(np.unique(ids)) == sp.Ntot == len(sp.Ex)


def test_restart_keeps_the_ionization_levels(tmp_path):
    """A checkpoint stores the per-particle charge of an ionizable species; after `restart_from_checkpoint` the ions
    have their levels (and deposition weights) back (checkpoint_restart.py:317-323)."""
    from fbpic_b200 import Simulation
    from fbpic_b200.openpmd_diag import set_periodic_checkpoint, restart_from_checkpoint

    def build():
        Nz, Nr, Nm, zmax, rmax = 48, 8, 2, 24.e-6, 8.e-6
        np.random.seed(2)
        sim = Simulation(Nz, zmax, Nr, rmax, Nm, zmax / Nz / c, zmin=0., n_order=-1, n_guard=12,
                         n_damp={'z': 12, 'r': 4}, boundaries={'z': 'open', 'r': 'reflective'})
        kw = dict(p_zmin=9.e-6, p_zmax=15.e-6, p_rmax=6.e-6, p_nz=2, p_nr=2, p_nt=4, continuous_injection=False)
        elec = sim.add_new_species(q=-e, m=m_e)
        ions = sim.add_new_species(q=0, m=14. * m_p, n=1.e18, **kw)
        ions.make_ionizable('N', target_species=elec, level_start=0, level_max=5)
        return sim, elec, ions
    a, elec_a, ions_a = build()
    _standing_field(a, 1.5e11)
    set_periodic_checkpoint(a, 6, checkpoint_dir=str(tmp_path))
    a.step(6, correct_currents=False)
    b, elec_b, ions_b = build()
    restart_from_checkpoint(b, checkpoint_dir=str(tmp_path))
    assert b.iteration == 6 and elec_b.Ntot == elec_a.Ntot > 0 and ions_b.Ntot == ions_a.Ntot
    oa, ob = np.lexsort((ions_a.x, ions_a.z)), np.lexsort((ions_b.x, ions_b.z))
    la, lb = ions_a.ionizer.ionization_level[oa], ions_b.ionizer.ionization_level[ob]
    assert np.array_equal(la, lb) and la.max() > 0
    assert np.allclose(ions_b.ionizer.w_times_level[ob], ions_a.ionizer.w_times_level[oa], rtol=1e-14, atol=0)
    b.step(2, correct_currents=False)                      # and the restarted run goes on ionizing
    assert ions_b.ionizer.ionization_level.sum() >= lb.sum()

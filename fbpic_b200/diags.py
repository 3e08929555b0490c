"""
Diagnostics and checkpoints that plug into `Simulation.step()` the way the reference's do
(`sim.diags`, `sim.checkpoints`: objects with a `write(iteration)` method, called after the gather resp. at the
end of every cycle, fbpic/main.py:474-481, 563-565; fbpic/openpmd_diag/{generic_diag,field_diag,particle_diag,
checkpoint_restart}.py).

Difference to the reference, stated plainly: the reference writes openPMD/HDF5 through h5py, which is not
available to this build; these classes keep its constructor arguments, periods, gathering rules and the openPMD
thetaMode array layout ([2 Nm - 1, Nr, Nz]: mode 0, then 2 Re / 2 Im of every mode m > 0; field_diag.py:177-190)
but store NumPy `.npz` archives whose keys are the openPMD record paths (`fields/E/r`, `particles/<name>/
position/x`, ...).  Any other object with a `write(iteration)` method (e.g. a user's own openPMD writer) can be
put in `sim.diags` as well.

Only the arrays a diagnostic asks for are read back from HBM (one D2H copy each); the simulation data stays on
the device -- the reference does a full `receive_fields_from_gpu()` / `send_fields_to_gpu()` round trip per
output (field_diag.py:98-99, 156-157).
"""
import glob
import os
import re
import numpy as np

from ._lib import DeviceArray


def _host(a):
    return a.get() if isinstance(a, DeviceArray) else np.asarray(a)


class NpzDiagnostic(object):
    """Period / iteration-window logic of OpenPMDDiagnostic (generic_diag.py:20-141)."""

    def __init__(self, period, comm, write_dir=None, iteration_min=0, iteration_max=np.inf, dt_period=None,
                 dt_sim=None):
        self.rank = comm.rank if comm is not None else 0
        if (period is None) and (dt_period is None):
            raise ValueError("You need to pass either `period` or `dt_period`to the diagnostics.")
        if (period is not None) and (dt_period is not None):
            raise ValueError("You need to pass either `period` or `dt_period`to the diagnostics, \n"
                             "but do not pass both.")
        if period is None:
            period = dt_period / dt_sim
        self.period = max(1, int(round(period)))
        self.iteration_min, self.iteration_max = iteration_min, iteration_max
        self.comm = comm
        self.write_dir = os.path.join(os.getcwd(), 'diags') if write_dir is None else os.path.abspath(write_dir)
        if self.rank == 0 or comm is None:
            os.makedirs(os.path.join(self.write_dir, 'npz'), exist_ok=True)

    def is_due(self, iteration):
        return iteration % self.period == 0 and self.iteration_min <= iteration < self.iteration_max

    def write(self, iteration):
        if self.is_due(iteration):
            self.write_npz(iteration)

    def path(self, kind, iteration):
        return os.path.join(self.write_dir, 'npz', '%s%08d.npz' % (kind, iteration))


class FieldDiagnostic(NpzDiagnostic):
    """Fields on the grid (field_diag.py:11-190).  With `comm` the data is gathered on every rank, guard / damp /
    PML cells removed, and rank 0 writes; without, each rank writes its own arrays, guard cells included."""

    def __init__(self, period=None, fldobject=None, comm=None, fieldtypes=["rho", "E", "B", "J"], write_dir=None,
                 iteration_min=0, iteration_max=np.inf, dt_period=None, keep_mode0_imag=False):
        """`keep_mode0_imag` (used by the checkpoints): also store the imaginary part of the mode-0 arrays, which
        the openPMD layout drops.  It is at rounding level for rho but not for E, B (the (r,t) -> (p,m)
        combination mixes Im Et into Re Er), so a restart without it is only approximate -- as in the
        reference."""
        if fldobject is None:
            raise ValueError("You need to pass the argument `fldobject` to `FieldDiagnostic`.")
        NpzDiagnostic.__init__(self, period, comm, write_dir, iteration_min, iteration_max, dt_period=dt_period,
                               dt_sim=fldobject.dt)
        self.fld, self.fieldtypes = fldobject, list(fieldtypes)
        self.keep_mode0_imag = keep_mode0_imag

    def _dataset(self, quantity):
        """[2 Nm - 1, Nr, Nz] real array of one field component (field_diag.py:177-212)."""
        modes = []
        for m in range(self.fld.Nm):
            a = _host(getattr(self.fld.interp[m], quantity))
            if self.comm is not None:
                a = self.comm.gather_grid_array(a)
            modes.append(a.T)
        out = np.empty((2 * self.fld.Nm - 1,) + modes[0].shape)
        out[0] = modes[0].real
        for m in range(1, self.fld.Nm):
            out[2 * m - 1], out[2 * m] = 2 * modes[m].real, 2 * modes[m].imag
        self._imag0 = np.ascontiguousarray(modes[0].imag)
        return out

    def _store(self, out, key, quantity):
        out[key] = self._dataset(quantity)
        if self.keep_mode0_imag:
            out[key + '/imag0'] = self._imag0

    def write_npz(self, iteration):
        fld, comm = self.fld, self.comm
        multi = (comm is not None) and (comm.size > 1)
        if "rho" in self.fieldtypes:      # bring the (smoothed) sources back from spectral space
            fld.spect2interp('rho_prev')
            if multi and not fld.exchanged_source['rho_prev']:
                comm.exchange_fields(fld.interp, 'rho', 'add')
        if "J" in self.fieldtypes:
            fld.spect2interp('J')
            if multi and not fld.exchanged_source['J']:
                comm.exchange_fields(fld.interp, 'J', 'add')
        g0 = fld.interp[0]
        if comm is None:
            zmin, Nz, Nr = g0.zmin, g0.Nz, g0.Nr
        else:
            zmin, _ = comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
            Nz, _ = comm.get_Nz_and_iz(local=False, with_damp=False, with_guard=False)
            Nr = comm.get_Nr(with_damp=False)
        out = {'meta/iteration': iteration, 'meta/time': iteration * fld.dt, 'meta/dt': fld.dt, 'meta/zmin': zmin,
               'meta/dz': g0.dz, 'meta/dr': g0.dr, 'meta/Nz': Nz, 'meta/Nr': Nr, 'meta/Nm': fld.Nm,
               'meta/geometry': 'thetaMode', 'meta/axisLabels': 'r,z'}
        for ft in self.fieldtypes:
            if ft == "rho":
                self._store(out, 'fields/rho', 'rho')
            elif ft in ("E", "B", "J"):
                for coord in ('r', 't', 'z'):
                    self._store(out, 'fields/%s/%s' % (ft, coord), ft + coord)
            elif ft.endswith("_pml"):
                self._store(out, 'fields/' + ft, ft)
            else:
                raise ValueError("Invalid string in fieldtypes: %s" % ft)
        if self.rank == 0 or comm is None:
            np.savez(self.path('fields', iteration), **out)


_QUANTITIES = {'position': ('x', 'y', 'z'), 'momentum': ('ux', 'uy', 'uz'), 'weighting': ('w',),
               'gamma': ('gamma',), 'E': ('Ex', 'Ey', 'Ez'), 'B': ('Bx', 'By', 'Bz'), 'id': ('id',)}


class ParticleDiagnostic(NpzDiagnostic):
    """Particle phase space (particle_diag.py:14-520).  `select`: {'uz': [1., None], ...} keeps the particles
    whose quantity lies in the interval; quantities: x, y, z, ux, uy, uz, gamma."""

    def __init__(self, period=None, species={}, comm=None, particle_data=["position", "momentum", "weighting"],
                 select=None, write_dir=None, iteration_min=0, iteration_max=np.inf, dt_period=None):
        if len(species) == 0:
            raise ValueError("You need to pass an non-empty `species_dict`.")
        dt = list(species.values())[0].dt
        NpzDiagnostic.__init__(self, period, comm, write_dir, iteration_min, iteration_max, dt_period=dt_period,
                               dt_sim=dt)
        for q in particle_data:
            if q not in _QUANTITIES:
                raise ValueError("Invalid string in particle_data: %s" % q)
        self.species_dict, self.particle_data, self.select = dict(species), list(particle_data), select
        self.dt = dt
        # tracked species get their ids written as well (particle_diag.py:111-116)
        if 'id' not in self.particle_data and any(sp.tracker is not None for sp in self.species_dict.values()):
            self.particle_data.append('id')

    @staticmethod
    def _attr(sp, name):
        if name == 'gamma':
            return 1. / _host(sp.inv_gamma)
        if name == 'id':
            if sp.tracker is None:
                raise ValueError('The species is not tracked: call `species.track(sim.comm)` first.')
            return _host(sp.tracker.id)
        return _host(getattr(sp, name))

    def write_npz(self, iteration):
        out = {'meta/iteration': iteration, 'meta/time': iteration * self.dt, 'meta/dt': self.dt,
               'meta/species': np.array(sorted(self.species_dict))}
        for name, sp in self.species_dict.items():
            n = sp.Ntot
            keep = np.ones(n, dtype=bool)
            if self.select is not None:         # particle_diag.py:366-410
                for q, (lo, hi) in self.select.items():
                    v = self._attr(sp, q)
                    if lo is not None:
                        keep &= (v > lo)
                    if hi is not None:
                        keep &= (v < hi)
            data = {}
            quantities = [q for q in self.particle_data if not (q == 'id' and sp.tracker is None)]
            for q in quantities:
                for comp in _QUANTITIES[q]:
                    data[comp] = self._attr(sp, comp)[:n][keep]
            if self.comm is not None and self.comm.size > 1:
                parts = [None] * self.comm.size
                self.comm._host_group().all_gather_object(parts, data)
                data = {k: np.concatenate([p[k] for p in parts]) for k in data}
            grp = 'particles/%s/' % name
            out[grp + 'charge'], out[grp + 'mass'] = sp.q, sp.m
            for q in quantities:
                for comp in _QUANTITIES[q]:
                    key = q if len(_QUANTITIES[q]) == 1 else '%s/%s' % (q, comp[-1])
                    out[grp + key] = data[comp]
        if self.rank == 0 or self.comm is None:
            np.savez(self.path('particles', iteration), **out)


# ---------------------------------------------------------------------------
# checkpoint / restart (checkpoint_restart.py:22-330)
# ---------------------------------------------------------------------------
def set_periodic_checkpoint(sim, period, checkpoint_dir='./checkpoints'):
    """Every `period` cycles each rank saves its E, B (+ PML components) with guard cells and all its particles
    (checkpoint_restart.py:22-75)."""
    comm = sim.comm
    if comm.rank == 0:
        os.makedirs(checkpoint_dir, exist_ok=True)
    write_dir = os.path.join(checkpoint_dir, 'proc%d/' % comm.rank)
    fieldtypes = ["E", "B"]
    if sim.use_pml:
        fieldtypes += ["Er_pml", "Et_pml", "Br_pml", "Bt_pml"]
    sim.checkpoints.append(FieldDiagnostic(period, sim.fld, fieldtypes=fieldtypes, write_dir=write_dir,
                                           keep_mode0_imag=True))
    species = {'species %d' % i: sp for i, sp in enumerate(sim.ptcl)}
    if species:
        sim.checkpoints.append(ParticleDiagnostic(period, species, write_dir=write_dir))


def restart_from_checkpoint(sim, iteration=None, checkpoint_dir='./checkpoints'):
    """Load E, B (+ PML components), the particles, the iteration / time and the position of the (moving) grid
    from a checkpoint written by a simulation with the same set-up and number of ranks
    (checkpoint_restart.py:77-189).  Call before `step()` (host copy of the data)."""
    from .particles import FIELD_ATTRS
    comm = sim.comm
    data_dir = os.path.join(checkpoint_dir, 'proc%d' % comm.rank, 'npz')
    its = sorted(int(re.search(r'fields(\d+)\.npz$', f).group(1)) for f in glob.glob(os.path.join(data_dir, 'fields*.npz')))
    if not its:
        raise RuntimeError('The directory %s, which is required to restart a simulation from checkpoints, '
                           'holds no checkpoint.' % data_dir)
    if iteration is None:
        iteration = its[-1]
    elif iteration not in its:
        raise RuntimeError('The iteration %d is not among the checkpoints (%s).' % (iteration, its))
    f = np.load(os.path.join(data_dir, 'fields%08d.npz' % iteration))
    g0 = sim.fld.interp[0]
    if (int(f['meta/Nz']), int(f['meta/Nr']), int(f['meta/Nm'])) != (g0.Nz, g0.Nr, sim.fld.Nm):
        raise RuntimeError('The checkpoint was written with a different grid: the local grid (with guard cells) '
                           'must be identical, which also requires the same number of ranks.')
    sim.iteration, sim.time = iteration, float(f['meta/time'])
    names = ['Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'] + (['Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml'] if sim.use_pml else [])
    for m in range(sim.fld.Nm):
        for k in names:
            key = 'fields/' + (k if k.endswith('_pml') else '%s/%s' % (k[0], k[1]))
            d = f[key]
            if m == 0:
                a = d[0] + (1.j * f[key + '/imag0'] if key + '/imag0' in f.files else 0.j)
            else:
                a = 0.5 * (d[2 * m - 1] + 1.j * d[2 * m])
            getattr(sim.fld.interp[m], k)[:, :] = a.T
    zmin_old, zmin_new = g0.zmin, float(f['meta/zmin'])
    for m in range(sim.fld.Nm):
        length = sim.fld.interp[m].zmax - sim.fld.interp[m].zmin
        sim.fld.interp[m].zmin, sim.fld.interp[m].zmax = zmin_new, zmin_new + length
    comm.shift_global_domain_positions(zmin_new - zmin_old)
    if sim.ptcl:
        p = np.load(os.path.join(data_dir, 'particles%08d.npz' % iteration))
        if len(p['meta/species']) != len(sim.ptcl):
            raise RuntimeError('Species numbers in checkpoint and simulation should be same, but got %d and %d. '
                               'Use add_new_species method to add species to simulation or sim.ptcl = [] to remove '
                               'them' % (len(p['meta/species']), len(sim.ptcl)))
        for i, sp in enumerate(sim.ptcl):
            grp = 'particles/species %d/' % i
            for attr, key in (('x', 'position/x'), ('y', 'position/y'), ('z', 'position/z'), ('ux', 'momentum/x'),
                              ('uy', 'momentum/y'), ('uz', 'momentum/z'), ('w', 'weighting')):
                setattr(sp, attr, np.ascontiguousarray(p[grp + key], dtype=np.float64))
            sp.Ntot = len(sp.x)
            if sp.tracker is not None:
                sp.tracker.overwrite_ids(p[grp + 'id'], comm)
            sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
            for k in FIELD_ATTRS:
                setattr(sp, k, np.zeros(sp.Ntot))
            if hasattr(sp.injector, 'reset_injection_positions'):
                sp.injector.reset_injection_positions()

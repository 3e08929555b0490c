"""
Diagnostics and checkpoints that plug into `Simulation.step()` the way the reference's do
(`sim.diags`, `sim.checkpoints`: objects with a `write(iteration)` method, called after the gather resp. at the
end of every cycle, fbpic/main.py:474-481, 563-565; fbpic/openpmd_diag/{generic_diag,field_diag,particle_diag,
particle_density_diag,boosted_field_diag,checkpoint_restart}.py).

They write the openPMD tree of the reference -- `/data/<iteration>/fields/...` in thetaMode layout ([2 Nm - 1, Nr, Nz]:
mode 0, then 2 Re / 2 Im of every mode m > 0; field_diag.py:177-196), `/data/<iteration>/particles/<species>/...` with
the momenta in kg m/s (particle_diag.py:497-500), and the openPMD attributes -- into `<write_dir>/hdf5/data%08d.h5`
through h5py when h5py is installed, else into `data%08d.npz` archives that hold the same tree
(fbpic_b200/openpmd_store.py; this build image has no HDF5 library).  tests/golden/diags_tree.npz is the tree the
unmodified reference wrote for the same simulation; the tests compare path by path.

Only the arrays a diagnostic asks for are read back from HBM (one D2H copy each; the lab-frame diagnostic reads back
one interpolated slice, 10 x (2 Nm - 1) x Nr numbers, per snapshot and cycle); the simulation data stays on the device
-- the reference does a full `receive_fields_from_gpu()` / `send_fields_to_gpu()` round trip per output
(field_diag.py:98-99, 156-157).
"""
import datetime
import os
import re
import numpy as np
from scipy.constants import c

from . import _lib
from ._lib import DeviceArray, call
from .openpmd_store import open_file, existing_file, read_tree

__version__ = '0.1'

# openPMD unitDimension: exponents of (length, mass, time, current, temperature, amount, luminous intensity).
# 'mass' carries the reference's value (data_dict.py:21, a length), not (0, 1, 0, 0): files compare equal attribute by
# attribute with the reference's.
_DIMENSION = {'rho': (-3, 0, 1, 1), 'J': (-2, 0, 0, 1), 'E': (1, 1, -3, -1), 'B': (0, 1, -2, -1),
              'charge': (0, 0, 1, 1), 'mass': (1, 0, 0, 0), 'weighting': (0, 0, 0, 0), 'position': (1, 0, 0, 0),
              'positionOffset': (1, 0, 0, 0), 'momentum': (1, 1, -1, 0), 'id': (0, 0, 0, 0), 'gamma': (0, 0, 0, 0)}
# particle records: (macroWeighted, weightingPower)
_WEIGHTING = {'charge': (0, 1.), 'mass': (0, 1.), 'weighting': (1, 1.), 'position': (0, 0.), 'positionOffset': (0, 0.),
              'momentum': (0, 1.), 'E': (0, 0.), 'B': (0, 0.), 'gamma': (0, 0.), 'id': (0, 0.)}


def _unit_dimension(record):
    if record.startswith('rho'):                 # rho_<species> of the charge-density diagnostic
        record = 'rho'
    if record.endswith('_pml'):                  # Er_pml, Bt_pml, ...
        record = record[0]
    return np.array(_DIMENSION[record] + (0, 0, 0), dtype=np.float64)


def _host(a):
    return a.get() if isinstance(a, DeviceArray) else np.asarray(a)


def _text(s):
    return np.bytes_(s)


class OpenPMDDiagnostic(object):
    """Period / iteration-window logic and the file-level openPMD attributes (generic_diag.py:20-230)."""

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
        if self.rank == 0:
            os.makedirs(os.path.join(self.write_dir, 'hdf5'), exist_ok=True)

    def write(self, iteration):
        if iteration % self.period == 0 and self.iteration_min <= iteration < self.iteration_max:
            self.write_hdf5(iteration)

    def file_stem(self, iteration):
        return os.path.join(self.write_dir, 'hdf5', 'data%08d' % iteration)

    def open_file(self, stem):
        """The container file on rank 0, None elsewhere (generic_diag.py:98-120)."""
        return open_file(stem, 'a') if self.rank == 0 else None

    def setup_openpmd_file(self, f, iteration, time, dt):
        now = datetime.datetime.now(datetime.timezone.utc).astimezone()
        for key, value in (('openPMD', '1.0.0'), ('software', 'fbpic_b200 ' + __version__),
                           ('date', now.strftime('%Y-%m-%d %H:%M:%S %z')), ('meshesPath', 'fields/'),
                           ('particlesPath', 'particles/'), ('iterationEncoding', 'fileBased'),
                           ('iterationFormat', 'data%T.h5'), ('basePath', '/data/%T/')):
            f.attrs[key] = _text(value)
        f.attrs['openPMDextension'] = np.uint32(1)
        base = f.require_group('/data/%d/' % iteration)
        base.attrs['time'], base.attrs['dt'], base.attrs['timeUnitSI'] = time, dt, 1.

    @staticmethod
    def setup_openpmd_record(node, record):
        node.attrs['unitDimension'] = _unit_dimension(record)
        node.attrs['timeOffset'] = 0.

    @staticmethod
    def setup_openpmd_component(node):
        node.attrs['unitSI'] = 1.


class FieldDiagnostic(OpenPMDDiagnostic):
    """Fields on the grid (field_diag.py:11-386).  With `comm` the data is gathered, guard / damp / PML cells
    removed, and rank 0 writes; without, each rank writes its own arrays, guard cells included."""
    needs_gathered_fields = False          # Simulation.step keeps the fused gather + push with this diagnostic

    def __init__(self, period=None, fldobject=None, comm=None, fieldtypes=["rho", "E", "B", "J"], write_dir=None,
                 iteration_min=0, iteration_max=np.inf, dt_period=None, keep_mode0_imag=False):
        """`keep_mode0_imag` (used by the checkpoints): also store the imaginary part of the mode-0 arrays, which
        the openPMD layout drops, under `/restart/<iteration>/` (outside of the openPMD base path).  It is at
        rounding level for rho but not for E, B (the (r,t) -> (p,m) combination mixes Im Et into Re Er), so a
        restart without it is only approximate -- as in the reference."""
        if fldobject is None:
            raise ValueError("You need to pass the argument `fldobject` to `FieldDiagnostic`.")
        OpenPMDDiagnostic.__init__(self, period, comm, write_dir, iteration_min, iteration_max, dt_period=dt_period,
                                   dt_sim=fldobject.dt)
        self.fld, self.fieldtypes = fldobject, list(fieldtypes)
        self.coords = ['r', 't', 'z']
        self.keep_mode0_imag = keep_mode0_imag

    def _components(self):
        """(path below the meshes group, attribute of the interpolation grid) of every dataset"""
        out = []
        for ft in self.fieldtypes:
            if ft.startswith("rho") or ft.endswith("_pml"):
                out.append((ft, 'rho' if ft.startswith("rho") else ft))
            elif ft in ("E", "B", "J"):
                out += [('%s/%s' % (ft, co), ft + co) for co in self.coords]
            else:
                raise ValueError("Invalid string in fieldtypes: %s" % ft)
        return out

    def output_grid(self):
        """zmin, Nz, Nr of the output (field_diag.py:106-118)"""
        g0, comm = self.fld.interp[0], self.comm
        if comm is None:
            return g0.zmin, g0.Nz, g0.Nr
        zmin, _ = comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
        Nz, _ = comm.get_Nz_and_iz(local=False, with_damp=False, with_guard=False)
        return zmin, Nz, comm.get_Nr(with_damp=False)

    def write_hdf5(self, iteration):
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
        zmin, Nz, Nr = self.output_grid()
        stem = self.file_stem(iteration)
        self.create_file_empty_meshes(stem, iteration, iteration * fld.dt, Nr, Nz, zmin, fld.interp[0].dz, fld.dt)
        f = self.open_file(stem)
        grp = f['/data/%d/fields/' % iteration] if f is not None else None
        for path, quantity in self._components():
            self.write_dataset(grp, path, quantity, f, iteration)
        if f is not None:
            f.close()

    def get_dataset(self, quantity, m):
        a = _host(getattr(self.fld.interp[m], quantity))
        return self.comm.gather_grid_array(a) if self.comm is not None else a

    def write_dataset(self, grp, path, quantity, f=None, iteration=None):
        """thetaMode rows of one component (field_diag.py:161-196)"""
        for m in range(self.fld.Nm):
            mode = self.get_dataset(quantity, m)
            if self.rank != 0:
                continue
            mode = mode.T
            dset = grp[path]
            if m == 0:
                dset[0, :, :] = mode.real
                if self.keep_mode0_imag:
                    extra = f.require_group('/restart/%d/' % iteration)
                    if quantity in extra:
                        del extra[quantity]
                    extra.create_dataset(quantity, data=np.ascontiguousarray(mode.imag))
            else:
                dset[2 * m - 1, :, :] = 2 * mode.real
                dset[2 * m, :, :] = 2 * mode.imag

    def create_file_empty_meshes(self, stem, iteration, time, Nr, Nz, zmin, dz, dt):
        """File, meshes group, zero-filled datasets and all their attributes (field_diag.py:226-303)"""
        f = self.open_file(stem)
        if f is None:
            return
        self.setup_openpmd_file(f, iteration, time, dt)
        grp = f.require_group('/data/%d/fields/' % iteration)
        self.setup_openpmd_meshes_group(grp)
        shape = (2 * self.fld.Nm - 1, Nr, Nz)
        for path, quantity in self._components():
            dset = grp.require_dataset(path, shape, dtype='f8')
            self.setup_openpmd_component(dset)
            dset.attrs['position'] = np.array([0.5, 0.5])
        for ft in self.fieldtypes:
            self.setup_openpmd_mesh_record(grp[ft], ft, dz, zmin)
        f.close()

    @staticmethod
    def setup_openpmd_meshes_group(grp):
        grp.attrs['fieldSolver'] = _text('PSATD')
        grp.attrs['fieldBoundary'] = np.array([_text('reflecting')] * 4)
        grp.attrs['particleBoundary'] = np.array([_text('absorbing')] * 4)
        grp.attrs['currentSmoothing'] = _text('Binomial')
        grp.attrs['currentSmoothingParameters'] = _text('period=1;numPasses=1;compensator=false')
        grp.attrs['chargeCorrection'] = _text('spectral')
        grp.attrs['chargeCorrectionParameters'] = _text('period=1')

    def setup_openpmd_mesh_record(self, node, record, dz, zmin):
        g0 = self.fld.interp[0]
        self.setup_openpmd_record(node, record)
        node.attrs['geometry'] = _text('thetaMode')
        node.attrs['geometryParameters'] = _text('m={:d};imag=+'.format(self.fld.Nm))
        node.attrs['gridSpacing'] = np.array([g0.dr, dz])
        node.attrs['gridGlobalOffset'] = np.array([g0.rmin, zmin])
        node.attrs['axisLabels'] = np.array([b'r', b'z'])
        node.attrs['dataOrder'] = _text('C')
        node.attrs['gridUnitSI'] = 1.
        node.attrs['fieldSmoothing'] = _text('none')


class ParticleChargeDensityDiagnostic(FieldDiagnostic):
    """Charge density of single species, one mesh `rho_<name>` each (particle_density_diag.py:11-139).  As in the
    reference the deposition overwrites rho_next in spectral space, which the PIC cycle recomputes later in the same
    iteration."""
    needs_gathered_fields = True           # deposits in the middle of the cycle: keep the plain kernel sequence

    def __init__(self, period=None, sim=None, species={}, write_dir=None, iteration_min=0, iteration_max=np.inf,
                 dt_period=None):
        if sim is None:
            raise ValueError("You need to pass the argument `sim`.")
        if len(species) == 0:
            raise ValueError("You need to pass a valid `species` dictionary.")
        FieldDiagnostic.__init__(self, period, fldobject=sim.fld, comm=sim.comm,
                                 fieldtypes=['rho_%s' % name for name in species], write_dir=write_dir,
                                 iteration_min=iteration_min, iteration_max=iteration_max, dt_period=dt_period)
        self.sim, self.species = sim, species

    def write_hdf5(self, iteration):
        sim, fld = self.sim, self.fld
        zmin, Nz, Nr = self.output_grid()
        stem = self.file_stem(iteration)
        self.create_file_empty_meshes(stem, iteration, iteration * fld.dt, Nr, Nz, zmin, fld.interp[0].dz, fld.dt)
        for name, species in self.species.items():
            sim.deposit('rho_next', species_list=[species], update_spectral=True, exchange=False)
            fld.spect2interp('rho_next')          # the filtered density, back on the grid
            if sim.comm is not None and sim.comm.size > 1:
                sim.comm.exchange_fields(fld.interp, 'rho', 'add')
            f = self.open_file(stem)
            grp = f['/data/%d/fields/' % iteration] if f is not None else None
            self.write_dataset(grp, 'rho_%s' % name, 'rho')
            if f is not None:
                f.close()


class InputScriptDiagnostic(OpenPMDDiagnostic):
    """Saves the text of the running input script (the first `*.py` among the command-line arguments) and optional
    extra attributes (`param_dict`, camelCase keys) into the file attributes of every dump
    (inputscript_diag.py:13-120)."""
    needs_gathered_fields = False

    def __init__(self, period=None, comm=None, param_dict=None, write_dir=None, iteration_min=0, iteration_max=np.inf,
                 dt_period=None, dt_sim=None):
        import sys
        OpenPMDDiagnostic.__init__(self, period, comm, write_dir, iteration_min, iteration_max, dt_period=dt_period,
                                   dt_sim=dt_sim)
        self.param_dict = param_dict
        self._input_script = None
        scripts = [a for a in sys.argv if '.py' in a]
        if scripts and os.path.isfile(scripts[0]):
            with open(scripts[0], 'r') as f:
                self._input_script = f.read()

    def write_hdf5(self, iteration):
        if self._input_script is None:
            return
        f = self.open_file(self.file_stem(iteration))
        if f is None:
            return
        f.attrs['inputScript'] = _text(self._input_script)
        for key, val in (self.param_dict or {}).items():
            if isinstance(val, str):
                f.attrs[key] = _text(val)
            elif isinstance(val, (list, tuple)):
                f.attrs[key] = np.array(val)
            else:
                f.attrs[key] = val
        f.close()


_COMPONENTS = {'position': ('x', 'y', 'z'), 'momentum': ('ux', 'uy', 'uz'), 'weighting': ('w',),
               'gamma': ('gamma',), 'E': ('Ex', 'Ey', 'Ez'), 'B': ('Bx', 'By', 'Bz'), 'id': ('id',),
               'charge': ('charge',)}


class ParticleDiagnostic(OpenPMDDiagnostic):
    """Particle phase space (particle_diag.py:14-514).  `select`: {'uz': [1., None], ...} keeps the particles
    whose quantity lies in the interval; quantities: x, y, z, ux, uy, uz, gamma."""

    def __init__(self, period=None, species={}, comm=None, particle_data=["position", "momentum", "weighting"],
                 select=None, write_dir=None, iteration_min=0, iteration_max=np.inf, subsampling_fraction=None,
                 dt_period=None):
        if len(species) == 0:
            raise ValueError("You need to pass an non-empty `species_dict`.")
        self.species_names_list = sorted(species.keys())
        self.dt = species[self.species_names_list[0]].dt
        OpenPMDDiagnostic.__init__(self, period, comm, write_dir, iteration_min, iteration_max, dt_period=dt_period,
                                   dt_sim=self.dt)
        for q in particle_data:
            if q not in _COMPONENTS:
                raise ValueError("Invalid string in particle_data: %s" % q)
        self.species_dict, self.select, self.subsampling_fraction = dict(species), select, subsampling_fraction
        self.particle_data = [q for q in particle_data if q not in ('id', 'charge')]
        # E, B at the particles only exist as arrays after the unfused gather (the fused gather + push keeps them
        # in registers): Simulation.step looks at this flag
        self.needs_gathered_fields = ('E' in self.particle_data) or ('B' in self.particle_data)

    def _records(self, species):
        """tracked species get their ids written as well, ionizable ones their per-particle charge
        (particle_diag.py:111-128)"""
        return self.particle_data + (['id'] if species.tracker is not None else []) \
            + (['charge'] if species.ionizer is not None else [])

    @staticmethod
    def _attr(sp, name):
        if name == 'gamma':
            return 1. / _host(sp.inv_gamma)
        if name == 'id':
            return _host(sp.tracker.id)
        if name == 'charge':
            from scipy.constants import e
            return e * _host(sp.ionizer.levels.id).astype(np.float64)
        return _host(getattr(sp, name))

    def apply_selection(self, species):
        """particle_diag.py:366-410"""
        keep = np.ones(species.Ntot, dtype=bool)
        if self.subsampling_fraction is not None:
            keep &= np.random.rand(species.Ntot) < self.subsampling_fraction
        if self.select is not None:
            for q, (lo, hi) in self.select.items():
                v = self._attr(species, q)[:species.Ntot]
                if lo is not None:
                    keep &= (v > lo)
                if hi is not None:
                    keep &= (v < hi)
        return keep

    def setup_openpmd_species_group(self, grp, species):
        """particle_diag.py:129-175: attributes of the species, the constant records mass / charge, positionOffset"""
        grp.attrs['particleShape'] = 1.
        for key, value in (('currentDeposition', 'directMorseNielson'), ('particleSmoothing', 'none'),
                           ('particlePush', 'Vay'), ('particleInterpolation', 'uniform')):
            grp.attrs[key] = _text(value)
        one = np.array([1], dtype=np.uint64)
        for record, value in (('mass', species.m), ('charge', species.q)):
            if record == 'charge' and species.ionizer is not None:
                continue                      # per-particle record instead (particle_diag.py:124-128)
            node = grp.require_group(record)
            self.setup_openpmd_species_record(node, record)
            self.setup_openpmd_component(node)
            node.attrs['shape'], node.attrs['value'] = one, value
        self.setup_openpmd_species_record(grp.require_group('positionOffset'), 'positionOffset')
        for co in 'xyz':
            node = grp.require_group('positionOffset/' + co)
            self.setup_openpmd_component(node)
            node.attrs['shape'], node.attrs['value'] = one, 0.

    def setup_openpmd_species_record(self, node, record):
        self.setup_openpmd_record(node, record)
        node.attrs['macroWeighted'] = np.uint32(_WEIGHTING[record][0])
        node.attrs['weightingPower'] = _WEIGHTING[record][1]

    def write_hdf5(self, iteration):
        f = None
        if self.rank == 0:
            f = open_file(self.file_stem(iteration), 'a')
            self.setup_openpmd_file(f, iteration, iteration * self.dt, self.dt)
        for name in self.species_names_list:
            species = self.species_dict[name]
            if species is None:
                continue
            grp = None
            if f is not None:
                grp = f.require_group('/data/%d/particles/%s' % (iteration, name))
                self.setup_openpmd_species_group(grp, species)
            keep = self.apply_selection(species)
            for record in self._records(species):
                for comp in _COMPONENTS[record]:
                    data = self._attr(species, comp)[:species.Ntot][keep]
                    if record == 'momentum' and species.m > 0:        # kg m/s (particle_diag.py:497-500)
                        data = data * (species.m * c)
                    if self.comm is not None and self.comm.size > 1:
                        data = self.comm.gather_ptcl_array(data)
                    if grp is None:
                        continue
                    path = record if len(_COMPONENTS[record]) == 1 else '%s/%s' % (record, comp[-1])
                    if path in grp:       # left over from an earlier run with another particle number
                        del grp[path]
                    dset = grp.create_dataset(path, (len(data),), dtype=('uint64' if comp == 'id' else 'f8'))
                    self.setup_openpmd_component(dset)
                    dset[:] = data
                if grp is not None:
                    self.setup_openpmd_species_record(grp[record], record)
        if f is not None:
            f.close()


# ---------------------------------------------------------------------------
# lab-frame output of a boosted-frame simulation (boosted_field_diag.py:26-823)
# ---------------------------------------------------------------------------
SLICE_FIELDS = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')


class LabSnapshot(object):
    """One lab-frame time: its file, its z range and the slices collected since the last flush
    (boosted_field_diag.py:375-527)."""

    def __init__(self, t_lab, zmin_lab, zmax_lab, write_dir, i, fld, Nr_output):
        self.stem = os.path.join(write_dir, 'hdf5', 'data%08d' % i)
        self.iteration, self.t_lab, self.zmin_lab, self.zmax_lab = i, t_lab, zmin_lab, zmax_lab
        self.current_z_lab = self.current_z_boost = 0.
        self.buffered_slices, self.buffer_z_indices = [], []
        self.shape = (len(SLICE_FIELDS), 2 * fld.Nm - 1, Nr_output)
        self.d_slice = None                      # device buffer of one slice, allocated at first use

    def update_current_output_positions(self, t_boost, inv_gamma, inv_beta):
        """where the plane t = t_lab is at the boosted-frame time t_boost (Lorentz transformation at fixed t_lab)"""
        self.current_z_boost = (self.t_lab * inv_gamma - t_boost) * c * inv_beta
        self.current_z_lab = (self.t_lab - t_boost * inv_gamma) * c * inv_beta

    def register_slice(self, slice_array, inv_dz_lab):
        """successive slices move one lab cell to the left: integer bookkeeping after the first one"""
        if not self.buffer_z_indices:
            iz_lab = int((self.current_z_lab - self.zmin_lab) * inv_dz_lab)
        else:
            iz_lab = self.buffer_z_indices[-1] - 1
        self.buffer_z_indices.append(iz_lab)
        self.buffered_slices.append(slice_array)

    def compact_slices(self):
        """([10, 2 Nm - 1, Nr, n] array in increasing z, iz_min, iz_max) or (None, None, None)"""
        if not self.buffer_z_indices:
            return None, None, None
        if np.any(np.diff(self.buffer_z_indices) != -1):
            raise UserWarning('In the boosted frame diagnostic, the buffered slices are not contiguous in z.\n'
                              'The boosted frame diagnostics may be inaccurate.')
        return (np.stack(self.buffered_slices[::-1], axis=-1), self.buffer_z_indices[-1],
                self.buffer_z_indices[0] + 1)


class SliceHandler(object):
    """Extraction of one z slice of all grid fields and the Lorentz transformation (boosted_field_diag.py:529-742)."""

    def __init__(self, gamma_boost, beta_boost, Nr_output):
        self.gamma_boost, self.beta_boost, self.Nr_output = gamma_boost, beta_boost, Nr_output
        self.field_to_index = {name: k for k, name in enumerate(SLICE_FIELDS)}

    def slice_index(self, fld, comm, z_boost, zmin_boost):
        """lower row and its weight for the cell-centred interpolation at z_boost (boosted_field_diag.py:577-586)"""
        dz = fld.interp[0].dz
        z_cell = (z_boost - zmin_boost - 0.5 * dz) / dz
        iz = int(z_cell)
        Sz = iz + 1 - z_cell
        if comm is not None:
            iz += comm.n_guard
            if comm.left_proc is None:
                iz += comm.nz_damp + comm.n_inject
        return iz, Sz

    def extract_slice(self, fld, comm, z_boost, zmin_boost, snapshot):
        """[10, 2 Nm - 1, Nr] host array of the boosted-frame fields at z_boost: one small kernel per mode on the
        device-resident grids and ONE read-back of the slice (b2_extract_slice)."""
        iz, Sz = self.slice_index(fld, comm, z_boost, zmin_boost)
        g0 = fld.interp[0]
        if not isinstance(g0.Er, DeviceArray):          # between step() calls the grids are host arrays
            out = np.empty(snapshot.shape)
            for k, name in enumerate(SLICE_FIELDS):
                for m in range(fld.Nm):
                    a = getattr(fld.interp[m], name)
                    row = (Sz * a[iz, :self.Nr_output]) + ((1. - Sz) * a[iz + 1, :self.Nr_output])
                    if m == 0:
                        out[k, 0] = row.real
                    else:
                        out[k, 2 * m - 1], out[k, 2 * m] = 2 * row.real, 2 * row.imag
            return out
        if snapshot.d_slice is None:
            snapshot.d_slice = DeviceArray(int(np.prod(snapshot.shape)), np.float64)
        for m in range(fld.Nm):
            grids = _lib.ptr_array([getattr(fld.interp[m], name) for name in SLICE_FIELDS])
            call.b2_extract_slice(_lib.context().handle, grids, m, fld.Nm, g0.Nz, g0.Nr, self.Nr_output, iz, Sz,
                                  snapshot.d_slice.ptr, None)
        return snapshot.d_slice.get().reshape(snapshot.shape)

    def transform_fields_to_lab_frame(self, fields):
        """In place, boost of velocity -beta c along z of [10, ...] packed fields (boosted_field_diag.py:686-742):
        E_perp' = gamma (E - c beta x B), B_perp' = gamma (B + beta x E / c), (c rho, Jz) as a four-vector."""
        g, cb, b_c = self.gamma_boost, c * self.beta_boost, self.beta_boost / c
        i = self.field_to_index
        for a, b, sign in (('Er', 'Bt', 1.), ('Et', 'Br', -1.)):
            ea, bb = fields[i[a]].copy(), fields[i[b]].copy()
            fields[i[a]] = g * (ea + sign * cb * bb)
            fields[i[b]] = g * (bb + sign * b_c * ea)
        rho, jz = fields[i['rho']].copy(), fields[i['Jz']].copy()
        fields[i['rho']] = g * (rho + b_c * jz)
        fields[i['Jz']] = g * (jz + cb * rho)


class BackTransformedFieldDiagnostic(FieldDiagnostic):
    """Fields *in the lab frame* from a simulation in the boosted frame, as a series of snapshots at fixed lab
    times inside a virtual window [zmin_lab, zmax_lab] + v_lab t (boosted_field_diag.py:26-373): every cycle each
    snapshot receives the z slice that its plane t_lab = const crosses; slices are buffered on the host and written
    every `period` cycles."""

    def __init__(self, zmin_lab, zmax_lab, v_lab, dt_snapshots_lab, Ntot_snapshots_lab, gamma_boost, period,
                 fldobject, comm=None, fieldtypes=["E", "B"], write_dir=None, t_min_snapshots_lab=0.,
                 t_max_snapshots_lab=np.inf):
        if write_dir is None:
            write_dir = 'lab_diags'
        FieldDiagnostic.__init__(self, period, fldobject, comm, fieldtypes, write_dir)
        self.gamma_boost = gamma_boost
        self.inv_gamma_boost = 1. / gamma_boost
        self.beta_boost = np.sqrt(1. - self.inv_gamma_boost**2)
        self.inv_beta_boost = 1. / self.beta_boost
        # one boosted-frame cycle moves the plane of a snapshot by dz_lab in the lab frame
        dz_lab = c * self.fld.dt * self.inv_beta_boost * self.inv_gamma_boost
        Nz = int((zmax_lab - zmin_lab) / dz_lab) + 1
        self.inv_dz_lab = 1. / dz_lab
        Nr = self.fld.interp[0].Nr if comm is None else comm.get_Nr(with_damp=False)
        self.snapshots = []
        for i in range(Ntot_snapshots_lab):
            t_lab = i * dt_snapshots_lab
            if t_min_snapshots_lab <= t_lab < t_max_snapshots_lab:
                snapshot = LabSnapshot(t_lab, zmin_lab + v_lab * t_lab, zmax_lab + v_lab * t_lab, self.write_dir, i,
                                       self.fld, Nr)
                self.snapshots.append(snapshot)
                self.create_file_empty_meshes(snapshot.stem, i, t_lab, Nr, Nz, snapshot.zmin_lab, dz_lab,
                                              self.fld.dt)
        self.slice_handler = SliceHandler(self.gamma_boost, self.beta_boost, Nr)

    def write(self, iteration):
        self.store_snapshot_slices(iteration)
        if iteration % self.period == 0:
            self.flush_to_disk()

    def store_snapshot_slices(self, iteration):
        fld, comm = self.fld, self.comm
        if "rho" in self.fieldtypes or "J" in self.fieldtypes:
            fld.spect2interp('rho_prev')          # rho at time n
            fld.spect2interp('J')
            if comm is not None and comm.size > 1:
                if not fld.exchanged_source['J']:
                    comm.exchange_fields(fld.interp, 'J', 'add')
                if not fld.exchanged_source['rho_prev']:
                    comm.exchange_fields(fld.interp, 'rho', 'add')
        if comm is None:
            zmin_boost, zmax_boost = fld.interp[0].zmin, fld.interp[0].zmax
        else:
            zmin_boost, zmax_boost = comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=self.rank)
        time = iteration * fld.dt
        for snapshot in self.snapshots:
            snapshot.update_current_output_positions(time, self.inv_gamma_boost, self.inv_beta_boost)
            if (zmin_boost < snapshot.current_z_boost < zmax_boost) and \
                    (snapshot.zmin_lab < snapshot.current_z_lab < snapshot.zmax_lab):
                data = self.slice_handler.extract_slice(fld, comm, snapshot.current_z_boost, zmin_boost, snapshot)
                snapshot.register_slice(data, self.inv_dz_lab)

    def flush_to_disk(self):
        for snapshot in self.snapshots:
            field_array, iz_min, iz_max = snapshot.compact_slices()
            if field_array is not None:
                self.slice_handler.transform_fields_to_lab_frame(field_array)
            if self.comm is not None and self.comm.size > 1:
                field_array, iz_min, iz_max = self.gather_slices(field_array, iz_min, iz_max)
            if self.rank == 0 and field_array is not None:
                self.write_slices(field_array, iz_min, iz_max, snapshot, self.slice_handler.field_to_index)
            snapshot.buffered_slices, snapshot.buffer_z_indices = [], []

    def gather_slices(self, field_array, iz_min, iz_max):
        """Stitch the slices of the ranks together on rank 0 (boosted_field_diag.py:236-311)."""
        parts = [None] * self.comm.size
        self.comm._host_group().all_gather_object(parts, (field_array, iz_min, iz_max))
        parts = [p for p in parts if p[0] is not None]
        if self.rank != 0 or not parts:
            return None, None, None
        lo, hi = min(p[1] for p in parts), max(p[2] for p in parts)
        out = np.zeros(parts[0][0].shape[:3] + (hi - lo,))
        for a, a_lo, a_hi in parts:
            out[..., a_lo - lo:a_hi - lo] = a
        return out, lo, hi

    def write_slices(self, field_array, iz_min, iz_max, snapshot, f2i):
        f = self.open_file(snapshot.stem)
        grp = f['/data/%d/fields/' % snapshot.iteration]
        for path, quantity in self._components():
            grp[path][:, :, iz_min:iz_max] = field_array[f2i[quantity]]
        f.close()


BoostedFieldDiagnostic = BackTransformedFieldDiagnostic


class ParticleCatcher(object):
    """Particles that crossed the plane of a lab-frame snapshot during the last cycle, moved onto the plane and
    Lorentz-transformed (boosted_particle_diag.py:431-760).  The crossing test runs on the device over the resident
    particle arrays (b2_select_crossing); only the selected particles are gathered (b2_permute) and read back.  The
    reference downloads the whole slab of cells around the plane and selects on the host, which needs a sorted
    species."""
    ATTRS = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w', 'inv_gamma')

    def __init__(self, gamma_boost, beta_boost, fldobject):
        self.gamma_boost, self.beta_boost, self.fld, self.dt = gamma_boost, beta_boost, fldobject, fldobject.dt
        self._idx = self._count = self._packed = None

    def extract_slice(self, species, current_z_boost, previous_z_boost, t, select=None):
        data = self.get_particle_slice(species, current_z_boost, previous_z_boost)
        data = self.interpolate_particles_to_lab_frame(data, current_z_boost, t)
        if select is not None:
            data = self.apply_selection(select, data)
        if species.m > 0:                          # openPMD: momenta in kg m/s
            for k in ('ux', 'uy', 'uz'):
                data[k] = data[k] * (species.m * c)
        if 'charge' in data:                       # ionization level -> charge in C
            from scipy.constants import e
            data['charge'] = data['charge'] * e
        return data

    def get_particle_slice(self, species, z_curr, z_prev):
        """{'x', ..., 'inv_gamma' (, 'id')} of the particles with z on one side of the plane now and on the other
        side one cycle ago"""
        n = species.Ntot
        if not isinstance(species.z, DeviceArray):            # between step() calls: host arrays
            z, uz, ig = (np.asarray(getattr(species, k))[:n] for k in ('z', 'uz', 'inv_gamma'))
            z_old = z - uz * ig * c * self.dt
            keep = np.flatnonzero(((z >= z_curr) & (z_old <= z_prev)) | ((z <= z_curr) & (z_old >= z_prev)))
            out = {k: np.take(np.asarray(getattr(species, k))[:n], keep) for k in self.ATTRS}
            if species.tracker is not None:
                out['id'] = np.take(np.asarray(species.tracker.id)[:n], keep)
            if species.ionizer is not None:
                out['charge'] = np.take(np.asarray(species.ionizer.levels.id)[:n], keep)
            return out
        import ctypes
        ctx = _lib.context().handle
        if self._count is None:
            self._count = DeviceArray(1, np.int64)
        cap = max(4096, n // 32) if self._idx is None else self._idx.size
        found = ctypes.c_int64(0)
        while True:
            if self._idx is None or self._idx.size < cap:
                self._idx = DeviceArray(cap, np.int64)
            call.b2_select_crossing(ctx, n, species.z.ptr, species.uz.ptr, species.inv_gamma.ptr, c, self.dt, z_curr,
                                    z_prev, self._idx.size, self._idx.ptr, self._count.ptr, ctypes.byref(found), None)
            if found.value <= self._idx.size:
                break
            cap = int(found.value)                 # the buffer was too small: the count is exact, run again
        k = int(found.value)
        tracked = species.tracker is not None
        extra = ([('id', species.tracker.id)] if tracked else []) + \
            ([('charge', species.ionizer.levels.id)] if species.ionizer is not None else [])
        if k == 0:
            out = {name: np.zeros(0) for name in self.ATTRS}
            out.update({name: np.zeros(0, dtype=np.uint64) for name, _ in extra})
            return out
        idx = self._idx.view((k,))
        idx.set(np.sort(idx.get()))                # particle-array order, whatever order the atomics produced
        rows = len(self.ATTRS) + len(extra)
        if self._packed is None or self._packed.size < rows * k:
            self._packed = DeviceArray(rows * k, np.float64)
        src = [getattr(species, name) for name in self.ATTRS] + [a for _, a in extra]
        dst = [self._packed.ptr + 8 * k * r for r in range(rows)]
        call.b2_permute(ctx, k, idx.ptr, rows, _lib.ptr_array(src), _lib.ptr_array(dst), None)
        packed = self._packed.view((rows, k)).get()
        out = {name: packed[r] for r, name in enumerate(self.ATTRS)}
        for j, (name, _) in enumerate(extra):
            out[name] = packed[len(self.ATTRS) + j].view(np.uint64)
        return out

    def interpolate_particles_to_lab_frame(self, d, current_z_boost, t):
        """Move every particle along its straight path to the time it met the plane (which travels at -c / beta in
        the boosted frame), then z and uz to the lab frame (boosted_particle_diag.py:631-686)."""
        ig = d.pop('inv_gamma')
        v_z, v_plane = d['uz'] * ig * c, -c / self.beta_boost
        t_cross = t - (current_z_boost - d['z']) / (v_plane - v_z)
        shift = c * (t_cross - t) * ig
        d['x'] = d['x'] + shift * d['ux']
        d['y'] = d['y'] + shift * d['uy']
        z = d['z'] + shift * d['uz']
        d['z'] = self.gamma_boost * (z + self.beta_boost * c * t_cross)
        d['uz'] = self.gamma_boost * d['uz'] + (1. / ig) * (self.beta_boost * self.gamma_boost)
        return d

    @staticmethod
    def apply_selection(select, d):
        keep = np.ones(len(d['w']), dtype=bool)
        for q, (lo, hi) in select.items():
            if lo is not None:
                keep &= d[q] > lo
            if hi is not None:
                keep &= d[q] < hi
        return {k: v[keep] for k, v in d.items()}


class BackTransformedParticleDiagnostic(ParticleDiagnostic):
    """Particles *in the lab frame* from a simulation in the boosted frame (boosted_particle_diag.py:26-429): every
    cycle each snapshot collects the particles that its plane t_lab = const swept over; they are appended to the
    datasets of the snapshot every `period` cycles.  "E", "B" and "gamma" are not available (as in the reference)."""
    needs_gathered_fields = False

    def __init__(self, zmin_lab, zmax_lab, v_lab, dt_snapshots_lab, Ntot_snapshots_lab, gamma_boost, period,
                 fldobject, particle_data=["position", "momentum", "weighting"], select=None, write_dir=None,
                 species={"electrons": None}, comm=None, t_min_snapshots_lab=0., t_max_snapshots_lab=np.inf):
        if write_dir is None:
            write_dir = 'lab_diags'
        for q in particle_data:
            if q not in ('position', 'momentum', 'weighting'):
                raise ValueError("Invalid quantity for particle output: %s" % q)
        ParticleDiagnostic.__init__(self, period, species, comm, particle_data, select, write_dir)
        self.needs_gathered_fields = False
        self.fld = fldobject
        self.gamma_boost = gamma_boost
        self.inv_gamma_boost = 1. / gamma_boost
        self.beta_boost = np.sqrt(1. - self.inv_gamma_boost**2)
        self.inv_beta_boost = 1. / self.beta_boost
        self.snapshots = []
        for i in range(Ntot_snapshots_lab):
            t_lab = i * dt_snapshots_lab
            if t_min_snapshots_lab <= t_lab < t_max_snapshots_lab:
                snap = LabSnapshot(t_lab, zmin_lab + v_lab * t_lab, zmax_lab + v_lab * t_lab, self.write_dir, i,
                                   fldobject, fldobject.interp[0].Nr)
                snap.buffered_particles = {name: [] for name in self.species_names_list}
                self.snapshots.append(snap)
                self.create_file_empty_slice(snap.stem, i, t_lab, self.dt)
        self.particle_catcher = ParticleCatcher(self.gamma_boost, self.beta_boost, self.fld)

    def _paths(self, species):
        """(key of the caught data, dataset path, dtype) of every array written for this species"""
        out = []
        for record in self._records(species):
            for comp in _COMPONENTS[record]:
                path = record if len(_COMPONENTS[record]) == 1 else '%s/%s' % (record, comp[-1])
                out.append((comp, path, 'uint64' if comp == 'id' else 'f8'))
        return out

    def create_file_empty_slice(self, stem, iteration, time, dt):
        f = self.open_file(stem)
        if f is None:
            return
        self.setup_openpmd_file(f, iteration, time, dt)
        for name in self.species_names_list:
            species = self.species_dict[name]
            grp = f.require_group('/data/%d/particles/%s/' % (iteration, name))
            self.setup_openpmd_species_group(grp, species)
            for comp, path, dtype in self._paths(species):
                self.setup_openpmd_component(grp.require_dataset(path, (0,), maxshape=(None,), dtype=dtype))
            for record in self._records(species):
                self.setup_openpmd_species_record(grp[record], record)
        f.close()

    def write(self, iteration):
        self.store_snapshot_slices(iteration)
        if iteration % self.period == 0:
            self.flush_to_disk()

    def store_snapshot_slices(self, iteration):
        g0 = self.fld.interp[0]
        time = iteration * self.dt
        for snap in self.snapshots:
            snap.update_current_output_positions(time, self.inv_gamma_boost, self.inv_beta_boost)
            prev_z_boost = (snap.t_lab * self.inv_gamma_boost - (time - self.dt)) * c * self.inv_beta_boost
            if (g0.zmin <= snap.current_z_boost < g0.zmax) and (snap.zmin_lab <= snap.current_z_lab < snap.zmax_lab):
                for name in self.species_names_list:
                    snap.buffered_particles[name].append(self.particle_catcher.extract_slice(
                        self.species_dict[name], snap.current_z_boost, prev_z_boost, time, self.select))

    def flush_to_disk(self):
        multi = self.comm is not None and self.comm.size > 1
        for snap in self.snapshots:
            for name in self.species_names_list:
                species = self.species_dict[name]
                caught = snap.buffered_particles[name]
                data = {}
                for comp, path, dtype in self._paths(species):
                    a = np.concatenate([d[comp] for d in caught]) if caught else np.zeros(0, dtype=dtype)
                    data[path] = self.comm.gather_ptcl_array(a) if multi else a
                if self.rank == 0:
                    f = self.open_file(snap.stem)
                    grp = f['/data/%d/particles/%s' % (snap.iteration, name)]
                    for path, a in data.items():
                        dset = grp[path]
                        start = dset.shape[0]
                        dset.resize(start + len(a), axis=0)
                        dset[start:] = a
                    f.close()
                snap.buffered_particles[name] = []


BoostedParticleDiagnostic = BackTransformedParticleDiagnostic


# ---------------------------------------------------------------------------
# reading a diagnostic back
# ---------------------------------------------------------------------------
def read_diag(write_dir, iteration):
    """The tree below `/data/<iteration>/` of one output file as a flat dictionary: 'fields/E/r' -> [2 Nm - 1, Nr, Nz]
    array, 'particles/<species>/position/x' -> array, '<path>@<attribute>' -> attribute, plus 'time', 'dt' and, when
    there are meshes, 'zmin', 'dz', 'dr' taken from the mesh attributes."""
    path = existing_file(os.path.join(os.path.abspath(write_dir), 'hdf5', 'data%08d' % iteration))
    if path is None:
        raise OSError('No diagnostic of iteration %d in %s' % (iteration, write_dir))
    base = '/data/%d' % iteration
    out = {}
    for key, value in read_tree(path).items():
        if key.startswith(base + '/'):
            out[key[len(base) + 1:]] = value
        elif key.startswith(base + '@'):
            out[key[len(base) + 1:]] = value
        elif key.startswith('/restart/%d/' % iteration):
            out['restart/' + key[len('/restart/%d/' % iteration):]] = value
    for key, value in list(out.items()):
        if key.endswith('@gridSpacing'):
            out['dr'], out['dz'] = float(value[0]), float(value[1])
        elif key.endswith('@gridGlobalOffset'):
            out['zmin'] = float(value[1])
    return out


def list_iterations(write_dir):
    d = os.path.join(os.path.abspath(write_dir), 'hdf5')
    if not os.path.isdir(d):
        return []
    return sorted({int(m.group(1)) for m in (re.match(r'data(\d+)\.(h5|npz)$', n) for n in os.listdir(d)) if m})


# ---------------------------------------------------------------------------
# checkpoint / restart (checkpoint_restart.py:22-330)
# ---------------------------------------------------------------------------
def set_periodic_checkpoint(sim, period, checkpoint_dir='./checkpoints'):
    """Every `period` cycles each rank saves its E, B (+ PML components) with guard cells and all its particles
    (checkpoint_restart.py:22-75)."""
    comm = sim.comm
    if comm.rank == 0:
        os.makedirs(checkpoint_dir, exist_ok=True)
    comm.barrier()
    write_dir = os.path.join(checkpoint_dir, 'proc%d/' % comm.rank)
    fieldtypes = ["E", "B"]
    if sim.use_pml:
        fieldtypes += ["Er_pml", "Et_pml", "Br_pml", "Bt_pml"]
    sim.checkpoints.append(FieldDiagnostic(period, sim.fld, fieldtypes=fieldtypes, write_dir=write_dir,
                                           keep_mode0_imag=True))
    species = {'species %d' % i: sp for i, sp in enumerate(sim.ptcl)}
    if species:
        sim.checkpoints.append(ParticleDiagnostic(period, species, write_dir=write_dir))


def check_restart(sim, iteration, checkpoint_dir):
    """checkpoint_restart.py:191-218"""
    if not os.path.exists(checkpoint_dir):
        raise RuntimeError('The directory %s, which is required to restart a simulation, does not exist.'
                           % checkpoint_dir)
    nproc = sum(1 for d in os.listdir(checkpoint_dir) if re.match(r'proc\d+', d))
    if nproc != sim.comm.size:
        raise RuntimeError('For a valid restart, the current simulation should use %d MPI processes.' % nproc)
    if sim.comm.moving_win is not None:
        raise RuntimeError('The moving window has already been initialized.\nFor valid restart, the moving window '
                           'should be initialized *after*\ncalling `restart_from_checkpoint`.')


def restart_from_checkpoint(sim, iteration=None, checkpoint_dir='./checkpoints'):
    """Load E, B (+ PML components), the particles, the iteration / time and the position of the (moving) grid
    from a checkpoint written by a simulation with the same set-up and number of ranks
    (checkpoint_restart.py:77-189; the reference reads the files back through openPMD-viewer, here the tree is read
    directly).  Call before `step()` and before `set_moving_window` (host copy of the data)."""
    from .particles import FIELD_ATTRS
    comm = sim.comm
    check_restart(sim, iteration, checkpoint_dir)
    data_dir = os.path.join(checkpoint_dir, 'proc%d' % comm.rank)
    its = list_iterations(data_dir)
    if not its:
        raise RuntimeError('The directory %s, which is required to restart a simulation from checkpoints, '
                           'holds no checkpoint.' % data_dir)
    if iteration is None:
        iteration = its[-1]
    elif iteration not in its:
        raise RuntimeError('The iteration %d is not among the checkpoints (%s).' % (iteration, its))
    d = read_diag(data_dir, iteration)
    g0 = sim.fld.interp[0]
    if d['fields/E/r'].shape != (2 * sim.fld.Nm - 1, g0.Nr, g0.Nz):
        raise RuntimeError('The checkpoint was written with a different grid: the local grid (with guard cells) '
                           'must be identical, which also requires the same number of ranks.')
    species_names = sorted({k.split('/')[1].split('@')[0] for k in d if k.startswith('particles/')})
    if len(species_names) != len(sim.ptcl):
        raise RuntimeError('Species numbers in checkpoint and simulation should be same, but got %d and %d. '
                           'Use add_new_species method to add species to simulation or sim.ptcl = [] to remove '
                           'them' % (len(species_names), len(sim.ptcl)))
    sim.iteration, sim.time = iteration, float(d['time'])
    names = ['Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'] + (['Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml'] if sim.use_pml else [])
    for m in range(sim.fld.Nm):
        for k in names:
            data = d['fields/' + (k if k.endswith('_pml') else '%s/%s' % (k[0], k[1]))]
            if m == 0:
                a = data[0] + (1.j * d['restart/' + k] if ('restart/' + k) in d else 0.j)
            else:
                a = 0.5 * (data[2 * m - 1] + 1.j * data[2 * m])
            getattr(sim.fld.interp[m], k)[:, :] = a.T
    zmin_old, zmin_new = g0.zmin, d['zmin']
    for m in range(sim.fld.Nm):
        length = sim.fld.interp[m].zmax - sim.fld.interp[m].zmin
        sim.fld.interp[m].zmin, sim.fld.interp[m].zmax = zmin_new, zmin_new + length
    comm.shift_global_domain_positions(zmin_new - zmin_old)
    for i, sp in enumerate(sim.ptcl):
        grp = 'particles/species %d/' % i
        for attr, key in (('x', 'position/x'), ('y', 'position/y'), ('z', 'position/z'), ('w', 'weighting')):
            setattr(sp, attr, np.ascontiguousarray(d[grp + key], dtype=np.float64))
        to_u = 1. / (sp.m * c) if sp.m > 0 else 1.
        for attr, key in (('ux', 'momentum/x'), ('uy', 'momentum/y'), ('uz', 'momentum/z')):
            setattr(sp, attr, np.ascontiguousarray(d[grp + key], dtype=np.float64) * to_u)
        sp.Ntot = len(sp.x)
        if (grp + 'id') in d:
            if sp.tracker is None:
                sp.track(comm)
            sp.tracker.overwrite_ids(d[grp + 'id'], comm)
        sp.inv_gamma = 1. / np.sqrt(1 + sp.ux**2 + sp.uy**2 + sp.uz**2)
        if sp.ionizer is not None:            # checkpoint_restart.py:317-323
            from scipy.constants import e
            sp.ionizer.levels.overwrite_ids(np.uint64(np.round(d[grp + 'charge'] / e)))
            sp.ionizer.w_times_level = sp.w * sp.ionizer.levels.id
        for k in FIELD_ATTRS:
            setattr(sp, k, np.zeros(sp.Ntot))
        if hasattr(sp.injector, 'reset_injection_positions'):
            sp.injector.reset_injection_positions()

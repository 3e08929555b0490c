"""
`Particles`: the species container of the operator surface, backed by the B200
kernels of libfbpic_b200.so.  Mirrors fbpic/particles/particles.py:52-1094 for
the hot path (gather, push_p, push_x, deposit, sort_particles,
rearrange_particle_arrays, send/receive_particles_*): same method names,
argument meaning and attribute names (`x..w`, `Ex..Bz`, `cell_idx`,
`sorted_idx`, `prefix_sum`, `sorted`, `Ntot`, `q`, `m`).

Beyond SURVEY section 8 (2g/2f), hooked in here: ionization (`make_ionizable`), Compton scattering
(`activate_compton`), tracking (`track`)  -- fbpic_b200/ionization.py, compton.py.
"""
import inspect
import warnings
import numpy as np

from . import _lib
from ._lib import DeviceArray, call, ptr_array

FLOAT_ATTRS = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'inv_gamma', 'w')
FIELD_ATTRS = ('Ex', 'Ey', 'Ez', 'Bx', 'By', 'Bz')


def _dens_func_args(dens_func):
    args = inspect.getfullargspec(dens_func).args
    if args and args[0] == 'self':
        args = args[1:]
    if args not in (['x', 'y', 'z'], ['z', 'r']):
        raise ValueError("The argument `dens_func` needs to be a function of z, r\n"
                         "or a function of x, y, z.")
    return args


def generate_evenly_spaced(Npz, zmin, zmax, Npr, rmin, rmax, Nptheta, n, dens_func,
                           ux_m, uy_m, uz_m, ux_th, uy_th, uz_th):
    """Host-side particle loader: regular (z, r, theta) lattice, one random azimuthal
    offset per (z, r) position, weights n*r*dtheta*dr*dz.  Restates
    fbpic/particles/injection/continuous_injection.py:203-275 (same draw order from
    np.random, so the same seed gives the same particles as the reference)."""
    if Npz * Npr * Nptheta <= 0:
        e = np.empty(0)
        return 0, e, e.copy(), e.copy(), e.copy(), e.copy(), e.copy(), e.copy(), e.copy()
    dz = (zmax - zmin) * 1. / Npz
    dr = (rmax - rmin) * 1. / Npr
    dtheta = 2 * np.pi / Nptheta
    z_reg = zmin + dz * (np.arange(Npz) + 0.5)
    r_reg = rmin + dr * (np.arange(Npr) + 0.5)
    theta_reg = dtheta * np.arange(Nptheta)
    zp, rp, thetap = np.meshgrid(z_reg, r_reg, theta_reg, copy=True, indexing='ij')
    thetap[:, :, :] = thetap + (2 * np.pi * np.random.rand(Npz, Npr))[:, :, np.newaxis]
    r = rp.flatten()
    x = r * np.cos(thetap.flatten())
    y = r * np.sin(thetap.flatten())
    z = zp.flatten()
    w = n * r * dtheta * dr * dz
    if dens_func is not None:
        if _dens_func_args(dens_func) == ['x', 'y', 'z']:
            w *= dens_func(x=x, y=y, z=z)
        else:
            w *= dens_func(z=z, r=r)
    if np.any(w < 0):
        warnings.warn('The specified particle density returned negative densities.\n'
                      'No particles were generated in areas of negative density.')
    keep = (w > 0)
    Ntot = int(keep.sum())
    x, y, z, w = x[keep], y[keep], z[keep], w[keep]
    uz = uz_m * np.ones(Ntot) + uz_th * np.random.normal(size=Ntot)
    ux = ux_m * np.ones(Ntot) + ux_th * np.random.normal(size=Ntot)
    uy = uy_m * np.ones(Ntot) + uy_th * np.random.normal(size=Ntot)
    inv_gamma = 1. / np.sqrt(1 + ux**2 + uy**2 + uz**2)
    return Ntot, x, y, z, ux, uy, uz, inv_gamma, w


class ContinuousInjector(object):
    """Book-keeping of the plasma that enters through the right edge of a moving window
    (fbpic/particles/injection/continuous_injection.py:13-200); host side, runs once per particle
    exchange."""

    def __init__(self, Npz, zmin, zmax, dz_particles, Npr, rmin, rmax, Nptheta, n, dens_func,
                 ux_m, uy_m, uz_m, ux_th, uy_th, uz_th):
        self.Npr, self.rmin, self.rmax, self.Nptheta, self.n = Npr, rmin, rmax, Nptheta, n
        self.dens_func = dens_func
        self.ux_m, self.uy_m, self.uz_m = ux_m, uy_m, uz_m
        self.ux_th, self.uy_th, self.uz_th = ux_th, uy_th, uz_th
        self.dz_particles = (zmax - zmin) / Npz if Npz != 0 else dz_particles
        c_light = 299792458.
        self.v_end_plasma = c_light * uz_m / np.sqrt(1 + ux_m**2 + uy_m**2 + uz_m**2)
        self.nz_inject = self.z_inject = self.z_end_plasma = None

    def initialize_injection_positions(self, comm, v_moving_window, species_z, dt):
        if comm.rank != comm.size - 1 or self.z_inject is not None:
            return
        _, zmax_damp = comm.get_zmin_zmax(local=False, with_damp=True, with_guard=False)
        self.z_inject = zmax_damp + (3 - comm.n_inject) * comm.dz \
            + comm.exchange_period * dt * (v_moving_window - self.v_end_plasma)
        self.nz_inject = 0
        if len(species_z) > 0:
            self.z_end_plasma = species_z.max() + 0.5 * self.dz_particles
        else:
            _, self.z_end_plasma = comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
        if self.dz_particles is None:
            raise ValueError('The simulation uses continuous injection of particles, but was unable to\n'
                             'calculate the spacing between particles: pass `dz_particles`.')

    def reset_injection_positions(self):
        self.nz_inject = self.z_inject = self.z_end_plasma = None

    def increment_injection_positions(self, v_moving_window, duration):
        self.z_inject += v_moving_window * duration
        self.z_end_plasma += self.v_end_plasma * duration
        nz_new = int((self.z_inject - self.z_end_plasma) / self.dz_particles)
        self.nz_inject += nz_new
        self.z_end_plasma += nz_new * self.dz_particles

    def generate_particles(self, time):
        dens_func = None
        if self.dens_func is not None:
            user, v = self.dens_func, self.v_end_plasma
            if _dens_func_args(user) == ['z', 'r']:
                def dens_func(z, r):
                    return user(z - v * time, r)
            else:
                def dens_func(x, y, z):
                    return user(x, y, z - v * time)
        zmax = self.z_end_plasma
        zmin = self.z_end_plasma - self.nz_inject * self.dz_particles
        out = generate_evenly_spaced(self.nz_inject, zmin, zmax, self.Npr, self.rmin, self.rmax, self.Nptheta,
                                     self.n, dens_func, self.ux_m, self.uy_m, self.uz_m,
                                     self.ux_th, self.uy_th, self.uz_th)
        self.nz_inject = 0
        return out


class BallisticBeforePlane(object):
    """Injection "through a plane": the particles of the species move ballistically until they cross the plane
    z = z_plane_lab (fixed in the lab frame), e.g. the plasma entrance for a bunch initialised in vacuum in a
    boosted-frame run (fbpic/particles/injection/ballistic_before_plane.py:10-61)."""

    def __init__(self, z_plane_lab, boost):
        self.z_plane_lab = z_plane_lab
        self.inv_gamma_boost = 1. / boost.gamma0 if boost is not None else 1.
        self.beta_boost = boost.beta0 if boost is not None else 0.

    def get_current_plane_position(self, t):
        return self.inv_gamma_boost * self.z_plane_lab - self.beta_boost * 299792458. * t


class ParticleTracker(object):
    """Unique integer ids of the macroparticles, for tracking in post-processing
    (fbpic/particles/tracking/tracking.py:15-130): rank r hands out r, r + size, r + 2 size, ...
    On the device the ids are one more 8-byte array that follows the particles through the cell sort and the
    particle exchange (moved bit for bit by the same permutation / partition kernels as the float attributes)."""

    def __init__(self, comm_size, comm_rank, N):
        self.next_attributed_id = comm_rank
        self.id_step = comm_size
        self.id = self.generate_new_ids(N)
        self.spare = None

    def generate_new_ids(self, N):
        stop = self.next_attributed_id + N * self.id_step
        ids = np.arange(start=self.next_attributed_id, stop=stop, step=self.id_step, dtype=np.uint64)
        self.next_attributed_id = stop
        return ids

    def overwrite_ids(self, pid, comm):
        """tracking.py:92-118"""
        self.id = np.array(pid, dtype=np.uint64)
        local_max = int(pid.max()) if len(pid) > 0 else 0
        global_max = int(max(comm.allreduce_max([float(local_max)])))
        n = int((global_max - comm.rank) / self.id_step) + 1
        self.next_attributed_id = comm.rank + n * self.id_step

    def send_to_gpu(self, capacity):
        n = len(self.id)
        d = DeviceArray(capacity, np.uint64).view((n,))
        d.set(self.id)
        self.id = d
        self.spare = DeviceArray(capacity, np.uint64).view((n,))

    def receive_from_gpu(self):
        self.id = self.id.get()
        self.spare = None

    def swap(self, n=None):
        """the spare buffer (just filled by a permutation / partition) becomes the id array"""
        self.id, self.spare = self.spare, self.id
        if n is not None:
            self.id, self.spare = self.id.view((n,)), self.spare.view((n,))


class LevelCarrier(ParticleTracker):
    """The ionization level of every macroparticle of an ionizable species: one more 8-byte array that follows the
    particles through the cell sort and the particle exchange exactly like the tracked ids (same `id` / `spare` /
    `swap` plumbing); particles that enter (continuous injection) start at `level_start`
    (particles.py:370-372; particle_buffer_handling.py:120-172, 413-417)."""

    def __init__(self, level_start, N):
        self.level_start = int(level_start)
        self.id = self.generate_new_ids(N)
        self.spare = None

    def generate_new_ids(self, N):
        return np.full(N, self.level_start, dtype=np.uint64)

    def overwrite_ids(self, levels, comm=None):
        self.id = np.array(levels, dtype=np.uint64)


class Particles(object):
    """One species.  At the end/start of a PIC cycle the momenta are half a step
    behind the positions (particles.py:62-63)."""

    def __init__(self, q, m, n, Npz, zmin, zmax, Npr, rmin, rmax, Nptheta, dt,
                 ux_m=0., uy_m=0., uz_m=0., ux_th=0., uy_th=0., uz_th=0.,
                 dens_func=None, continuous_injection=True, grid_shape=None,
                 particle_shape='linear', use_cuda=True, dz_particles=None, is_tracer=False):
        if particle_shape not in ('linear', 'cubic'):
            raise ValueError("`particle_shape` should be either 'linear' or 'cubic' "
                             "but is `%s`" % particle_shape)
        self.use_cuda = True            # this build only has the GPU path
        self.data_is_on_gpu = False
        Ntot, x, y, z, ux, uy, uz, inv_gamma, w = generate_evenly_spaced(
            Npz, zmin, zmax, Npr, rmin, rmax, Nptheta, n, dens_func,
            ux_m, uy_m, uz_m, ux_th, uy_th, uz_th)
        self.Ntot, self.q, self.m, self.dt = Ntot, q, m, dt
        self.is_tracer = is_tracer
        self.x, self.y, self.z = x, y, z
        self.ux, self.uy, self.uz = ux, uy, uz
        self.inv_gamma, self.w = inv_gamma, w
        for k in FIELD_ATTRS:
            setattr(self, k, np.zeros(Ntot))
        self.continuous_injection = continuous_injection
        self.injector = ContinuousInjector(Npz, zmin, zmax, dz_particles, Npr, rmin, rmax, Nptheta, n,
                                           dens_func, ux_m, uy_m, uz_m, ux_th, uy_th, uz_th) \
            if continuous_injection else None
        self.tracker = None
        self.ionizer = None
        self.compton_scatterer = None
        self.n_integer_quantities = 0
        self.n_float_quantities = 8
        self.particle_shape = particle_shape
        self.keep_fields_sorted = False
        if grid_shape is None:
            raise ValueError("A `grid_shape` is needed when running on the GPU.\n"
                             "Please provide it when initializing particles.")
        self.grid_shape = grid_shape
        self.cell_idx = None
        self.sorted_idx = None
        self.prefix_sum = None
        self.sorting_buffers = None
        self.prefix_sum_shift = 0
        self.sorted = False
        # set by Simulation.step() when the fused gather+push is used: Ex..Bz are then neither uploaded nor
        # read back (they are never written on the device), which removes 6 of the 14 per-particle arrays from
        # the host<->device copies of a step() call
        self.fields_resident_only = False

    @property
    def ballistic_before_plane(self):
        return isinstance(self.injector, BallisticBeforePlane)

    def track(self, comm):
        """Activate particle tracking: a unique id per macroparticle, written by the particle diagnostics
        (fbpic/particles/particles.py:376-392)."""
        if self.data_is_on_gpu:
            raise _lib.B200Error('track() acts on the host copy of the particles: call it before step()')
        self.tracker = ParticleTracker(comm.size, comm.rank, self.Ntot)
        self.n_integer_quantities += 1

    # ------------------------------------------------------------------ device residency
    HEADROOM = 1.08     # per-particle device arrays are allocated with room for migration

    def _capacity_for(self, n):
        return int(n * self.HEADROOM) + 4096

    def _alloc_sort_arrays(self):
        """(Re)allocate the sort work arrays.  All per-particle arrays share one capacity so that
        particle exchange can reuse them without touching the allocator."""
        Nz, Nr = self.grid_shape
        cap = max(getattr(self, '_capacity', 0), self._capacity_for(self.Ntot))
        self._capacity = cap
        self.cell_idx = DeviceArray(cap, np.int32).view((self.Ntot,))
        self.sorted_idx = DeviceArray(cap, np.int64).view((self.Ntot,))
        if self.prefix_sum is None or self.prefix_sum.size != Nz * (Nr + 1):
            self.prefix_sum = DeviceArray(Nz * (Nr + 1), np.int32)
        # double buffers for the one-pass SoA permutation (8 state + 6 field arrays)
        self.sorting_buffers = [DeviceArray(cap, np.float64).view((self.Ntot,)) for _ in range(14)]
        self._order_matches_prefix = False
        self._keys_fresh = False
        self._j_since_sort = 0

    def send_particles_to_gpu(self):
        """particles.py:252-291"""
        if self.data_is_on_gpu:
            return
        self._capacity = self._capacity_for(self.Ntot)
        for k in FLOAT_ATTRS + FIELD_ATTRS:
            d = DeviceArray(self._capacity, np.float64).view((self.Ntot,))
            if k in FIELD_ATTRS and self.fields_resident_only:
                d.fill(0)       # gathered fields never leave the device in this mode: nothing to upload
            else:
                d.set(np.asarray(getattr(self, k), dtype=np.float64))
            setattr(self, k, d)
        for carrier in self.uint_carriers():
            carrier.send_to_gpu(self._capacity)
        if self.ionizer is not None:
            self.ionizer.send_to_gpu(self)
        self._alloc_sort_arrays()
        self.sorted = False
        self.data_is_on_gpu = True

    def receive_particles_from_gpu(self):
        """particles.py:293-333"""
        if not self.data_is_on_gpu:
            return
        for k in FLOAT_ATTRS:
            setattr(self, k, _lib.to_host(getattr(self, k)))
        for k in FIELD_ATTRS:
            # fused gather+push keeps the gathered fields in registers: the device arrays still hold the zeros
            # they were created with, so the host gets zeros without a copy
            setattr(self, k, np.zeros(self.Ntot) if self.fields_resident_only else _lib.to_host(getattr(self, k)))
        if self.ionizer is not None:
            self.ionizer.receive_from_gpu(self)
        for carrier in self.uint_carriers():
            carrier.receive_from_gpu()
        self.data_is_on_gpu = False

    def resize_device_arrays(self, new_arrays, n_new):
        """Adopt `new_arrays` (dict of the 8 state arrays, length n_new, built in the spare sort
        buffers) after a particle exchange; every other per-particle array is re-viewed at the new
        length.  Falls back to fresh allocations when n_new exceeds the capacity."""
        old = [getattr(self, k) for k in FLOAT_ATTRS]
        if n_new <= self._capacity:
            for k in FLOAT_ATTRS:
                setattr(self, k, new_arrays[k])
            spare = [a.view((n_new,)) for a in old] + [b.view((n_new,)) for b in self.sorting_buffers[8:]]
            self.sorting_buffers = spare
            for k in FIELD_ATTRS:
                setattr(self, k, getattr(self, k).view((n_new,)))
            self.cell_idx = self.cell_idx.view((n_new,))
            self.sorted_idx = self.sorted_idx.view((n_new,))
            self.Ntot = n_new
        else:
            self.Ntot = n_new
            self._capacity = self._capacity_for(n_new)
            for k in FLOAT_ATTRS:
                d = DeviceArray(self._capacity, np.float64).view((n_new,))
                d.copy_from(new_arrays[k])
                setattr(self, k, d)
            for k in FIELD_ATTRS:
                setattr(self, k, DeviceArray(self._capacity, np.float64).view((n_new,)))
            self._alloc_sort_arrays()
        self._order_matches_prefix = False
        self._keys_fresh = False
        self._j_since_sort = 0
        self.sorted = False

    def exchange_buffers(self, n_new):
        """Destination arrays (length n_new) for a particle exchange: the spare sort buffers when
        they are large enough, else temporary allocations."""
        if n_new <= self._capacity:
            return {k: self.sorting_buffers[i].view((n_new,)) for i, k in enumerate(FLOAT_ATTRS)}
        return {k: DeviceArray(n_new, np.float64) for k in FLOAT_ATTRS}

    def generate_continuously_injected_particles(self, time):
        """Float buffer (8, N) of the particles entering through the right edge
        (particles.py:335-374; no tracker / ionizer quantities in this build)."""
        assert self.continuous_injection is True
        Ntot, x, y, z, ux, uy, uz, inv_gamma, w = self.injector.generate_particles(time)
        return np.ascontiguousarray(np.stack((x, y, z, ux, uy, uz, inv_gamma, w)))

    def _need_gpu(self):
        if not self.data_is_on_gpu:
            raise _lib.B200Error('particle data is on the host: call send_particles_to_gpu() '
                                 '(fbpic_b200 has no CPU path)')

    # ------------------------------------------------------------------ push / gather
    def push_p(self, t):
        """Vay momentum push (particles.py:557-636)."""
        if self.q == 0:
            return
        self._need_gpu()
        ctx = _lib.context()
        if self.ionizer is not None:                              # particles.py:590-597
            if isinstance(self.injector, BallisticBeforePlane):
                raise NotImplementedError('Ballistic injection before a plane is not implemented for ionizable '
                                          'particles.')
            call.b2_push_p_ioniz(ctx.handle, self.Ntot, self.ionizer.levels.id.ptr, self.ux.ptr, self.uy.ptr,
                                 self.uz.ptr, self.inv_gamma.ptr, self.Ex.ptr, self.Ey.ptr, self.Ez.ptr, self.Bx.ptr,
                                 self.By.ptr, self.Bz.ptr, self.m, self.dt, None)
            return
        if isinstance(self.injector, BallisticBeforePlane):       # particles.py:577-578, 599-606
            call.b2_push_p_after_plane(ctx.handle, self.Ntot, self.z.ptr, self.injector.get_current_plane_position(t),
                                       self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                                       self.Ex.ptr, self.Ey.ptr, self.Ez.ptr, self.Bx.ptr, self.By.ptr, self.Bz.ptr,
                                       self.q, self.m, self.dt, None)
            return
        call.b2_push_p(ctx.handle, self.Ntot, self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                       self.Ex.ptr, self.Ey.ptr, self.Ez.ptr, self.Bx.ptr, self.By.ptr, self.Bz.ptr,
                       self.q, self.m, self.dt, None)

    def push_x(self, dt, x_push=1., y_push=1., z_push=1.):
        """Position push (particles.py:639-671)."""
        self._need_gpu()
        ctx = _lib.context()
        call.b2_push_x(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                       self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                       dt, x_push, y_push, z_push, None)
        self.sorted = False
        self._keys_fresh = False

    @staticmethod
    def _eb_grid_ptrs(grid):
        return ptr_array([getattr(g, k) for g in grid for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz')])

    def gather(self, grid, comm):
        """E, B grid -> particles, all azimuthal modes in one pass (particles.py:673-837)."""
        if self.q == 0:
            return
        self._need_gpu()
        ctx = _lib.context()
        g0 = grid[0]
        call.b2_gather(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                       comm.get_rmax(with_damp=False), g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr,
                       len(grid), self._eb_grid_ptrs(grid), int(self.particle_shape == 'cubic'),
                       self.Ex.ptr, self.Ey.ptr, self.Ez.ptr, self.Bx.ptr, self.By.ptr, self.Bz.ptr, None)

    def gather_and_push(self, grid, comm, dt_x, key_zmin=None):
        """Fused gather + push_p + push_x(dt_x): the call sequence main.py:470-490 in one
        kernel (the gathered fields stay in registers; Ex..Bz are not written).  With
        `key_zmin` the kernel also emits the cell keys of the new positions for a grid whose
        left edge is key_zmin, so that the next sort skips its cell-index pass."""
        if self.q == 0:
            self.push_x(dt_x)
            return
        self._need_gpu()
        ctx = _lib.context()
        g0 = grid[0]
        keys = None
        if key_zmin is not None:
            if self.cell_idx is None or self.cell_idx.size != self.Ntot:
                self._alloc_sort_arrays()
            keys = self.cell_idx.ptr
        call.b2_gather_push(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                            self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                            comm.get_rmax(with_damp=False), g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin,
                            g0.Nr, len(grid), self._eb_grid_ptrs(grid), int(self.particle_shape == 'cubic'),
                            self.q, self.m, self.dt, dt_x, keys, 0. if key_zmin is None else key_zmin, None)
        self.sorted = False
        self._keys_fresh = key_zmin is not None

    def push_x_and_key(self, dt, fld, wrap=None, key_zmin=None):
        """push_x(dt) fused with the periodic wrap of z into `wrap=(zmin, zmax)` and with the
        cell-key computation of the following sort (on a grid starting at key_zmin)."""
        self._need_gpu()
        g0 = fld.interp[0]
        keys = None
        if key_zmin is not None:
            if self.cell_idx is None or self.cell_idx.size != self.Ntot:
                self._alloc_sort_arrays()
            keys = self.cell_idx.ptr
        wz = wrap if wrap is not None else (0., 0.)
        call.b2_push_x_key(_lib.context().handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                           self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr, dt,
                           int(wrap is not None), wz[0], wz[1], g0.invdz,
                           0. if key_zmin is None else key_zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, keys, None)
        self.sorted = False
        self._keys_fresh = key_zmin is not None

    # ------------------------------------------------------------------ sorting
    def sort_particles(self, fld):
        """cell key -> stable sort -> inclusive prefix sum -> permute the SoA
        (particles.py:1049-1094, cuda_sorting.py)."""
        self._need_gpu()
        ctx = _lib.context()
        g0 = fld.interp[0]
        if self.cell_idx is None or self.cell_idx.size != self.Ntot:
            self._alloc_sort_arrays()
        if not getattr(self, '_keys_fresh', False):
            call.b2_cell_index(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                               g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, self.cell_idx.ptr, None)
        self._keys_fresh = False
        call.b2_sort_cells(ctx.handle, self.Ntot, self.cell_idx.ptr, self.sorted_idx.ptr,
                           self.prefix_sum.ptr, g0.Nz, g0.Nr, None)
        self.prefix_sum_shift = 0
        self.rearrange_particle_arrays()
        self._order_matches_prefix = True

    def rearrange_particle_arrays(self):
        """One-pass permutation of all SoA attributes, then swap with the spare
        buffers (particles.py:510-555 does one launch per attribute)."""
        names = list(FLOAT_ATTRS)
        if self.keep_fields_sorted:
            names += list(FIELD_ATTRS)
        src = [getattr(self, k) for k in names]
        dst = self.sorting_buffers[:len(names)]
        call.b2_permute(_lib.context().handle, self.Ntot, self.sorted_idx.ptr, len(names),
                        ptr_array(src), ptr_array(dst), None)
        for i, k in enumerate(names):
            setattr(self, k, dst[i])
            self.sorting_buffers[i] = src[i]
        self._permute_ids(self.sorted_idx.ptr)

    def uint_carriers(self):
        """the 8-byte integer arrays that travel with the particles: tracked ids, ionization levels"""
        out = [self.tracker] if self.tracker is not None else []
        if self.ionizer is not None:
            out.append(self.ionizer.levels)
        return out

    def _permute_ids(self, sorted_idx_ptr):
        """The tracked ids and ionization levels follow the sort (particles.py:541-545); sorted_idx_ptr None: the
        permutation of the last b2_sort_cells on this context."""
        for t in self.uint_carriers():
            call.b2_permute(_lib.context().handle, self.Ntot, sorted_idx_ptr, 1, ptr_array([t.id]),
                            ptr_array([t.spare]), None)
            t.swap()
        if self.ionizer is not None:
            self.ionizer.update_weights(self)

    # ------------------------------------------------------------------ deposition
    def deposit_fused(self, fld, fieldtype, push=None):
        """Fast path used by Simulation.step(fused=True); same sums as deposit().
        * not sorted: cell sort (keys possibly already emitted by the push kernel), then ONE
          kernel that applies the permutation to the SoA and deposits (`b2_deposit_permute`);
        * rho with the arrays still in the order of the last sort (particles moved by about a cell
          since): deposited as they lie, no re-sort -- the deposition kernel reduces runs of equal
          cell key, whatever the order; the second sort of the PIC cycle (particles.py:866-871 sorts
          before every deposit) disappears.
        The API-visible `cell_idx` / `sorted_idx` are refreshed only by sort_particles()."""
        if self.q == 0:
            return
        if self.ionizer is not None:        # per-particle charge: the plain sort + deposit sequence, weight w * level
            assert push is None
            return self.deposit(fld, fieldtype)
        assert fieldtype in ('rho', 'J')
        self._need_gpu()
        ctx = _lib.context()
        grid = fld.interp
        g0, Nm = grid[0], len(grid)
        cubic = (self.particle_shape == 'cubic')
        attr = 'd_ruyten_cubic_coef' if cubic else 'd_ruyten_linear_coef'
        r0, rh = getattr(grid[0], attr), getattr(grid[1 if Nm > 1 else 0], attr)
        if fieldtype == 'rho':
            grids = ptr_array([g.rho for g in grid])
        else:
            grids = ptr_array([getattr(g, k) for g in grid for k in ('Jr', 'Jt', 'Jz')])
        if push is not None:
            # second half position push + periodic wrap + rho deposition at the new position,
            # one pass (main.py:519 and :528); push = (dt, wrap interval or None)
            assert fieldtype == 'rho'
            dt_x, wrap = push
            wz = wrap if wrap is not None else (0., 0.)
            call.b2_push_deposit_rho(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr, self.w.ptr,
                                     self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr, dt_x,
                                     int(wrap is not None), wz[0], wz[1], self.q, g0.invdz, g0.zmin, g0.Nz,
                                     g0.invdr, g0.rmin, g0.Nr, Nm, grids, r0.ptr, rh.ptr, int(cubic), None)
            self.sorted = False
            self._keys_fresh = False
            return
        if self.sorted:
            return self.deposit(fld, fieldtype)
        if fieldtype == 'rho' and getattr(self, '_order_matches_prefix', False):
            # the arrays are still in the order of the last sort and the particles moved by about a
            # cell since: the run-based deposition kernel needs no re-sort
            call.b2_deposit_rho(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr, self.w.ptr,
                                self.q, g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm, grids,
                                self.prefix_sum.ptr, r0.ptr, rh.ptr, int(cubic), None)
            return
        if fieldtype == 'J' and getattr(self, '_order_matches_prefix', False) \
                and self._j_since_sort < getattr(self, 'sort_period', 1) - 1:
            # re-sorting is only a locality optimisation for the run-based kernels: with
            # sort_period > 1 the arrays are re-sorted every sort_period-th current deposition
            self._j_since_sort += 1
            self._keys_fresh = False
            call.b2_deposit_J(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr, self.w.ptr, self.q,
                              self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                              g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm, grids,
                              self.prefix_sum.ptr, r0.ptr, rh.ptr, int(cubic), None)
            return
        if self.cell_idx is None or self.cell_idx.size != self.Ntot:
            self._alloc_sort_arrays()
        if not getattr(self, '_keys_fresh', False):
            call.b2_cell_index(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr,
                               g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, self.cell_idx.ptr, None)
        self._keys_fresh = False
        self._j_since_sort = 0
        call.b2_sort_cells(ctx.handle, self.Ntot, self.cell_idx.ptr, None, self.prefix_sum.ptr,
                           g0.Nz, g0.Nr, None)
        self.prefix_sum_shift = 0
        names = ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'inv_gamma')
        src = [getattr(self, k) for k in names]
        dst = self.sorting_buffers[:8]
        call.b2_deposit_permute(ctx.handle, int(fieldtype == 'J'), self.Ntot, ptr_array(src), ptr_array(dst),
                                self.q, g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm, grids,
                                self.prefix_sum.ptr, r0.ptr, rh.ptr, int(cubic), None)
        for i, k in enumerate(names):
            setattr(self, k, dst[i])
            self.sorting_buffers[i] = src[i]
        self._permute_ids(None)
        self.sorted = True
        self._order_matches_prefix = True

    def deposit(self, fld, fieldtype):
        """rho or J on the interpolation grid (particles.py:839-985): sorts first if
        needed, then one launch for all modes."""
        if self.q == 0:
            return
        assert fieldtype in ('rho', 'J')
        self._need_gpu()
        if not self.sorted:
            self.sort_particles(fld=fld)
            self.sorted = True
        ctx = _lib.context()
        # ionizable species: charge e, weight w * level (particles.py:875-880)
        weight = self.ionizer.w_times_level if self.ionizer is not None else self.w
        grid = fld.interp
        g0 = grid[0]
        Nm = len(grid)
        cubic = (self.particle_shape == 'cubic')
        attr = 'd_ruyten_cubic_coef' if cubic else 'd_ruyten_linear_coef'
        r0 = getattr(grid[0], attr)
        rh = getattr(grid[1 if Nm > 1 else 0], attr)
        if fieldtype == 'rho':
            call.b2_deposit_rho(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr, weight.ptr, self.q,
                                g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm,
                                ptr_array([g.rho for g in grid]), self.prefix_sum.ptr, r0.ptr, rh.ptr,
                                int(cubic), None)
        else:
            call.b2_deposit_J(ctx.handle, self.Ntot, self.x.ptr, self.y.ptr, self.z.ptr, weight.ptr, self.q,
                              self.ux.ptr, self.uy.ptr, self.uz.ptr, self.inv_gamma.ptr,
                              g0.invdz, g0.zmin, g0.Nz, g0.invdr, g0.rmin, g0.Nr, Nm,
                              ptr_array([getattr(g, k) for g in grid for k in ('Jr', 'Jt', 'Jz')]),
                              self.prefix_sum.ptr, r0.ptr, rh.ptr, int(cubic), None)

    # ------------------------------------------------------------------ elementary processes
    def handle_elementary_processes(self, t):
        """Ionization, Compton scattering (particles.py:497-509)"""
        if self.ionizer is not None:
            self.ionizer.handle_ionization(self)
        if self.compton_scatterer is not None:
            self.compton_scatterer.handle_scattering(self, t)

    def make_ionizable(self, element, target_species, level_start=0, level_max=None):
        """ADK ionization of this species; the freed electrons go to `target_species` (a `Particles` object, or
        {level: Particles}) (particles.py:398-495).  The charge becomes e: the deposition uses w * level as weight."""
        from .ionization import Ionizer
        if self.data_is_on_gpu:
            raise _lib.B200Error('make_ionizable acts on the host copy of the data: call it before step()')
        from scipy.constants import e
        self.ionizer = Ionizer(element, self, target_species, level_start, level_max=level_max)
        self.q = e

    def grow_device_arrays(self, n_new, new_uint=None):
        """Room for n_new - Ntot more particles at the end of every per-particle device array (the caller fills the
        8 state arrays of the new ones; their gathered fields are zero until the next gather).  `new_uint`: callable
        carrier -> values of the new entries of that 8-byte array (default: `carrier.generate_new_ids`)."""
        old_n = self.Ntot
        add = n_new - old_n
        if n_new <= self._capacity:
            for k in FLOAT_ATTRS + FIELD_ATTRS:
                setattr(self, k, getattr(self, k).view((n_new,)))
            self.sorting_buffers = [b.view((n_new,)) for b in self.sorting_buffers]
            self.cell_idx, self.sorted_idx = self.cell_idx.view((n_new,)), self.sorted_idx.view((n_new,))
            for t in self.uint_carriers():
                t.id, t.spare = t.id.view((n_new,)), t.spare.view((n_new,))
        else:
            cap = self._capacity_for(n_new)
            for k in FLOAT_ATTRS + FIELD_ATTRS:
                d = DeviceArray(cap, np.float64).view((n_new,))
                d.view((old_n,)).copy_from(getattr(self, k))
                setattr(self, k, d)
            for t in self.uint_carriers():
                d = DeviceArray(cap, np.uint64).view((n_new,))
                d.view((old_n,)).copy_from(t.id)
                t.id, t.spare = d, DeviceArray(cap, np.uint64).view((n_new,))
            self._capacity = cap
            self.Ntot = n_new
            self._alloc_sort_arrays()
        for k in FIELD_ATTRS:
            call.b2_memset(getattr(self, k).ptr + 8 * old_n, 0, 8 * add, _lib.context().stream)
        for t in self.uint_carriers():
            values = t.generate_new_ids(add) if new_uint is None else new_uint(t)
            t.id.view((add,), byte_offset=8 * old_n).set(values)
        self.Ntot = n_new
        self._order_matches_prefix = False
        self._keys_fresh = False
        self._j_since_sort = 0
        self.sorted = False

    def activate_compton(self, target_species, laser_energy, laser_wavelength, laser_waist, laser_ctau,
                         laser_initial_z0, ratio_w_electron_photon, boost=None):
        """Compton scattering of a counter-propagating Gaussian laser pulse off this (electron) species; the photons
        go to `target_species` (q = 0, m = 0) (particles.py:376-396)."""
        from .compton import ComptonScatterer
        self.compton_scatterer = ComptonScatterer(self, target_species, laser_energy, laser_wavelength, laser_waist,
                                                  laser_ctau, laser_initial_z0, ratio_w_electron_photon, boost)

    def shift_periodic(self, zmin, zmax):
        """Single periodic domain: wrap z back into the box
        (boundaries/particle_buffer_handling.py:514-560)."""
        self._need_gpu()
        call.b2_shift_periodic(_lib.context().handle, self.Ntot, self.z.ptr, zmin, zmax, None)
        self.sorted = False
        self._keys_fresh = False

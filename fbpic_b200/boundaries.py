"""
`BoundaryCommunicator`: z-slab domain decomposition, guard-cell exchange of the
fields, particle migration and open-boundary damping, on NCCL instead of
mpi4py.  Mirrors fbpic/boundaries/boundary_communicator.py:28-1222 for the hot
path (same constructor arguments, attribute and method names); the gather /
scatter routines used by the diagnostics are out of scope.

One process drives one GPU.  rank/size come from the launcher's environment
(RANK, WORLD_SIZE: torchrun) -- the role `mpi4py.MPI.COMM_WORLD` plays in the
reference (fbpic/utils/mpi.py:10-76).  The NCCL communicator lives inside the C
library; its 128-byte unique id is broadcast once through torch.distributed
(gloo), which is host plumbing only.

The slab arithmetic (`decompose_z`, `halo_plan`) is pure Python so that it can
be tested without a GPU.
"""
import ctypes
import os
import numpy as np
from scipy.constants import c

from . import _lib
from . import host_tables as ht
from ._lib import DeviceArray, call, ptr_array


# ---------------------------------------------------------------------------
# process group (replaces fbpic/utils/mpi.py)
# ---------------------------------------------------------------------------
class ProcessGroup(object):
    """rank/size of the job and the NCCL communicator between the ranks."""

    def __init__(self):
        self.rank = int(os.environ.get('RANK', '0'))
        self.size = int(os.environ.get('WORLD_SIZE', '1'))
        self._nccl_ready = False

    def init_nccl(self):
        if self._nccl_ready or self.size == 1:
            return
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('gloo')
        ident = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            call.b2_nccl_unique_id(ident.ctypes.data)
        t = torch.from_numpy(ident)
        dist.broadcast(t, src=0)
        call.b2_nccl_init(_lib.context().handle, ident.ctypes.data, self.rank, self.size)
        self._nccl_ready = True


_world = None


def world():
    global _world
    if _world is None:
        _world = ProcessGroup()
    return _world


# ---------------------------------------------------------------------------
# pure slab arithmetic (testable on CPU)
# ---------------------------------------------------------------------------
def decompose_z(Nz_global, size, rank, n_guard, nz_damp, n_inject, with_damp=True, with_guard=True):
    """(Nz, iz_start) of the local grid of `rank`; iz counted from the first cell of the
    global physical domain (boundary_communicator.py:399-457: every rank gets
    int(Nz/size) cells, the last one takes the remainder)."""
    per = int(Nz_global / size)
    Nz = per
    iz = rank * per
    if rank == size - 1:
        Nz += Nz_global % size
    if with_damp:
        if rank == 0:
            Nz += nz_damp + n_inject
            iz -= nz_damp + n_inject
        if rank == size - 1:
            Nz += nz_damp + n_inject
    if with_guard:
        Nz += 2 * n_guard
        iz -= n_guard
    return Nz, iz


def halo_plan(Nz, ng, method):
    """Row ranges of a local [Nz, Nr] array involved in a guard-cell exchange
    (field_buffer_handling.py:180-188, 270-347; boundaries/cuda_methods.py:56-69, 239-252).

    Returns dict(send_l, send_r, recv_l, recv_r) of (start, stop) row ranges:
      replace: send the inner `ng` valid rows, overwrite the guard rows;
      add    : send guard+inner (2*ng rows), ADD into the same 2*ng rows."""
    if method == 'replace':
        return dict(send_l=(ng, 2 * ng), send_r=(Nz - 2 * ng, Nz - ng),
                    recv_l=(0, ng), recv_r=(Nz - ng, Nz))
    if method == 'add':
        return dict(send_l=(0, 2 * ng), send_r=(Nz - 2 * ng, Nz),
                    recv_l=(0, 2 * ng), recv_r=(Nz - 2 * ng, Nz))
    raise ValueError(method)


# ---------------------------------------------------------------------------
class BoundaryCommunicator(object):

    def __init__(self, Nz, zmin, zmax, Nr, rmax, Nm, dt, v_comoving, use_galilean, boundaries,
                 n_order, n_guard, n_damp, cdt_over_dr, n_inject=None, exchange_period=None,
                 use_all_mpi_ranks=True):
        self.Nm = Nm
        self._Nr = Nr
        self._Nz_global_domain = Nz
        self._zmin_global_domain = zmin
        self.dz = (zmax - zmin) / self._Nz_global_domain
        self.dr = rmax / self._Nr
        if not isinstance(boundaries, dict) or 'z' not in boundaries or 'r' not in boundaries:
            raise ValueError("The argument `boundaries` should be a dictionary,\n"
                             "whose keys are 'z' and 'r'.")
        if boundaries['z'] not in ('periodic', 'open'):
            raise ValueError("Unrecognized `boundaries['z']`: '%s'" % boundaries['z'])
        if boundaries['r'] not in ('reflective', 'open'):
            raise ValueError("Unrecognized `boundaries['r']`: '%s'" % boundaries['r'])
        self.use_all_mpi_ranks = use_all_mpi_ranks
        w = world()
        if use_all_mpi_ranks and w.size > 1:
            self.mpi_comm, self.rank, self.size = w, w.rank, w.size
        else:
            self.mpi_comm, self.rank, self.size = None, 0, 1
        self.left_proc, self.right_proc = self.rank - 1, self.rank + 1
        self.boundaries = boundaries
        if boundaries['z'] == 'periodic':
            if self.rank == 0:
                self.left_proc = self.size - 1
            if self.rank == self.size - 1:
                self.right_proc = 0
        else:
            if self.rank == 0:
                self.left_proc = None
            if self.rank == self.size - 1:
                self.right_proc = None
        # guard cells (boundary_communicator.py:225-254)
        if n_guard is None:
            if n_order == -1:
                self.n_guard = 64
                if self.size != 1:
                    raise ValueError('When running with domain decomposition, you need to set\n'
                                     'the argument `n_order` of the `Simulation` object to a\n'
                                     'positive value (e.g. n_order=32).')
            else:
                self.n_guard = ht.stencil_reach(self._Nz_global_domain, self.dz, c * dt, n_order,
                                                v_comoving, use_galilean) + 1
        else:
            self.n_guard = n_guard
        if boundaries['z'] == 'periodic' and self.size == 1:
            self.n_guard = 0
        self.nz_damp, self.nr_damp = n_damp['z'], n_damp['r']
        if boundaries['z'] == 'periodic':
            self.nz_damp, self.n_inject = 0, 0
        else:
            self.n_inject = int(self.n_guard / 2) if n_inject is None else n_inject
        # radial PML (boundary_communicator.py:272-278, 332-335; pml_damping.py:23-44, 86-108)
        if boundaries['r'] == 'reflective':
            self.nr_damp = 0
        self.use_pml = (boundaries['r'] == 'open')
        self.pml_damp_array = ht.pml_damp_array(self.nr_damp, cdt_over_dr) if self.use_pml else None
        self.d_pml_damp_array = None
        # exchange period (boundary_communicator.py:281-304)
        if exchange_period is None:
            cells_per_step = 2. * c * dt / self.dz
            self.exchange_period = int(((self.n_guard / 2) - 3) / cells_per_step)
            if self.size == 1 and boundaries['z'] == 'periodic':
                self.exchange_period = 1
            if self.exchange_period < 1:
                raise ValueError('Guard region size is too small for chosen timestep.')
        else:
            self.exchange_period = exchange_period
        self.moving_win = None
        self.left_damp = self.right_damp = None
        self.d_left_damp = self.d_right_damp = None
        if (self.nz_damp + self.n_inject) > 0:
            if self.left_proc is None:
                self.left_damp = self.generate_damp_array(self.n_guard, self.nz_damp, self.n_inject)
            if self.right_proc is None:
                self.right_damp = self.generate_damp_array(self.n_guard, self.nz_damp, self.n_inject)
        self._halo_buf = {}

    # ---- geometry (boundary_communicator.py:338-512) ----
    def divide_into_domain(self):
        zmin, zmax = self.get_zmin_zmax(local=True, with_damp=True, with_guard=True, rank=self.rank)
        Nz, _ = self.get_Nz_and_iz(local=True, with_damp=True, with_guard=True, rank=self.rank)
        if Nz < 4 * self.n_guard:
            raise ValueError('The boundary guard region is larger than the physical domain size. '
                             'Use fewer ranks or a smaller order of the field solver.')
        return zmin, zmax, Nz

    def get_Nr(self, with_damp):
        return self._Nr + (self.nr_damp if with_damp else 0)

    def get_rmax(self, with_damp):
        return (self._Nr + (self.nr_damp if with_damp else 0)) * self.dr

    def get_Nz_and_iz(self, local, with_damp, with_guard, rank=None):
        if local:
            if rank is None:
                raise ValueError('For a local number of cells, the rank considered is needed.')
            return decompose_z(self._Nz_global_domain, self.size, rank, self.n_guard, self.nz_damp,
                               self.n_inject, with_damp, with_guard)
        Nz, iz = self._Nz_global_domain, 0
        if with_damp:
            Nz += 2 * (self.nz_damp + self.n_inject)
            iz -= self.nz_damp + self.n_inject
        if with_guard:
            Nz += 2 * self.n_guard
            iz -= self.n_guard
        return Nz, iz

    def get_zmin_zmax(self, local, with_damp, with_guard, rank=None):
        Nz, iz_start = self.get_Nz_and_iz(local=local, with_damp=with_damp, with_guard=with_guard, rank=rank)
        zmin = self._zmin_global_domain + iz_start * self.dz
        return zmin, zmin + Nz * self.dz

    def shift_global_domain_positions(self, z_shift):
        self._zmin_global_domain += z_shift

    # ---- field guard cells (boundary_communicator.py:556-707) ----
    def exchange_fields(self, interp, fldtype, method):
        """Exchange the guard cells of `fldtype` ('E','B','J','rho') with the z-neighbours:
        'replace' overwrites the local guard rows, 'add' sums guard+inner rows."""
        if self.size == 1:
            return
        self.mpi_comm.init_nccl()
        ctx = _lib.context()
        if fldtype == 'EB':         # fused step: E and B in one NCCL group (same data as 'E' then 'B')
            names = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz')
        else:
            names = ('rho',) if fldtype == 'rho' else (fldtype + 'r', fldtype + 't', fldtype + 'z')
        if self.use_pml and fldtype in ('E', 'B', 'EB'):
            # the split PML components travel with their field (boundary_communicator.py:621-627)
            names = names + tuple(f + c + '_pml' for f in fldtype for c in 'rt')
        arrays = [getattr(interp[m], n) for m in range(self.Nm) for n in names]
        if len(arrays) > _lib.MAX_ARRAYS:
            # more slabs than one staging launch carries: exchange them in chunks
            for i in range(0, len(arrays), _lib.MAX_ARRAYS):
                self._exchange_arrays(arrays[i:i + _lib.MAX_ARRAYS], method)
            return
        self._exchange_arrays(arrays, method)

    def _exchange_arrays(self, arrays, method):
        ctx = _lib.context()
        Nz, Nr = arrays[0].shape
        ng = self.n_guard
        plan = halo_plan(Nz, ng, method)
        nrow = plan['send_l'][1] - plan['send_l'][0]
        na = len(arrays)
        key = ('fld', na, nrow, Nr)
        if key not in self._halo_buf:
            self._halo_buf[key] = [DeviceArray((na, nrow, Nr), np.complex128) for _ in range(4)]
        send_l, send_r, recv_l, recv_r = self._halo_buf[key]
        nbytes = na * nrow * Nr * 16
        ptrs = ptr_array(arrays)
        # one staging kernel per face packs the slabs of all the arrays (rows are contiguous), ONE NCCL
        # message per neighbour and direction, one kernel per face replaces / adds the received rows
        if self.left_proc is not None:
            call.b2_halo_stage(ctx.handle, 0, na, ptrs, plan['send_l'][0], nrow, Nr, send_l.ptr, None)
        if self.right_proc is not None:
            call.b2_halo_stage(ctx.handle, 0, na, ptrs, plan['send_r'][0], nrow, Nr, send_r.ptr, None)
        # NCCL pairs the k-th send to a peer with the k-th receive posted for that peer.  What goes
        # out through my LEFT face arrives at my left neighbour's RIGHT face, so sends are posted
        # (left, right) and receives (right, left): with 2 ranks on a ring both neighbours are the
        # same peer and this order is what keeps the two directions apart (the reference uses MPI
        # tags 1/2 for that, boundary_communicator.py:688-699).
        call.b2_comm_begin(ctx.handle)
        if self.left_proc is not None:
            call.b2_nccl_send(ctx.handle, send_l.ptr, nbytes, self.left_proc, None)
        if self.right_proc is not None:
            call.b2_nccl_send(ctx.handle, send_r.ptr, nbytes, self.right_proc, None)
            call.b2_nccl_recv(ctx.handle, recv_r.ptr, nbytes, self.right_proc, None)
        if self.left_proc is not None:
            call.b2_nccl_recv(ctx.handle, recv_l.ptr, nbytes, self.left_proc, None)
        call.b2_comm_end(ctx.handle)
        mode = 1 if method == 'replace' else 2
        if self.left_proc is not None:
            call.b2_halo_stage(ctx.handle, mode, na, ptrs, plan['recv_l'][0], nrow, Nr, recv_l.ptr, None)
        if self.right_proc is not None:
            call.b2_halo_stage(ctx.handle, mode, na, ptrs, plan['recv_r'][0], nrow, Nr, recv_r.ptr, None)

    # ---- particles (boundary_communicator.py:710-826) ----
    def exchange_particles(self, species, fld, time):
        if self.n_guard == 0:
            species.shift_periodic(fld.interp[0].zmin, fld.interp[0].zmax)
        else:
            self.exchange_particles_aperiodic_subdomain(species, fld, time)

    def exchange_particles_aperiodic_subdomain(self, species, fld, time):
        """Particles that left the physical part of the slab go to the z-neighbours (or are dropped
        at an open end), the others stay; received ones are placed before (from the left) and after
        (from the right) the kept block, new plasma injected by a moving window counts as received
        from the right (boundary_communicator.py:750-826).  Selection rule and ordering are those of
        the reference CPU path (remove_particles_cpu / add_buffers_cpu,
        particle_buffer_handling.py:58-175, 424-512): left if z < zmin + n_guard*dz, right if
        z > zmax - n_guard*dz, a stable 3-way partition of the SoA done on the device."""
        from .particles import FLOAT_ATTRS
        species._need_gpu()
        ctx = _lib.context()
        g0 = fld.interp[0]
        ng = self.n_guard
        zlo = g0.zmin + ng * g0.dz
        zhi = g0.zmax - ng * g0.dz
        N = species.Ntot
        # new plasma entering through the right edge (boundary_communicator.py:803-810): generated on the host FIRST
        # -- it does not depend on the classification, whose call returns counts and therefore waits for every cycle
        # still queued on the device; the 4 ms of NumPy work per exchange now run under those cycles
        injected = None
        if (self.moving_win is not None) and (self.rank == self.size - 1) and species.continuous_injection:
            injected = species.generate_continuously_injected_particles(time)
        counts = (ctypes.c_int64 * 3)()
        call.b2_exchange_classify(ctx.handle, N, species.z.ptr, zlo, zhi, counts, None)
        n_stay, n_left, n_right = int(counts[0]), int(counts[1]), int(counts[2])
        n_send_l = n_left if self.left_proc is not None else 0
        n_send_r = n_right if self.right_proc is not None else 0
        n_recv_l = n_recv_r = 0
        if self.size > 1:
            self.mpi_comm.init_nccl()
            cnt = DeviceArray.from_numpy(np.array([n_send_l, n_send_r, 0, 0], dtype=np.int64))
            call.b2_nccl_group_start()       # sends (left, right), receives (right, left): see exchange_fields
            if self.left_proc is not None:
                call.b2_nccl_send(ctx.handle, cnt.ptr, 8, self.left_proc, None)
            if self.right_proc is not None:
                call.b2_nccl_send(ctx.handle, cnt.ptr + 8, 8, self.right_proc, None)
                call.b2_nccl_recv(ctx.handle, cnt.ptr + 24, 8, self.right_proc, None)
            if self.left_proc is not None:
                call.b2_nccl_recv(ctx.handle, cnt.ptr + 16, 8, self.left_proc, None)
            call.b2_nccl_group_end()
            h = cnt.get()
            n_recv_l, n_recv_r = int(h[2]), int(h[3])
        if injected is not None:
            n_recv_r = injected.shape[1]
        n_new = n_recv_l + n_stay + n_recv_r
        new = species.exchange_buffers(n_new)       # spare sort buffers: no allocation in steady state
        send_l = self._send_buffer('l', 8 * n_send_l)
        send_r = self._send_buffer('r', 8 * n_send_r)
        src = [getattr(species, k) for k in FLOAT_ATTRS]
        stay = [new[k].ptr + 8 * n_recv_l for k in FLOAT_ATTRS]
        left = [send_l.ptr + 8 * n_send_l * i for i in range(8)] if n_send_l else None
        right = [send_r.ptr + 8 * n_send_r * i for i in range(8)] if n_send_r else None
        if N:
            call.b2_exchange_scatter(ctx.handle, N, species.z.ptr, zlo, zhi, 8, ptr_array(src), ptr_array(stay),
                                     ptr_array(left) if left else None, ptr_array(right) if right else None, None)
        if self.size > 1:
            call.b2_nccl_group_start()
            for i, k in enumerate(FLOAT_ATTRS):
                if self.left_proc is not None and n_send_l:
                    call.b2_nccl_send(ctx.handle, left[i], 8 * n_send_l, self.left_proc, None)
                if self.right_proc is not None and n_send_r:
                    call.b2_nccl_send(ctx.handle, right[i], 8 * n_send_r, self.right_proc, None)
                if self.right_proc is not None and n_recv_r and injected is None:
                    call.b2_nccl_recv(ctx.handle, new[k].ptr + 8 * (n_recv_l + n_stay), 8 * n_recv_r,
                                      self.right_proc, None)
                if self.left_proc is not None and n_recv_l:
                    call.b2_nccl_recv(ctx.handle, new[k].ptr, 8 * n_recv_l, self.left_proc, None)
            call.b2_nccl_group_end()
        if injected is not None and n_recv_r:
            for i, k in enumerate(FLOAT_ATTRS):
                new[k].view((n_recv_r,), byte_offset=8 * (n_recv_l + n_stay)).set(injected[i])
        # periodic images: shift z by the box length (boundary_communicator.py:815-821)
        Ltot = self._Nz_global_domain * self.dz
        if self.right_proc == 0 and n_recv_r:
            self._shift_z(new['z'], n_recv_l + n_stay, n_recv_r, +Ltot)
        if self.left_proc == self.size - 1 and n_recv_l:
            self._shift_z(new['z'], 0, n_recv_l, -Ltot)
        for carrier in species.uint_carriers():       # tracked ids, ionization levels
            self._exchange_ids(species, carrier, N, zlo, zhi, n_new, n_stay, n_send_l, n_send_r, n_recv_l, n_recv_r,
                               injected is not None)
        species.resize_device_arrays(new, n_new)
        if species.ionizer is not None:
            species.ionizer.update_weights(species)

    def _exchange_ids(self, species, t, N, zlo, zhi, n_new, n_stay, n_send_l, n_send_r, n_recv_l, n_recv_r, injected):
        """An 8-byte integer array `t.id` (tracked ids; ionization levels) takes the same 3-way partition as the float
        attributes (the classification of
        b2_exchange_classify is still valid: one more b2_exchange_scatter with the id array), travel to the
        neighbours in a message of their own, and new ids are drawn for injected plasma
        (particle_buffer_handling.py:117-167, 409-411; particles.py:367-368)."""
        ctx = _lib.context()
        cap = max(species._capacity, species._capacity_for(n_new))
        dest = t.spare.view((n_new,)) if t.spare.capacity >= n_new else DeviceArray(cap, np.uint64).view((n_new,))
        id_l = DeviceArray(max(n_send_l, 1), np.uint64)
        id_r = DeviceArray(max(n_send_r, 1), np.uint64)
        if N:
            call.b2_exchange_scatter(ctx.handle, N, species.z.ptr, zlo, zhi, 1, ptr_array([t.id]),
                                     ptr_array([dest.ptr + 8 * n_recv_l]), ptr_array([id_l]) if n_send_l else None,
                                     ptr_array([id_r]) if n_send_r else None, None)
        if self.size > 1:
            call.b2_nccl_group_start()
            if self.left_proc is not None and n_send_l:
                call.b2_nccl_send(ctx.handle, id_l.ptr, 8 * n_send_l, self.left_proc, None)
            if self.right_proc is not None and n_send_r:
                call.b2_nccl_send(ctx.handle, id_r.ptr, 8 * n_send_r, self.right_proc, None)
            if self.right_proc is not None and n_recv_r and not injected:
                call.b2_nccl_recv(ctx.handle, dest.ptr + 8 * (n_recv_l + n_stay), 8 * n_recv_r, self.right_proc, None)
            if self.left_proc is not None and n_recv_l:
                call.b2_nccl_recv(ctx.handle, dest.ptr, 8 * n_recv_l, self.left_proc, None)
            call.b2_nccl_group_end()
        if injected and n_recv_r:
            dest.view((n_recv_r,), byte_offset=8 * (n_recv_l + n_stay)).set(t.generate_new_ids(n_recv_r))
        old = t.id
        t.id = dest
        t.spare = old.view((n_new,)) if old.capacity >= n_new else DeviceArray(cap, np.uint64).view((n_new,))

    def _send_buffer(self, side, n_doubles):
        """Grow-only device staging buffer for the particles leaving through one face."""
        buf = self._halo_buf.get(('ptcl', side))
        if buf is None or buf.size < max(n_doubles, 1):
            buf = DeviceArray(max(int(n_doubles * 1.5), 1 << 16), np.float64)
            self._halo_buf[('ptcl', side)] = buf
        return buf

    @staticmethod
    def _shift_z(z, start, count, dz_shift):
        """z[start:start+count] += dz_shift for the periodic images received across the ring
        closure (only the two end ranks, only on particle-exchange steps)."""
        call.b2_add_scalar(_lib.context().handle, count, z.ptr + 8 * start, dz_shift, None)

    # ---- open-boundary damping (boundary_communicator.py:828-945) ----
    def generate_damp_array(self, n_guard, nz_damp, n_inject):
        return ht.damp_array(n_guard, nz_damp, n_inject)

    def damp_EB_open_boundary(self, interp):
        if self.nz_damp == 0:
            return
        nd = self.n_guard + self.nz_damp + self.n_inject
        left, right = self.left_proc is None, self.right_proc is None
        if not (left or right):
            return
        if self.d_left_damp is None:
            arr = self.left_damp if self.left_damp is not None else self.right_damp
            self.d_left_damp = self.d_right_damp = DeviceArray.from_numpy(arr)
        names = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz')
        if interp[0].use_pml:       # boundary_communicator.py:856-860, 893-897
            names = names + ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')
        for m in range(len(interp)):
            arrays = [getattr(interp[m], n) for n in names]
            call.b2_damp_z(_lib.context().handle, len(arrays), ptr_array(arrays), self.d_left_damp.ptr, nd,
                           int(left), int(right), interp[0].Nz, interp[0].Nr, None)

    def damp_pml_EB(self, interp):
        """Anisotropic damping of E, B in the last nr_damp radial cells (pml_damping.py:46-83)."""
        if not self.use_pml:
            return
        if self.d_pml_damp_array is None:
            self.d_pml_damp_array = DeviceArray.from_numpy(self.pml_damp_array)
        for g in interp:
            call.b2_damp_pml(_lib.context().handle, g.Et.ptr, g.Et_pml.ptr, g.Ez.ptr, g.Bt.ptr, g.Bt_pml.ptr,
                             g.Bz.ptr, self.d_pml_damp_array.ptr, self.nr_damp, g.Nz, g.Nr, None)

    def move_grids(self, fld, ptcl, dt, time):
        """boundary_communicator.py:531-553"""
        self.moving_win.move_grids(fld, ptcl, self, time)

    # ---- global <-> local grid arrays on the host (boundary_communicator.py:1011-1130); used by the one-off
    #      set-up routines (laser injection, bunch space charge), never inside the PIC cycle ----
    def _host_group(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('gloo')
        return dist

    def barrier(self):
        if self.size > 1:
            self._host_group().barrier()

    def gather_ptcl_array(self, array, n_rank=None, Ntot=None):
        """The particle arrays of all ranks, concatenated in rank order (boundary_communicator.py:1132-1168; every
        rank receives the result)."""
        if self.size == 1:
            return array
        parts = [None] * self.size
        self._host_group().all_gather_object(parts, np.ascontiguousarray(array))
        return np.concatenate(parts)

    def allreduce_sum(self, values):
        """Sum of a short list of floats over the ranks (mpi_comm.allreduce in the reference)."""
        if self.size == 1:
            return list(values)
        import torch
        t = torch.tensor(list(values), dtype=torch.float64)
        self._host_group().all_reduce(t)
        return [float(v) for v in t]

    def allreduce_max(self, values):
        if self.size == 1:
            return list(values)
        import torch
        import torch.distributed as dist
        t = torch.tensor(list(values), dtype=torch.float64)
        self._host_group().all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather_grid_array(self, array, root=0, with_damp=False):
        """Global array (no guard cells; damp cells if with_damp) assembled from the local host arrays.
        Unlike the reference (root only) every rank receives it: the set-up solves that use it run
        redundantly on every GPU instead of on rank 0 + scatter."""
        Nr = self.get_Nr(with_damp=with_damp)
        Nz_local, iz_dom = self.get_Nz_and_iz(local=True, with_damp=with_damp, with_guard=False, rank=self.rank)
        _, iz_arr = self.get_Nz_and_iz(local=True, with_damp=True, with_guard=True, rank=self.rank)
        i0 = iz_dom - iz_arr
        local = np.ascontiguousarray(array[i0:i0 + Nz_local, :Nr])
        if self.size == 1:
            return local.copy()
        parts = [None] * self.size
        self._host_group().all_gather_object(parts, local)
        return np.concatenate(parts, axis=0)

    def gather_grid(self, grid, root=0):
        """A global InterpolationGrid (physical domain only: no guard, damp or PML cells) holding the host
        data of the local grids of all ranks (boundary_communicator.py:964-1009); available on every rank."""
        from .fields import InterpolationGrid, INTERP_FIELDS
        Nz_g, _ = self.get_Nz_and_iz(local=False, with_guard=False, with_damp=False)
        zmin_g, zmax_g = self.get_zmin_zmax(local=False, with_guard=False, with_damp=False)
        out = InterpolationGrid(Nz_g, self.get_Nr(with_damp=False), grid.m, zmin_g, zmax_g,
                                self.get_rmax(with_damp=False))
        for k in INTERP_FIELDS:
            a = getattr(grid, k)
            if hasattr(a, 'ptr'):
                raise _lib.B200Error('gather_grid acts on the host copy of the fields: call it after step() or '
                                     'receive_data_from_gpu()')
            setattr(out, k, self.gather_grid_array(a, root))
        return out

    def scatter_grid_array(self, array, root=0, with_damp=False):
        """The local part (no guard cells) of a global array that every rank holds."""
        Nz_global, iz_glob = self.get_Nz_and_iz(local=False, with_damp=with_damp, with_guard=False)
        Nr = self.get_Nr(with_damp=with_damp)
        assert array.shape == (Nz_global, Nr)
        Nz_local, iz_dom = self.get_Nz_and_iz(local=True, with_damp=with_damp, with_guard=False, rank=self.rank)
        i0 = iz_dom - iz_glob
        return np.array(array[i0:i0 + Nz_local], dtype=np.complex128)

    def bcast_int(self, value):
        """Rank 0's integer on every rank (the n_move broadcast of moving_window.py:97)."""
        if self.size == 1:
            return value
        import torch
        import torch.distributed as dist
        t = torch.tensor([0 if value is None else int(value)], dtype=torch.int64)
        dist.broadcast(t, src=0)
        return int(t[0])

"""
`MovingWindow`: the simulation box follows the laser/beam at speed v.  Mirrors
fbpic/boundaries/moving_window.py:14-278: every step the window position advances; whenever it has
crossed whole cells the grid coordinates are shifted by n_move cells and the spectral E, B (and
rho_prev) arrays are multiplied by exp(i kz dz)^n_move -- a translation in real space; new plasma is
injected at the right edge by the species' `ContinuousInjector` at the next particle exchange.
"""
import numpy as np

from . import _lib
from ._lib import DeviceArray, call, ptr_array


class MovingWindow(object):

    def __init__(self, comm, dt, v, time):
        if ((comm.rank == comm.size - 1) and (comm.right_proc is not None)) \
                or ((comm.rank == 0) and (comm.left_proc is not None)):
            raise ValueError('The simulation is using a moving window, but the boundaries are periodic.\n'
                             'Please select open boundaries when initializing the Simulation object.')
        self.v = v
        self.t_last_move = time - dt
        zmin_global, _ = comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
        if comm.rank == 0:
            self.zmin = zmin_global

    def move_grids(self, fld, ptcl, comm, time):
        """moving_window.py:60-132"""
        fld.join_side()        # the shift touches the spectral E, B
        dz = comm.dz
        if comm.rank == 0:
            self.zmin += self.v * (time - self.t_last_move)
            zmin_global, _ = comm.get_zmin_zmax(local=False, with_damp=False, with_guard=False)
            n_move = int((self.zmin - zmin_global) / dz)
        else:
            n_move = None
        if comm.size > 1:
            n_move = comm.bcast_int(n_move)
        if n_move != 0:
            comm.shift_global_domain_positions(n_move * dz)
            for m in range(len(fld.interp)):
                fld.interp[m].zmin += n_move * fld.interp[m].dz
                fld.interp[m].zmax += n_move * fld.interp[m].dz
                self.shift_spect_grid(fld.spect[m], n_move)
        for species in ptcl:
            species.prefix_sum_shift += n_move
        if comm.rank == comm.size - 1:
            for species in ptcl:
                if species.continuous_injection and species.injector is not None:
                    species.injector.increment_injection_positions(self.v, time - self.t_last_move)
        self.t_last_move = time

    def shift_spect_grid(self, grid, n_move, shift_rho=True, shift_currents=True):
        """moving_window.py:134-202: one launch for all the arrays of the mode.  (The reference's
        signature defaults to shift_currents=True -- its docstring says False -- so the J returned
        at the end of step() is the shifted one.)"""
        names = ['Ep', 'Em', 'Ez', 'Bp', 'Bm', 'Bz']
        if grid.use_pml:        # moving_window.py:172-176
            names += ['Ep_pml', 'Em_pml', 'Bp_pml', 'Bm_pml']
        if shift_rho:
            names.append('rho_prev')
        if shift_currents:
            names += ['Jp', 'Jm', 'Jz']
        arrays = [getattr(grid, k) for k in names]
        if not isinstance(arrays[0], DeviceArray):
            raise _lib.B200Error('field data is on the host: call send_fields_to_gpu()')
        if getattr(grid, 'd_field_shift', None) is None:
            grid.d_field_shift = DeviceArray.from_numpy(np.ascontiguousarray(grid.field_shift))
        call.b2_shift_spect(_lib.context().handle, len(arrays), ptr_array(arrays), grid.d_field_shift.ptr,
                            int(n_move), grid.Nz, grid.Nr, None)

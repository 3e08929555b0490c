"""
`Simulation`: the PIC cycle driver, same constructor and `step()` contract as
fbpic/main.py:38-1111, running every per-step operation on the B200 through
libfbpic_b200.so (no Numba, no CuPy, no CPU fallback).

Hot-path scope: gather / push / deposit / sort, z-FFT + Hankel GEMM,
current correction + PSATD push, z guard-cell exchange and particle migration.
Widened per SURVEY 8f: moving window with continuous injection (rank 1), laser antennas
(rank 2), radial PML (`boundaries['r']='open'`) and the cross-deposition current correction
(rank 4); plus mirrors, external fields and the boosted-frame conversion of the set-up
(`gamma_boost`), the `sim.diags` / `sim.checkpoints` hooks (fbpic_b200/diags.py), ADK ionization and
Compton scattering (fbpic_b200/ionization.py, compton.py).  `step(show_progress=...)` defaults to False here
(reference: True).
"""
import numpy as np
from scipy.constants import m_e, m_p, e, c

from . import _lib
from .fields import Fields
from .particles import Particles
from .boundaries import BoundaryCommunicator


class Simulation(object):

    def __init__(self, Nz, zmax, Nr, rmax, Nm, dt,
                 p_zmin=-np.inf, p_zmax=np.inf, p_rmin=0, p_rmax=np.inf,
                 p_nz=None, p_nr=None, p_nt=None, n_e=None, zmin=0.,
                 n_order=-1, dens_func=None, filter_currents=True,
                 v_comoving=None, use_galilean=True,
                 initialize_ions=False, use_cuda=True, n_guard=None,
                 n_damp={'z': 64, 'r': 32}, exchange_period=None,
                 current_correction='curl-free',
                 boundaries={'z': 'periodic', 'r': 'reflective'},
                 gamma_boost=None, use_all_mpi_ranks=True,
                 particle_shape='linear', verbose_level=0,
                 smoother=None, use_ruyten_shapes=True, use_modified_volume=True,
                 fused=True, sort_period=1):
        """Arguments as in fbpic/main.py:51-229.  `use_cuda` is accepted for drop-in
        compatibility; this implementation only has the GPU path.  `fused=True` lets
        `step()` use the fused kernels (gather+push, correct+push) -- same arithmetic,
        fewer passes over HBM; `fused=False` issues one kernel per reference operator.
        `sort_period` (fused mode): the particle arrays are re-sorted by cell every sort_period-th
        step instead of before every deposition; the deposition and gather kernels are correct for
        any particle order, sorting only restores memory locality."""
        self.use_cuda = True
        self.fused = fused
        self.sort_period = max(int(sort_period), 1)
        self.v_comoving = v_comoving
        self.use_galilean = use_galilean if v_comoving is not None else False
        if gamma_boost is not None:                   # main.py:275-280
            from .lpa_utils.boosted_frame import BoostConverter
            self.boost = BoostConverter(gamma_boost)
            zmin, zmax, dt = self.boost.copropag_length([zmin, zmax, dt])
        else:
            self.boost = None
        self.dt = dt
        cdt_over_dr = c * dt / (rmax / Nr)
        self.comm = BoundaryCommunicator(Nz, zmin, zmax, Nr, rmax, Nm, dt, self.v_comoving,
                                         self.use_galilean, boundaries, n_order, n_guard, n_damp,
                                         cdt_over_dr, None, exchange_period, use_all_mpi_ranks)
        self.use_pml = self.comm.use_pml
        zmin, zmax, Nz = self.comm.divide_into_domain()
        Nr = self.comm.get_Nr(with_damp=True)
        rmax = self.comm.get_rmax(with_damp=True)
        self.fld = Fields(Nz, zmax, Nr, rmax, Nm, dt, n_order=n_order, zmin=zmin,
                          v_comoving=v_comoving, use_pml=self.use_pml, use_galilean=use_galilean,
                          current_correction=current_correction, smoother=smoother,
                          use_ruyten_shapes=use_ruyten_shapes,
                          use_modified_volume=use_modified_volume)
        self.grid_shape = self.fld.interp[0].Ez.shape
        self.particle_shape = particle_shape
        self.ptcl = []
        if n_e is not None:
            self.add_new_species(q=-e, m=m_e, n=n_e, dens_func=dens_func, p_nz=p_nz, p_nr=p_nr, p_nt=p_nt,
                                 p_zmin=p_zmin, p_zmax=p_zmax, p_rmin=p_rmin, p_rmax=p_rmax)
            if initialize_ions:
                self.add_new_species(q=e, m=m_p, n=n_e, dens_func=dens_func, p_nz=p_nz, p_nr=p_nr,
                                     p_nt=p_nt, p_zmin=p_zmin, p_zmax=p_zmax, p_rmin=p_rmin, p_rmax=p_rmax)
        self.time = 0.
        self.iteration = 0
        self.filter_currents = filter_currents
        self.external_fields, self.diags, self.checkpoints = [], [], []
        self.laser_antennas, self.mirrors = [], []
        from .printing import print_simulation_setup
        print_simulation_setup(self, verbose_level=verbose_level)

    # ------------------------------------------------------------------ data residency
    def send_data_to_gpu(self, step_entry=False):
        """fbpic/utils/cuda.py:101-118"""
        self.fld.send_fields_to_gpu(step_entry=step_entry)
        for species in self.ptcl:
            species.send_particles_to_gpu()

    def receive_data_from_gpu(self):
        """fbpic/utils/cuda.py:120-137"""
        self.fld.receive_fields_from_gpu()
        for species in self.ptcl:
            species.receive_particles_from_gpu()

    # ------------------------------------------------------------------ the PIC cycle
    def step(self, N=1, correct_currents=True, correct_divE=False, use_true_rho=False,
             move_positions=True, move_momenta=True, show_progress=False, keep_on_gpu=False):
        """N PIC cycles, call order of fbpic/main.py:346-586.  `keep_on_gpu=True` skips the
        final device->host copy (the data stays in HBM for the next call)."""
        ptcl, fld, dt = self.ptcl, self.fld, self.dt
        if self.comm.size > 1 and correct_divE:
            raise ValueError('correct_divE cannot be used in multi-proc mode.')
        if self.comm.size > 1 and use_true_rho and correct_currents:
            raise ValueError('`use_true_rho` cannot be used together with `correct_currents` '
                             'in multi-proc mode.')
        if self.comm.moving_win is not None:          # main.py:390-395
            for species in self.ptcl:
                if species.continuous_injection and species.injector is not None:
                    # (the positions are read only by the first call on the last rank, to find the end of the
                    #  plasma: no 8-byte-per-particle download at the entry of every later step() call)
                    need_z = species.injector.z_inject is None and self.comm.rank == self.comm.size - 1
                    z_host = np.empty(0) if not need_z else \
                        (species.z.get() if species.data_is_on_gpu else species.z)
                    species.injector.initialize_injection_positions(
                        self.comm, self.comm.moving_win.v, z_host, self.dt)
        single = (self.comm.size == 1)
        periodic_single = single and self.comm.n_guard == 0
        # external fields act on the gathered E, B of the particles and particle diagnostics may output them: both
        # need the unfused gather
        fuse_gp = self.fused and move_positions and move_momenta and not self.external_fields and \
            not any(getattr(diag, 'needs_gathered_fields', True) for diag in self.diags)
        fuse_cp = self.fused and correct_currents and single and fld.current_correction == 'curl-free'

        for species in ptcl:
            if not species.data_is_on_gpu:
                species.fields_resident_only = bool(fuse_gp) and not species.ballistic_before_plane and \
                    species.ionizer is None
            elif not fuse_gp:
                species.fields_resident_only = False      # an unfused gather will write them: read them back
        import time as _time
        bytes0 = dict(_lib.TRANSFERRED)
        t_start = _time.perf_counter()
        # (J, rho and every spectral array are recomputed before their first use when at least one cycle runs:
        #  their host copies stay where they are)
        self.send_data_to_gpu(step_entry=(N >= 1))
        t_sent = _time.perf_counter()
        self.comm.exchange_fields(fld.interp, 'E', 'replace')
        self.comm.exchange_fields(fld.interp, 'B', 'replace')
        self.comm.damp_EB_open_boundary(fld.interp)
        fld.interp2spect('E')
        fld.interp2spect('B')
        if self.use_pml:                              # main.py:413-415
            fld.interp2spect('E_pml')
            fld.interp2spect('B_pml')

        # Single periodic domain, fused mode: z is wrapped inside the second position push and
        # rho_prev of step n+1 is (bit for bit) the rho_next that push_rho already moved over, so
        # the per-step exchange_particles + re-deposition of rho_prev (main.py:435-449, needed in
        # the reference only because particles may have been added/removed) is done at i_step==0 only.
        wrap_in_push = self.fused and periodic_single and move_positions and not self.laser_antennas
        progress_bar = None
        if show_progress and self.comm.rank == 0:     # main.py:398-399; never synchronises the device
            from .printing import ProgressBar
            progress_bar = ProgressBar(N)
        for i_step in range(N):
            if progress_bar is not None:
                progress_bar.time(i_step)
                progress_bar.print_progress()
            exchange_now = (self.iteration % self.comm.exchange_period == 0 or i_step == 0)
            if exchange_now and not (wrap_in_push and i_step > 0):
                for species in ptcl:
                    self.comm.exchange_particles(species, fld, self.time)
                for antenna in self.laser_antennas:       # main.py:443-444
                    antenna.update_current_rank(self.comm)
                self.deposit('rho_prev', exchange=(use_true_rho is True))
            if i_step == 0:
                self.deposit('J', exchange=True)

            for species in ptcl:
                species.keep_fields_sorted = True
            gal_shift = self.v_comoving * 0.5 * dt if self.use_galilean else 0.
            if fuse_gp:
                for diag in self.diags:                   # E, B, rho, x at time n (main.py:474-481)
                    diag.write(self.iteration)
                for species in ptcl:
                    if species.ballistic_before_plane or species.ionizer is not None:
                        # the fused kernel pushes every particle: this species takes the three-kernel route
                        species.gather(fld.interp, self.comm)
                        species.push_p(self.time + 0.5 * dt)
                        species.push_x(0.5 * dt)
                        continue
                    will_sort = (not getattr(species, '_order_matches_prefix', False)) or \
                        species._j_since_sort >= self.sort_period - 1
                    species.gather_and_push(fld.interp, self.comm, 0.5 * dt,
                                            key_zmin=(fld.interp[0].zmin + gal_shift) if will_sort else None)
            else:
                for species in ptcl:
                    species.gather(fld.interp, self.comm)
                for ext_field in self.external_fields:    # main.py:472-473
                    ext_field.apply_expression(self.ptcl, self.time)
                # (E, B, rho, x are defined at time n ; J, p at time n-1/2: main.py:474-481)
                for diag in self.diags:
                    diag.write(self.iteration)
                if move_momenta:
                    for species in ptcl:
                        species.push_p(self.time + 0.5 * dt)
                if move_positions:
                    for species in ptcl:
                        species.push_x(0.5 * dt)
            for antenna in self.laser_antennas:           # main.py:491-494
                antenna.update_v(self.time + 0.5 * dt, dt)
                antenna.push_x(0.5 * dt)
            if self.use_galilean:
                self.shift_galilean_boundaries(0.5 * dt)
            # elementary processes at t = (n + 1/2) dt, positions and momenta synchronised (main.py:499-503)
            for species in ptcl:
                species.handle_elementary_processes(self.time + 0.5 * dt)
            for species in ptcl:
                species.keep_fields_sorted = False

            cross = correct_currents and fld.current_correction == 'cross-deposition'
            # fused mode: nothing reads the spectral current before correct_currents, so its FFTs and Hankel
            # transforms wait for the charge deposited at the end of the step and share its launches
            defer_J = self.fused and not cross and not self.use_pml and \
                not (self.comm.size > 1 and ((correct_currents is False) or (use_true_rho is True)))
            self.deposit('J', exchange=(correct_currents is False), defer_spectral=defer_J)
            if cross:                                 # main.py:512-514
                self.cross_deposit(move_positions)
            # fused mode: the second half push rides inside the rho deposition kernel when every
            # species deposits and already has sort locality
            fuse_pr = self.fused and move_positions and len(ptcl) > 0 and (not cross) and \
                all((sp.q != 0) and (not sp.is_tracer) and getattr(sp, '_order_matches_prefix', False)
                    and sp.ionizer is None for sp in ptcl)
            for antenna in self.laser_antennas:           # main.py:520-522
                antenna.push_x(0.5 * dt)
            if fuse_pr:
                if self.use_galilean:
                    self.shift_galilean_boundaries(0.5 * dt)
                z0 = fld.interp[0].zmin
                wrap = (z0, fld.interp[0].zmax) if (wrap_in_push and i_step < N - 1) else None
                self.deposit('rho_next', exchange=(use_true_rho is True), push=(0.5 * dt, wrap))
            elif move_positions:
                if self.fused:
                    z0 = fld.interp[0].zmin + gal_shift
                    # (no wrap after the last step of this call: the reference leaves x^{n+1} unwrapped until
                    #  the exchange_particles at the start of the next step, main.py:435-442)
                    wrap = (z0, z0 + (fld.interp[0].zmax - fld.interp[0].zmin)) \
                        if (wrap_in_push and i_step < N - 1) else None
                    kz0 = None          # rho is deposited without a re-sort: no keys needed here
                    for species in ptcl:
                        species.push_x_and_key(0.5 * dt, fld, wrap=wrap, key_zmin=kz0)
                else:
                    for species in ptcl:
                        species.push_x(0.5 * dt)
            if not fuse_pr:
                if self.use_galilean:
                    self.shift_galilean_boundaries(0.5 * dt)
                self.deposit('rho_next', exchange=(use_true_rho is True))

            if fuse_cp:
                fld.correct_currents_and_push(use_true_rho)
                fld.exchanged_source['J'] = True
            else:
                if correct_currents:
                    fld.correct_currents(check_exchanges=(self.comm.size > 1))
                    if self.comm.size > 1:
                        fld.spect2partial_interp('J')
                        self.comm.exchange_fields(fld.interp, 'J', 'add')
                        fld.partial_interp2spect('J')
                    fld.exchanged_source['J'] = True
                fld.push(use_true_rho, check_exchanges=(self.comm.size > 1))
            if correct_divE:                          # main.py:543-544
                fld.correct_divE()
            if self.comm.moving_win is not None:
                self.comm.move_grids(fld, ptcl, dt, self.time)
            self.exchange_and_damp_EB(skip_identity=(periodic_single and self.fused and not self.use_pml
                                                     and not self.mirrors))
            self.time += dt
            self.iteration += 1
            for checkpoint in self.checkpoints:       # main.py:563-565
                checkpoint.write(self.iteration)

        fld.spect2interp('J')
        if (not fld.exchanged_source['J']) and (self.comm.size > 1):
            self.comm.exchange_fields(self.fld.interp, 'J', 'add')
        fld.spect2interp('rho_prev')
        if (not fld.exchanged_source['rho_prev']) and (self.comm.size > 1):
            self.comm.exchange_fields(self.fld.interp, 'rho', 'add')
        t_enqueued = _time.perf_counter()
        _lib.context().sync()
        t_done = _time.perf_counter()
        if not keep_on_gpu:
            self.receive_data_from_gpu()
        # wall-clock split of this call: host->device copy, the N cycles (device-synchronised), device->host copy
        self.last_step_timing = dict(h2d_s=t_sent - t_start, cycles_s=t_done - t_sent,
                                     d2h_s=_time.perf_counter() - t_done)
        # host time spent issuing the launches of the N cycles (the device runs behind it, asynchronously)
        self.last_enqueue_s = t_enqueued - t_sent
        if progress_bar is not None:
            progress_bar.print_summary()
        # bytes copied host->device / device->host by this call (counted from the arrays actually copied)
        self.last_step_bytes = {k: _lib.TRANSFERRED[k] - bytes0[k] for k in bytes0}

    def deposit(self, fieldtype, exchange=False, update_spectral=True, species_list=None, push=None,
                defer_spectral=False):
        """fbpic/main.py:588-670.  `defer_spectral` (fused mode): the transforms of this source are batched
        with those of the next fused deposit."""
        fld = self.fld
        if species_list is None:            # everything deposits (main.py:618-624)
            species_list = [s for s in self.ptcl if not s.is_tracer]
            antennas_list = self.laser_antennas
        else:
            antennas_list = []
        if fieldtype.startswith('rho'):
            grid_type = 'rho'
        elif fieldtype == 'J':
            grid_type = 'J'
        else:
            raise ValueError('Unknown fieldtype: %s' % fieldtype)
        fld.erase(grid_type)
        for species in species_list:
            if self.fused:
                species.deposit_fused(fld, grid_type, push=push)
            else:
                species.deposit(fld, grid_type)
        for antenna in antennas_list:       # main.py:634-636, 651-653
            antenna.deposit(fld, grid_type)
        fld.sum_reduce_deposition_array(grid_type)
        if self.fused and update_spectral and not (exchange and self.comm.size > 1):
            # divide_by_volume, the transforms and the filter as FFTs + one batched Hankel launch
            pending = getattr(self, '_pending_spect', [])
            if defer_spectral:
                self._pending_spect = pending + [fieldtype]
            else:
                self._pending_spect = []
                fld.fused_deposit2spect(pending + [fieldtype], self.filter_currents)
            fld.exchanged_source[fieldtype] = exchange
            return
        assert not getattr(self, '_pending_spect', []), 'a deferred transform must be followed by a fused deposit'
        fld.divide_by_volume(grid_type)
        if exchange and self.comm.size > 1:
            self.comm.exchange_fields(fld.interp, grid_type, 'add')
        if update_spectral:
            fld.interp2spect(fieldtype)
            if self.filter_currents:
                fld.filter_spect(fieldtype)
            fld.exchanged_source[fieldtype] = exchange

    def exchange_and_damp_EB(self, skip_identity=False):
        """fbpic/main.py:719-769.  On a single periodic domain the iFFT / exchange / FFT round
        trip is an identity (no neighbour, no damping): `skip_identity` drops those 12 FFTs per
        mode and goes straight to spect2interp."""
        fld = self.fld
        if self.use_pml:
            # exchange / damp act in z AND r: full transforms both ways (main.py:732-761); fused mode batches
            # them (two Hankel launches + one multi-lane FFT call each way instead of 10 launches per mode)
            if self.fused:
                fld.fused_spect2interp_EB_pml()
            else:
                for ft in ('E', 'B', 'E_pml', 'B_pml'):
                    fld.spect2interp(ft)
            self.comm.exchange_fields(fld.interp, 'E', 'replace')
            self.comm.exchange_fields(fld.interp, 'B', 'replace')
            self.comm.damp_EB_open_boundary(fld.interp)
            self.comm.damp_pml_EB(fld.interp)
            for mirror in self.mirrors:
                mirror.set_fields_to_zero(fld.interp, self.comm, self.time)
            if self.fused:
                fld.fused_interp2spect_EB_pml()
            else:
                for ft in ('E', 'B', 'E_pml', 'B_pml'):
                    fld.interp2spect(ft)
            return
        if not skip_identity:
            if self.fused:
                # E and B together: one batch of concurrent z-FFTs each way, one NCCL group
                fld.spect2partial_interp('EB')
                self.comm.exchange_fields(fld.interp, 'EB', 'replace')
                self.comm.damp_EB_open_boundary(fld.interp)
                for mirror in self.mirrors:           # main.py:751-753 (rows in z: valid in (z, kr) space)
                    mirror.set_fields_to_zero(fld.interp, self.comm, self.time)
                # (the spectral arrays are next read by the field push of the following cycle: these 6*Nm forward
                #  transforms go to the second stream and run under that cycle's particle kernels)
                fld.partial_interp2spect('EB', side=True)
                # the exchanged (z, kr) arrays go straight to real space: inverse Hankel only
                fld.fused_partial2interp_EB()
                return
            fld.spect2partial_interp('E')
            fld.spect2partial_interp('B')
            self.comm.exchange_fields(fld.interp, 'E', 'replace')
            self.comm.exchange_fields(fld.interp, 'B', 'replace')
            self.comm.damp_EB_open_boundary(fld.interp)
            for mirror in self.mirrors:
                mirror.set_fields_to_zero(fld.interp, self.comm, self.time)
            fld.partial_interp2spect('E')
            fld.partial_interp2spect('B')
        if self.fused:
            fld.fused_spect2interp_EB()
        else:
            fld.spect2interp('E')
            fld.spect2interp('B')

    def cross_deposit(self, move_positions):
        """fbpic/main.py:672-717: with the particles at time n+1/2, deposit rho at (z[n], x[n+1]) and at
        (z[n+1], x[n]) for the cross-deposition current correction."""
        dt = self.dt
        for frac, sx, sz, ft in ((0.5, 1., -1., 'rho_next_xy'), (1., -1., 1., 'rho_next_z'), (0.5, 1., -1., None)):
            if move_positions:
                for species in self.ptcl:
                    species.push_x(frac * dt, x_push=sx, y_push=sx, z_push=sz)
            for antenna in self.laser_antennas:
                antenna.push_x(frac * dt, x_push=sx, y_push=sx, z_push=sz)
            if self.use_galilean:
                self.shift_galilean_boundaries(sz * frac * dt)
            if ft is not None:
                self.deposit(ft)

    def shift_galilean_boundaries(self, dt):
        """fbpic/main.py:772-789"""
        shift_distance = self.v_comoving * dt
        self.comm.shift_global_domain_positions(shift_distance)
        for m in range(self.fld.Nm):
            self.fld.interp[m].zmin += shift_distance
            self.fld.interp[m].zmax += shift_distance

    # ------------------------------------------------------------------ species
    def add_new_species(self, q, m, n=None, dens_func=None, p_nz=None, p_nr=None, p_nt=None,
                        p_zmin=-np.inf, p_zmax=np.inf, p_rmin=0, p_rmax=np.inf,
                        uz_m=0., ux_m=0., uy_m=0., uz_th=0., ux_th=0., uy_th=0.,
                        continuous_injection=True, boost_positions_in_dens_func=False, is_tracer=False):
        """fbpic/main.py:792-1001.  With `gamma_boost`, positions, density and momenta are given in the
        lab frame and converted to the boosted frame here (main.py:909-950)."""
        if n is not None:
            for var in (p_nz, p_nr, p_nt):
                if var is None:
                    raise ValueError('If the density `n` is passed to `add_new_species`,\n'
                                     'then the arguments `p_nz`, `p_nr` and `p_nt` need to be passed too.')
            if self.boost is not None:
                gamma_m = np.sqrt(1. + uz_m**2 + ux_m**2 + uy_m**2)
                beta_m_lab = uz_m / gamma_m
                p_zmin, p_zmax = self.boost.copropag_length([p_zmin, p_zmax], beta_object=beta_m_lab)
                n, = self.boost.copropag_density([n], beta_object=beta_m_lab)
                # approximate transform of the longitudinal thermal spread (perturbation of the Lorentz
                # transform of uz), then of the mean momentum
                if uz_m == 0:
                    uz_th = self.boost.gamma0 * uz_th
                else:
                    uz_th = self.boost.gamma0 * (1. - self.boost.beta0 * beta_m_lab) * uz_th
                uz_m = self.boost.gamma0 * (uz_m - self.boost.beta0 * gamma_m)
                if boost_positions_in_dens_func and (dens_func is not None):
                    from .particles import _dens_func_args
                    coef = self.boost.gamma0 * (1 - beta_m_lab * self.boost.beta0)
                    lab_dens_func = dens_func
                    if _dens_func_args(lab_dens_func) == ['z', 'r']:
                        dens_func = lambda z, r: lab_dens_func(coef * z, r)      # noqa: E731
                    else:
                        dens_func = lambda x, y, z: lab_dens_func(x, y, coef * z)  # noqa: E731
            zmin_local, zmax_local = self.comm.get_zmin_zmax(local=True, rank=self.comm.rank,
                                                            with_damp=False, with_guard=False)
            p_zmin = max(zmin_local, p_zmin)
            p_zmax = min(zmax_local, p_zmax)
            p_rmax = min(self.comm.get_rmax(with_damp=False), p_rmax)
            p_zmin, p_zmax, Npz = adapt_to_grid(self.fld.interp[0].z, p_zmin, p_zmax, p_nz)
            p_rmin, p_rmax, Npr = adapt_to_grid(self.fld.interp[0].r, p_rmin, p_rmax, p_nr)
            dz_particles = self.comm.dz / p_nz
        else:
            n = 0
            p_zmin = p_zmax = p_rmin = p_rmax = 0
            Npz = Npr = p_nt = 0
            continuous_injection = False
            dz_particles = 0.
        sp = Particles(q=q, m=m, n=n, dens_func=dens_func, Npz=Npz, zmin=p_zmin, zmax=p_zmax,
                       Npr=Npr, rmin=p_rmin, rmax=p_rmax, Nptheta=p_nt, dt=self.dt,
                       particle_shape=self.particle_shape, grid_shape=self.grid_shape,
                       ux_m=ux_m, uy_m=uy_m, uz_m=uz_m, ux_th=ux_th, uy_th=uy_th, uz_th=uz_th,
                       continuous_injection=continuous_injection, dz_particles=dz_particles,
                       is_tracer=is_tracer)
        sp.sort_period = self.sort_period
        self.ptcl.append(sp)
        return sp

    def set_moving_window(self, v=c, **kw):
        """fbpic/main.py:1004-1032 (the deprecated keyword arguments are accepted and ignored)."""
        from .moving_window import MovingWindow
        self.comm.moving_win = MovingWindow(self.comm, self.dt, v, self.time)


    def reverse_time(self):
        """Reverse the propagation direction of waves and particles: B and the momenta change sign
        (fbpic/main.py:1034-1053).  Acts on the host copy of the data."""
        if self.fld.data_is_on_gpu or any(sp.data_is_on_gpu for sp in self.ptcl):
            raise _lib.B200Error('reverse_time acts on the host copy of the data: call receive_data_from_gpu() first')
        for m in range(self.fld.Nm):
            for k in ('Bp', 'Bm', 'Bz'):
                getattr(self.fld.spect[m], k)[...] *= -1
            for k in ('Br', 'Bt', 'Bz'):
                getattr(self.fld.interp[m], k)[...] *= -1
        for species in self.ptcl:
            for k in ('ux', 'uy', 'uz'):
                setattr(species, k, -np.asarray(getattr(species, k)))


class GpuMemoryManager(object):
    """`with GpuMemoryManager(sim):` -- fields and particles live in HBM inside the block and come back as
    NumPy arrays at its end (fbpic/utils/cuda.py:139-182)."""

    def __init__(self, sim):
        self.sim = sim

    def __enter__(self):
        self.sim.send_data_to_gpu()

    def __exit__(self, type, value, traceback):
        self.sim.receive_data_from_gpu()


def adapt_to_grid(x, p_xmin, p_xmax, p_nx, ncells_empty=0):
    """Snap particle bounds to the grid and count the particles (fbpic/main.py:1056-1111)."""
    xmin, xmax = x.min(), x.max()
    dx = x[1] - x[0]
    if p_xmin < xmin - 0.5 * dx:
        p_xmin = xmin - 0.5 * dx
    if p_xmax > xmax + (0.5 - ncells_empty) * dx:
        p_xmax = xmax + (0.5 - ncells_empty) * dx
    x_load = x[(x > p_xmin) & (x < p_xmax)]
    Npx = len(x_load) * p_nx
    if Npx > 0:
        p_xmin = x_load.min() - 0.5 * dx
        p_xmax = x_load.max() + 0.5 * dx
    return p_xmin, p_xmax, Npx

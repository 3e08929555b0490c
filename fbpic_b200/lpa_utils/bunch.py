"""
Relativistic particle bunches and their initial space-charge field, same entry points as
`fbpic.lpa_utils.bunch` (fbpic/lpa_utils/bunch.py:18-1007): `add_particle_bunch`, `add_particle_bunch_gaussian`,
`add_particle_bunch_file`, `add_particle_bunch_from_arrays`, the `add_elec_bunch*` shorthands and
`get_space_charge_fields`.

The bunch charge and current are deposited by the regular deposition kernel; the Poisson-like solve for the
field of a bunch of Lorentz factor gamma is done in spectral space -- forward and inverse transforms on the
GPU (cuFFT + DMMA Hankel), the element-wise spectral formula on the host arrays of this one-off set-up (the
reference runs all of it on the CPU, bunch.py:886-915).
"""
import warnings
import numpy as np
from scipy.constants import c, e, m_e, epsilon_0, mu_0


def _set_injection_plane(ptcl_bunch, z_injection_plane, boost):
    """bunch.py:117-119: ballistic motion before the plane z_injection_plane (lab frame)"""
    if z_injection_plane is not None:
        from ..particles import BallisticBeforePlane
        assert ptcl_bunch.injector is None      # don't overwrite a previous injector
        ptcl_bunch.injector = BallisticBeforePlane(z_injection_plane, boost)


def add_particle_bunch(sim, q, m, gamma0, n, p_zmin, p_zmax, p_rmin, p_rmax, p_nr=2, p_nz=2, p_nt=4,
                       dens_func=None, boost=None, direction='forward', z_injection_plane=None,
                       initialize_self_field=True, boost_positions_in_dens_func=False):
    """Uniform (or dens_func-shaped) mono-energetic bunch of Lorentz factor gamma0 (bunch.py:18-123)."""
    uz_m = (gamma0**2 - 1.)**0.5
    if direction == 'backward':
        uz_m *= -1.
    ptcl_bunch = sim.add_new_species(q=q, m=m, n=n, p_nz=p_nz, p_nr=p_nr, p_nt=p_nt, p_zmin=p_zmin, p_zmax=p_zmax,
                                     p_rmin=p_rmin, p_rmax=p_rmax, continuous_injection=False,
                                     dens_func=dens_func, uz_m=uz_m,
                                     boost_positions_in_dens_func=boost_positions_in_dens_func)
    _set_injection_plane(ptcl_bunch, z_injection_plane, boost)
    if initialize_self_field:
        get_space_charge_fields(sim, ptcl_bunch, direction=direction)
    return ptcl_bunch


def add_particle_bunch_gaussian(sim, q, m, sig_r, sig_z, n_emit, gamma0, sig_gamma, n_physical_particles,
                                n_macroparticles, tf=0., zf=0., boost=None, save_beam=None,
                                z_injection_plane=None, initialize_self_field=True, symmetrize=False):
    """Gaussian bunch with normalised emittance n_emit, focused at zf at time tf (bunch.py:126-284; same
    order of draws from np.random, so the same seed gives the same bunch as the reference)."""
    if symmetrize:
        assert n_macroparticles % 4 == 0, "When using symmetrize, `n_macroparticles` must be a multiple of 4."
        n_macroparticles = n_macroparticles // 4
    if sig_gamma > 0.:
        gamma = np.random.normal(gamma0, sig_gamma, n_macroparticles)
    else:
        gamma = np.full(n_macroparticles, gamma0)
        if sig_gamma < 0.:
            warnings.warn("Negative energy spread sig_gamma detected. sig_gamma will be set to zero. \n")
    inv_gamma = 1. / gamma
    x = sig_r * np.random.normal(0., 1., n_macroparticles)
    y = sig_r * np.random.normal(0., 1., n_macroparticles)
    z = zf + sig_z * np.random.normal(0., 1., n_macroparticles)
    sig_ur = n_emit / sig_r
    ux = sig_ur * np.random.normal(0., 1., n_macroparticles)
    uy = sig_ur * np.random.normal(0., 1., n_macroparticles)
    uz_sqr = (gamma**2 - 1) - ux**2 - uy**2
    keep = uz_sqr >= 0
    N_new = np.count_nonzero(keep)
    if N_new < n_macroparticles:
        warnings.warn("Particles with uz**2<0 detected. %d Particles will be removed from the beam. \n"
                      "However, the charge will be kept constant. \n" % (n_macroparticles - N_new))
        x, y, z, ux, uy, inv_gamma, uz_sqr = [a[keep] for a in (x, y, z, ux, uy, inv_gamma, uz_sqr)]
    uz = np.sqrt(uz_sqr)
    w = n_physical_particles / N_new * np.ones_like(x)
    if tf != 0.:     # ballistic back-propagation from the focus
        x = x - ux * inv_gamma * c * tf
        y = y - uy * inv_gamma * c * tf
        z = z - uz * inv_gamma * c * tf
    if symmetrize:   # 4-fold rotational symmetry: zero initial offset in x and y
        w *= 0.25
        x, y, z, ux, uy, uz, w = map(np.concatenate, zip([x, y, z, ux, uy, uz, w], [-y, x, z, -uy, ux, uz, w],
                                                         [-x, -y, z, -ux, -uy, uz, w], [y, -x, z, uy, -ux, uz, w]))
    if save_beam is not None:
        np.savez(save_beam, x=x, y=y, z=z, ux=ux, uy=uy, uz=uz, inv_gamma=inv_gamma, w=w)
    return add_particle_bunch_from_arrays(sim, q, m, x, y, z, ux, uy, uz, w, boost=boost,
                                          z_injection_plane=z_injection_plane,
                                          initialize_self_field=initialize_self_field)


def add_particle_bunch_file(sim, q, m, filename, n_physical_particles, z_off=0., boost=None, direction='forward',
                            z_injection_plane=None, initialize_self_field=True):
    """Bunch from a text file with the columns x y z ux uy uz (bunch.py:287-356)."""
    data = np.loadtxt(filename)
    x, y, z = data[:, 0], data[:, 1], data[:, 2] + z_off
    ux, uy, uz = data[:, 3], data[:, 4], data[:, 5]
    w = n_physical_particles / len(x) * np.ones_like(x)
    return add_particle_bunch_from_arrays(sim, q, m, x, y, z, ux, uy, uz, w, boost=boost, direction=direction,
                                          z_injection_plane=z_injection_plane,
                                          initialize_self_field=initialize_self_field)


def add_particle_bunch_openPMD(sim, q, m, ts_path, z_off=0., species=None, select=None, iteration=None, boost=None,
                               z_injection_plane=None, initialize_self_field=True):
    """Bunch from the particle record of an openPMD series (bunch.py:359-453): positions, momenta (kg m/s in the
    file, divided by the mass record and c) and weights of `species` at `iteration` (default: the last one),
    optionally restricted by `select` = {'uz': [lo, hi], ...}; the bunch is re-centred so that its weighted mean z is
    `z_off`.  `ts_path` is the directory that holds the `data%08d.h5` files -- or the `.npz` archives that
    fbpic_b200's diagnostics write when h5py is not installed.  The reference reads the series with openPMD-viewer;
    here the tree is read directly (fbpic_b200/openpmd_store.py)."""
    import os
    from scipy.constants import c
    from ..diags import read_diag, list_iterations
    write_dir = os.path.dirname(os.path.abspath(ts_path).rstrip('/')) if os.path.basename(
        os.path.abspath(ts_path).rstrip('/')) == 'hdf5' else ts_path
    its = list_iterations(write_dir)
    if not its:
        raise OSError('No openPMD file (data%%08d.h5 / .npz) found in %s' % ts_path)
    if iteration is None:
        iteration = its[-1]
    d = read_diag(write_dir, iteration)
    names = sorted({k.split('/')[1].split('@')[0] for k in d if k.startswith('particles/')})
    if species is None:
        if len(names) != 1:
            raise ValueError('Several species in the file (%s): pass `species`.' % ', '.join(names))
        species = names[0]
    grp = 'particles/%s/' % species
    if grp + 'position/x' not in d:
        raise ValueError('The file holds no species `%s` (available: %s).' % (species, ', '.join(names)))
    mass = float(d[grp + 'mass@value'])
    to_u = 1. / (mass * c) if mass > 0 else 1.
    data = {'x': d[grp + 'position/x'], 'y': d[grp + 'position/y'], 'z': d[grp + 'position/z'],
            'ux': d[grp + 'momentum/x'] * to_u, 'uy': d[grp + 'momentum/y'] * to_u, 'uz': d[grp + 'momentum/z'] * to_u,
            'w': d[grp + 'weighting']}
    if select is not None:
        keep = np.ones(len(data['w']), dtype=bool)
        for quantity, (lo, hi) in select.items():
            v = np.sqrt(1. + data['ux']**2 + data['uy']**2 + data['uz']**2) if quantity == 'gamma' else data[quantity]
            if lo is not None:
                keep &= v > lo
            if hi is not None:
                keep &= v < hi
        data = {k: v[keep] for k, v in data.items()}
    z = data['z'] - np.average(data['z'], weights=data['w']) + z_off
    return add_particle_bunch_from_arrays(sim, q, m, data['x'], data['y'], z, data['ux'], data['uy'], data['uz'],
                                          data['w'], boost=boost, z_injection_plane=z_injection_plane,
                                          initialize_self_field=initialize_self_field)


def add_elec_bunch_openPMD(sim, ts_path, z_off=0., species=None, select=None, iteration=None, boost=None,
                           z_injection_plane=None):
    """bunch.py:742-793"""
    return add_particle_bunch_openPMD(sim, -e, m_e, ts_path, z_off=z_off, species=species, select=select,
                                      iteration=iteration, boost=boost, z_injection_plane=z_injection_plane)


def add_particle_bunch_from_arrays(sim, q, m, x, y, z, ux, uy, uz, w, boost=None, direction='forward',
                                   z_injection_plane=None, initialize_self_field=True):
    """Bunch from arrays given in the lab frame (bunch.py:456-547); with `boost` they are Lorentz-transformed
    and propagated to t' = 0.  Particles outside the local physical domain are dropped."""
    from ..particles import FIELD_ATTRS
    inv_gamma = 1. / np.sqrt(1. + ux**2 + uy**2 + uz**2)
    if boost is not None:
        x, y, z, ux, uy, uz, inv_gamma = boost.boost_particle_arrays(x, y, z, ux, uy, uz, inv_gamma)
    zmin, zmax = sim.comm.get_zmin_zmax(local=True, with_damp=False, with_guard=False, rank=sim.comm.rank)
    sel = (z >= zmin) & (z < zmax)
    ptcl_bunch = sim.add_new_species(q=q, m=m)
    for k, a in (('x', x), ('y', y), ('z', z), ('ux', ux), ('uy', uy), ('uz', uz), ('inv_gamma', inv_gamma),
                 ('w', w)):
        setattr(ptcl_bunch, k, np.ascontiguousarray(np.asarray(a, dtype=np.float64)[sel]))
    ptcl_bunch.Ntot = int(sel.sum())
    for k in FIELD_ATTRS:
        setattr(ptcl_bunch, k, np.zeros(ptcl_bunch.Ntot))
    _set_injection_plane(ptcl_bunch, z_injection_plane, boost)
    if initialize_self_field:
        get_space_charge_fields(sim, ptcl_bunch, direction=direction)
    return ptcl_bunch


# ---- electron shorthands (bunch.py:550-835) ----
def add_elec_bunch(sim, gamma0, n_e, p_zmin, p_zmax, p_rmin, p_rmax, p_nr=2, p_nz=2, p_nt=4, dens_func=None,
                   boost=None, direction='forward', z_injection_plane=None):
    return add_particle_bunch(sim, -e, m_e, gamma0, n_e, p_zmin, p_zmax, p_rmin, p_rmax, p_nr=p_nr, p_nz=p_nz,
                              p_nt=p_nt, dens_func=dens_func, boost=boost, direction=direction,
                              z_injection_plane=z_injection_plane)


def add_elec_bunch_gaussian(sim, sig_r, sig_z, n_emit, gamma0, sig_gamma, Q, N, tf=0., zf=0., boost=None,
                            save_beam=None, z_injection_plane=None, symmetrize=False):
    return add_particle_bunch_gaussian(sim, -e, m_e, sig_r, sig_z, n_emit, gamma0, sig_gamma, Q / e, N, tf=tf,
                                       zf=zf, boost=boost, save_beam=save_beam,
                                       z_injection_plane=z_injection_plane, symmetrize=symmetrize)


def add_elec_bunch_file(sim, filename, Q_tot, z_off=0., boost=None, direction='forward', z_injection_plane=None):
    return add_particle_bunch_file(sim, -e, m_e, filename, Q_tot / e, z_off=z_off, boost=boost,
                                   direction=direction, z_injection_plane=z_injection_plane)


def add_elec_bunch_from_arrays(sim, x, y, z, ux, uy, uz, w, boost=None, direction='forward',
                               z_injection_plane=None):
    return add_particle_bunch_from_arrays(sim, -e, m_e, x, y, z, ux, uy, uz, w, boost=boost, direction=direction,
                                          z_injection_plane=z_injection_plane)


# ---- space charge (bunch.py:838-1007) ----
def get_space_charge_spect(spect, gamma, direction='forward', neglect_transverse_currents=True):
    """Field of a charge / current distribution moving rigidly at Lorentz factor gamma, in spectral space:
    phi = rho / (eps0 K^2), A = mu0 J / K^2 with K^2 = kr^2 + kz^2 / gamma^2; E = -grad phi - d_t A,
    B = curl A (bunch.py:946-1007).  Element-wise, on the host arrays of `spect`."""
    beta = np.sqrt(1. - 1. / gamma**2)
    if direction == 'backward':
        beta *= -1.
    kz, kr = spect.kz, spect.kr
    K2 = kr**2 + kz**2 * 1. / gamma**2
    inv_K2 = np.where(K2 != 0, 1. / np.where(K2 != 0, K2, 1.), 0.)
    phi = spect.rho_prev[:, :] * inv_K2 / epsilon_0
    Az = spect.Jz[:, :] * inv_K2 * mu_0
    spect.Ep[:, :] += 0.5 * kr * phi
    spect.Em[:, :] += -0.5 * kr * phi
    spect.Ez[:, :] += -1.j * kz * phi + 1.j * beta * c * kz * Az
    spect.Bp[:, :] += -0.5j * kr * Az
    spect.Bm[:, :] += -0.5j * kr * Az
    if not neglect_transverse_currents:
        Ap = spect.Jp[:, :] * inv_K2 * mu_0
        Am = spect.Jm[:, :] * inv_K2 * mu_0
        spect.Ep[:, :] += 1.j * beta * c * kz * Ap
        spect.Em[:, :] += 1.j * beta * c * kz * Am
        spect.Bp[:, :] += kz * Ap
        spect.Bm[:, :] -= kz * Am
        spect.Bz[:, :] += 1.j * kr * Ap + 1.j * kr * Am


def get_space_charge_fields(sim, ptcl, direction='forward'):
    """Add the space-charge field of the relativistic species `ptcl` (all particles at about the same gamma)
    to the interpolation grids of `sim` (bunch.py:838-944)."""
    from ..fields import Fields
    comm, fld = sim.comm, sim.fld
    if ptcl.data_is_on_gpu or fld.data_is_on_gpu:
        raise RuntimeError('get_space_charge_fields acts on the host copy of the data: call it before step() '
                           'or after receive_data_from_gpu()')
    w_sum, w_gamma_sum = comm.allreduce_sum([float(np.sum(ptcl.w)), float(np.sum(ptcl.w * 1. / ptcl.inv_gamma))])
    if w_sum == 0:
        warnings.warn("Tried to calculate space charge, but found 0 macroparticles in \n"
                      "the corresponding species. Skipping space charge calculation...\n")
        return
    gamma = w_gamma_sum / w_sum
    # charge and current of the bunch on the local grid, through the regular deposition kernel
    sim.send_data_to_gpu()
    sim.deposit('rho', exchange=True, species_list=[ptcl], update_spectral=False)
    sim.deposit('J', exchange=True, species_list=[ptcl], update_spectral=False)
    sim.receive_data_from_gpu()
    # the same on the global grid (damp cells included, guard cells not)
    Nz_g, _ = comm.get_Nz_and_iz(local=False, with_damp=True, with_guard=False)
    zmin_g, zmax_g = comm.get_zmin_zmax(local=False, with_damp=True, with_guard=False)
    gfld = Fields(Nz_g, zmax_g, fld.Nr, fld.rmax, fld.Nm, fld.dt, n_order=fld.n_order, smoother=fld.smoother,
                  zmin=zmin_g)
    for m in range(fld.Nm):
        for k in ('Jr', 'Jt', 'Jz', 'rho'):
            getattr(gfld.interp[m], k)[:, :] = comm.gather_grid_array(getattr(fld.interp[m], k), with_damp=True)
    gfld.send_fields_to_gpu()
    gfld.interp2spect('rho_prev')
    gfld.interp2spect('J')
    if sim.filter_currents:
        gfld.filter_spect('rho_prev')
        gfld.filter_spect('J')
    gfld.receive_fields_from_gpu()
    for m in range(gfld.Nm):
        get_space_charge_spect(gfld.spect[m], gamma, direction)
    gfld.send_fields_to_gpu()
    gfld.spect2interp('E')
    gfld.spect2interp('B')
    gfld.receive_fields_from_gpu()
    Nz_loc, iz_dom = comm.get_Nz_and_iz(local=True, with_damp=True, with_guard=False, rank=comm.rank)
    _, iz_arr = comm.get_Nz_and_iz(local=True, with_damp=True, with_guard=True, rank=comm.rank)
    i_loc = iz_dom - iz_arr
    for m in range(fld.Nm):
        for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz'):
            getattr(fld.interp[m], k)[i_loc:i_loc + Nz_loc, :] += \
                comm.scatter_grid_array(getattr(gfld.interp[m], k), with_damp=True)

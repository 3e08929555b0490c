"""
oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy + the OpenMP C kernels of oracle/pic_oracle.c) of the
per-step PIC hot loop of the FBPIC reference, single z-periodic domain:
gather -> Vay push -> push_x -> deposit J / rho -> FFT+Hankel -> current
correction -> PSATD push -> inverse transforms.  It is the checker for the CUDA
path of `fbpic_b200` and the timed CPU baseline of bench.py; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product never does.

Parity pinned: every stage is compared with outputs of the unmodified
reference (imported in the build container through oracle/ref_shim) by
oracle/gen_golden.py -> tests/golden/*.npz -> tests/test_oracle_golden.py.

The one-off host tables (Hankel matrices, PSATD coefficients, filters, Ruyten
coefficients) are the product's `fbpic_b200.host_tables`; they are pinned
bit-exactly against the reference by the same golden files.

Reference call order followed here: fbpic/main.py:346-586 (step), :588-670
(deposit); fbpic/fields/fields.py:247-625; fbpic/fields/numba_methods.py;
fbpic/fields/spectral_transform/spectral_transformer.py:89-223.
"""
import ctypes
import os
import subprocess
import numpy as np
import scipy.fft as sfft
from scipy.constants import c, mu_0, epsilon_0

from fbpic_b200 import host_tables as ht

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc + OpenMP)."""
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'pic_oracle.c')
    if force or (not os.path.exists(so)) or \
            (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_max_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_d, _i, _l = ctypes.c_double, ctypes.c_int, ctypes.c_int64


def nthreads_default():
    return int(os.environ.get('ORACLE_NUM_THREADS', lib().orc_max_threads()))


# ---------------------------------------------------------------------------
# particle kernels (thin ctypes wrappers)
# ---------------------------------------------------------------------------
def cell_index(x, y, z, invdz, zmin, Nz, invdr, rmin, Nr):
    out = np.empty(len(x), dtype=np.int32)
    lib().orc_cell_index(_l(len(x)), _p(x), _p(y), _p(z), _d(invdz), _d(zmin), _i(Nz),
                         _d(invdr), _d(rmin), _i(Nr), _p(out))
    return out


def sort_contract(cell_idx, Nz, Nr):
    """The sorting contract of the build: stable argsort by cell key and
    the inclusive per-cell prefix sum (cuda_sorting.py:91-190)."""
    sorted_idx = np.argsort(cell_idx, kind='stable').astype(np.int64)
    prefix = np.cumsum(np.bincount(cell_idx, minlength=Nz * (Nr + 1))).astype(np.int32)
    return sorted_idx, prefix


def push_p(ux, uy, uz, inv_gamma, Ex, Ey, Ez, Bx, By, Bz, q, m, dt):
    lib().orc_push_p(_l(len(ux)), _p(ux), _p(uy), _p(uz), _p(inv_gamma), _p(Ex), _p(Ey), _p(Ez),
                     _p(Bx), _p(By), _p(Bz), _d(q), _d(m), _d(dt))


def push_x(x, y, z, ux, uy, uz, inv_gamma, dt, x_push=1., y_push=1., z_push=1.):
    lib().orc_push_x(_l(len(x)), _p(x), _p(y), _p(z), _p(ux), _p(uy), _p(uz), _p(inv_gamma),
                     _d(dt), _d(x_push), _d(y_push), _d(z_push))


def shift_periodic(z, zmin, zmax):
    lib().orc_shift_periodic(_l(len(z)), _p(z), _d(zmin), _d(zmax))


def gather(x, y, z, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, grids, cubic,
           Ex, Ey, Ez, Bx, By, Bz):
    """`grids`: list over modes of (Er, Et, Ez, Br, Bt, Bz) complex [Nz, Nr] arrays."""
    Nm = len(grids)
    flat = [np.ascontiguousarray(a) for g in grids for a in g]
    ptrs = (ctypes.c_void_p * (6 * Nm))(*[a.ctypes.data for a in flat])
    lib().orc_gather(_l(len(x)), _p(x), _p(y), _p(z), _d(rmax_gather), _d(invdz), _d(zmin), _i(Nz),
                     _d(invdr), _d(rmin), _i(Nr), _i(Nm), ptrs, _i(int(cubic)),
                     _p(Ex), _p(Ey), _p(Ez), _p(Bx), _p(By), _p(Bz))


_DEP_BUFFERS = {}


def deposit(what, x, y, z, w, q, ux, uy, uz, inv_gamma, invdz, zmin, Nz, invdr, rmin, Nr, Nm,
            cubic, beta0, beta_hi, nthreads=None):
    """Deposit 'rho' or 'J'; returns raw (not volume-divided) complex sums,
    shape [ncomp, Nm, Nz, Nr] (comp order rho | Jr, Jt, Jz), guard cells folded
    as in fbpic/fields/numba_methods.py:410-461."""
    nthreads = nthreads or nthreads_default()
    ncomp = 3 if what == 'J' else 1
    # the per-thread guarded copies persist between calls and are erased in parallel, like
    # Fields.rho_global / J*_global + numba_erase_threading_buffer (fields.py:205-218, 539-564)
    shape = (nthreads, ncomp, Nm, Nz + 4, Nr + 4)
    glob = _DEP_BUFFERS.get(shape)
    if glob is None:
        _DEP_BUFFERS.clear()
        glob = _DEP_BUFFERS.setdefault(shape, np.zeros(shape, dtype=np.complex128))
    else:
        lib().orc_zero(_p(glob), _l(glob.size * 2))
    dummy = x
    lib().orc_deposit(_i(1 if what == 'J' else 0), _l(len(x)), _p(x), _p(y), _p(z), _p(w), _d(q),
                      _p(ux if what == 'J' else dummy), _p(uy if what == 'J' else dummy),
                      _p(uz if what == 'J' else dummy), _p(inv_gamma if what == 'J' else dummy),
                      _d(invdz), _d(zmin), _i(Nz), _d(invdr), _d(rmin), _i(Nr), _i(Nm), _i(int(cubic)),
                      _p(beta0), _p(beta_hi), _i(nthreads), _p(glob))
    out = np.zeros((ncomp, Nm, Nz, Nr), dtype=np.complex128)
    for k in range(ncomp):
        for m in range(Nm):
            lib().orc_sum_reduce(_p(glob), _i(nthreads), _i(ncomp), _i(Nm), _i(Nz), _i(Nr),
                                 _i(k), _i(m), _p(out[k, m]))
    return out


# ---------------------------------------------------------------------------
# spectral side (NumPy)
# ---------------------------------------------------------------------------
class OracleTransformer(object):
    """FFT along z + Hankel transforms of one azimuthal mode
    (spectral_transformer.py:21-223, hankel.py:182-243, fourier.py:104-168)."""

    def __init__(self, Nz, Nr, m, rmax, workers=None):
        self.Nz, self.Nr, self.m = Nz, Nr, m
        self.M0, self.iM0, self.nu = ht.hankel_matrices(m, m, Nr, rmax)
        self.Mp, self.iMp, _ = ht.hankel_matrices(m + 1, m, Nr, rmax)
        self.Mm, self.iMm, _ = ht.hankel_matrices(m - 1, m, Nr, rmax)
        self.workers = workers or nthreads_default()

    def fft(self, a):
        return sfft.fft(a, axis=0, workers=self.workers)

    def ifft(self, a):
        return sfft.ifft(a, axis=0, workers=self.workers)

    @staticmethod
    def dht(F, mat):
        # real GEMM on the [2Nz, Nr] split, as hankel.py:207-212 / 238-243
        Nz = F.shape[0]
        out = np.dot(np.concatenate((F.real, F.imag), axis=0), mat)
        return out[:Nz] + 1.j * out[Nz:]

    def interp2spect_scal(self, f):
        return self.dht(self.fft(f), self.M0)

    def interp2spect_vect(self, fr, ft):
        r, t = self.fft(fr), self.fft(ft)
        return self.dht(0.5 * (r - 1.j * t), self.Mp), self.dht(0.5 * (r + 1.j * t), self.Mm)

    def spect2interp_scal(self, s):
        return self.ifft(self.dht(s, self.iM0))

    def spect2interp_vect(self, sp, sm):
        p, mm = self.dht(sp, self.iMp), self.dht(sm, self.iMm)
        return self.ifft(p + mm), self.ifft(1.j * (p - mm))


class OracleSim(object):
    """Single-domain, z-periodic PIC cycle (main.py:346-586) on NumPy arrays."""

    FIELDS_I = ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz', 'Jr', 'Jt', 'Jz', 'rho')
    FIELDS_S = ('Ep', 'Em', 'Ez', 'Bp', 'Bm', 'Bz', 'Jp', 'Jm', 'Jz', 'rho_prev', 'rho_next')

    PML_I = ('Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml')
    PML_S = ('Ep_pml', 'Em_pml', 'Bp_pml', 'Bm_pml')

    def __init__(self, Nz, zmax, Nr, rmax, Nm, dt, zmin=0., n_order=-1,
                 v_comoving=None, use_galilean=True, particle_shape='linear',
                 filter_currents=True, nthreads=None, nr_damp=0,
                 current_correction='curl-free'):
        # radial PML (boundaries['r']='open'): the grid is enlarged by nr_damp cells
        # (main.py:293-294, boundary_communicator.py:371-397); particles are only gathered
        # inside the physical radius (particles.py:700-704)
        self.nr_damp = nr_damp
        self.use_pml = nr_damp > 0
        self.rmax_gather = rmax
        rmax = (Nr + nr_damp) * (rmax / Nr)
        Nr = Nr + nr_damp
        self.cdt_over_dr = c * dt / (rmax / Nr)
        self.current_correction = current_correction
        self.Nz, self.Nr, self.Nm, self.dt = Nz, Nr, Nm, dt
        self.zmin, self.zmax, self.rmax = zmin, zmax, rmax
        self.dz = (zmax - zmin) / Nz
        self.dr = rmax / Nr
        self.invdz, self.invdr = 1. / self.dz, 1. / self.dr
        self.cubic = (particle_shape == 'cubic')
        self.v_comoving = v_comoving
        self.use_galilean = use_galilean if v_comoving is not None else False
        self.filter_currents = filter_currents
        self.nthreads = nthreads or nthreads_default()
        self.time, self.iteration = 0., 0
        kz_true = 2 * np.pi * np.fft.fftfreq(Nz, self.dz)
        self.kz = ht.modified_kz(kz_true, n_order, self.dz)
        self.trans, self.interp, self.spect = [], [], []
        self.kr, self.coef, self.invvol, self.ruyten, self.filt, self.inv_k2 = [], [], [], [], [], []
        for m in range(Nm):
            tr = OracleTransformer(Nz, Nr, m, rmax, self.nthreads)
            self.trans.append(tr)
            kr = 2 * np.pi * tr.nu
            self.kr.append(kr)
            self.coef.append(ht.psatd_coefficients(self.kz, kr, dt, v_comoving, self.use_galilean))
            vol = ht.cell_volumes(m, Nr, rmax, self.dz)
            self.invvol.append(1. / vol)
            self.ruyten.append(ht.ruyten_coefs(vol, self.dr, self.dz))
            self.filt.append(ht.binomial_filters(kz_true, kr, self.dz, self.dr))
            self.inv_k2.append(ht.inverse_k2(self.kz, kr))
            names_i = self.FIELDS_I + (self.PML_I if self.use_pml else ())
            names_s = self.FIELDS_S + (self.PML_S if self.use_pml else ()) + \
                (('rho_next_z', 'rho_next_xy') if current_correction == 'cross-deposition' else ())
            self.interp.append({k: np.zeros((Nz, Nr), dtype=np.complex128) for k in names_i})
            self.spect.append({k: np.zeros((Nz, Nr), dtype=np.complex128) for k in names_s})
        # pml_damping.py:86-108
        self.pml_damp = np.exp(-4. * self.cdt_over_dr * (np.arange(nr_damp) * 1. / max(nr_damp, 1))**2)
        self.species = []

    # -- species: dict of SoA arrays + q, m
    def add_species(self, q, m, x, y, z, ux, uy, uz, inv_gamma, w):
        f64 = lambda a: np.array(a, dtype=np.float64, copy=True)
        n = len(x)
        sp = dict(q=q, m=m, x=f64(x), y=f64(y), z=f64(z), ux=f64(ux), uy=f64(uy), uz=f64(uz),
                  inv_gamma=f64(inv_gamma), w=f64(w),
                  Ex=np.zeros(n), Ey=np.zeros(n), Ez=np.zeros(n),
                  Bx=np.zeros(n), By=np.zeros(n), Bz=np.zeros(n))
        self.species.append(sp)
        return sp

    # -- particle operators
    def gather(self, sp):
        if sp['q'] == 0:
            return
        grids = [tuple(g[k] for k in ('Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz')) for g in self.interp]
        gather(sp['x'], sp['y'], sp['z'], self.rmax_gather, self.invdz, self.zmin, self.Nz,
               self.invdr, 0., self.Nr, grids, self.cubic,
               sp['Ex'], sp['Ey'], sp['Ez'], sp['Bx'], sp['By'], sp['Bz'])

    def push_p(self, sp):
        if sp['q'] == 0:
            return
        push_p(sp['ux'], sp['uy'], sp['uz'], sp['inv_gamma'], sp['Ex'], sp['Ey'], sp['Ez'],
               sp['Bx'], sp['By'], sp['Bz'], sp['q'], sp['m'], self.dt)

    def push_x(self, sp, dt, x_push=1., y_push=1., z_push=1.):
        push_x(sp['x'], sp['y'], sp['z'], sp['ux'], sp['uy'], sp['uz'], sp['inv_gamma'], dt,
               x_push, y_push, z_push)

    def deposit_interp(self, what):
        """erase + deposit all species + fold + divide by volume (main.py:628-657)."""
        ncomp = 3 if what == 'J' else 1
        tot = np.zeros((ncomp, self.Nm, self.Nz, self.Nr), dtype=np.complex128)
        r_hi = 1 if self.Nm > 1 else 0
        bsel = 1 if self.cubic else 0
        for sp in self.species:
            if sp['q'] == 0:
                continue
            tot += deposit(what, sp['x'], sp['y'], sp['z'], sp['w'], sp['q'], sp['ux'], sp['uy'], sp['uz'],
                           sp['inv_gamma'], self.invdz, self.zmin, self.Nz, self.invdr, 0., self.Nr, self.Nm,
                           self.cubic, self.ruyten[0][bsel], self.ruyten[r_hi][bsel], self.nthreads)
        names = ('Jr', 'Jt', 'Jz') if what == 'J' else ('rho',)
        for m in range(self.Nm):
            for k, name in enumerate(names):
                self.interp[m][name][:, :] = tot[k, m] * self.invvol[m][np.newaxis, :]

    def deposit(self, fieldtype):
        """main.py:588-670 (single domain: no exchange)."""
        if fieldtype.startswith('rho'):
            self.deposit_interp('rho')
            for m in range(self.Nm):
                s = self.trans[m].interp2spect_scal(self.interp[m]['rho'])
                if self.filter_currents:
                    s *= self.filt[m][0][:, None] * self.filt[m][1][None, :]
                self.spect[m][fieldtype][:, :] = s
        elif fieldtype == 'J':
            self.deposit_interp('J')
            for m in range(self.Nm):
                g, tr = self.interp[m], self.trans[m]
                jz = tr.interp2spect_scal(g['Jz'])
                jp, jm = tr.interp2spect_vect(g['Jr'], g['Jt'])
                f = self.filt[m][0][:, None] * self.filt[m][1][None, :] if self.filter_currents else 1.
                self.spect[m]['Jz'][:, :] = jz * f
                self.spect[m]['Jp'][:, :] = jp * f
                self.spect[m]['Jm'][:, :] = jm * f
        else:
            raise ValueError(fieldtype)

    # -- spectral operators
    def interp2spect(self, ft):
        for m in range(self.Nm):
            g, s, tr = self.interp[m], self.spect[m], self.trans[m]
            if ft in ('E', 'B'):
                s[ft + 'z'][:, :] = tr.interp2spect_scal(g[ft + 'z'])
                s[ft + 'p'][:, :], s[ft + 'm'][:, :] = tr.interp2spect_vect(g[ft + 'r'], g[ft + 't'])
            elif ft in ('E_pml', 'B_pml'):        # fields.py:341-352
                f = ft[0]
                s[f + 'p_pml'][:, :], s[f + 'm_pml'][:, :] = tr.interp2spect_vect(g[f + 'r_pml'], g[f + 't_pml'])
            else:
                raise ValueError(ft)

    def spect2interp(self, ft):
        for m in range(self.Nm):
            g, s, tr = self.interp[m], self.spect[m], self.trans[m]
            if ft in ('E', 'B', 'J'):
                g[ft + 'z'][:, :] = tr.spect2interp_scal(s[ft + 'z'])
                g[ft + 'r'][:, :], g[ft + 't'][:, :] = tr.spect2interp_vect(s[ft + 'p'], s[ft + 'm'])
            elif ft in ('rho_prev', 'rho_next'):
                g['rho'][:, :] = tr.spect2interp_scal(s[ft])
            elif ft in ('E_pml', 'B_pml'):        # fields.py:398-409
                f = ft[0]
                g[f + 'r_pml'][:, :], g[f + 't_pml'][:, :] = tr.spect2interp_vect(s[f + 'p_pml'], s[f + 'm_pml'])
            else:
                raise ValueError(ft)

    def damp_pml_EB(self):
        """pml_damping.py:46-83 (CPU branch)"""
        n, d = self.nr_damp, self.pml_damp[np.newaxis, :]
        for g in self.interp:
            g['Et'][:, -n:] -= g['Et_pml'][:, -n:]
            g['Bt'][:, -n:] -= g['Bt_pml'][:, -n:]
            g['Et_pml'][:, -n:] *= d
            g['Bt_pml'][:, -n:] *= d
            g['Et'][:, -n:] += g['Et_pml'][:, -n:]
            g['Bt'][:, -n:] += g['Bt_pml'][:, -n:]
            g['Bz'][:, -n:] *= d
            g['Ez'][:, -n:] *= d

    def correct_currents_cross(self):
        """numba_methods.py:88-116 (standard), :243-275 (comoving)."""
        inv_dt = 1. / self.dt
        for m in range(self.Nm):
            s, t = self.spect[m], self.coef[m]
            kz = np.broadcast_to(self.kz[:, None], (self.Nz, self.Nr))
            kr = np.broadcast_to(self.kr[m][None, :], (self.Nz, self.Nr))
            rn, rp, rz, rxy = s['rho_next'], s['rho_prev'], s['rho_next_z'], s['rho_next_xy']
            if self.v_comoving is None:
                Dz = 1.j * kz * s['Jz'] + 0.5 * inv_dt * (rn - rxy + rz - rp)
                Dxy = kr * (s['Jp'] - s['Jm']) + 0.5 * inv_dt * (rn - rz + rxy - rp)
            else:
                a, b = 0.5 * t['T_cc'] * t['j_corr_coef'], t['T_eb']
                Dz = 1.j * kz * s['Jz'] + a * (rn - b * rxy + rz - b * rp)
                Dxy = kr * (s['Jp'] - s['Jm']) + a * (rn + b * rxy - rz - b * rp)
            with np.errstate(divide='ignore', invalid='ignore'):
                dxy = np.where(kr != 0, 0.5 * Dxy / kr, 0.)
                dz = np.where(kz != 0, 1.j * Dz / kz, 0.)
            s['Jp'] -= dxy
            s['Jm'] += dxy
            s['Jz'] += dz

    def cross_deposit(self, move_positions):
        """main.py:672-717"""
        dt = self.dt
        if move_positions:
            for sp in self.species:
                self.push_x(sp, 0.5 * dt, 1., 1., -1.)
        if self.use_galilean:
            self.shift_galilean(-0.5 * dt)
        self.deposit('rho_next_xy')
        if move_positions:
            for sp in self.species:
                self.push_x(sp, dt, -1., -1., 1.)
        if self.use_galilean:
            self.shift_galilean(dt)
        self.deposit('rho_next_z')
        if move_positions:
            for sp in self.species:
                self.push_x(sp, 0.5 * dt, 1., 1., -1.)
        if self.use_galilean:
            self.shift_galilean(-0.5 * dt)

    def correct_currents(self):
        """numba_methods.py:64-86 (standard), :217-241 (comoving)."""
        if self.current_correction == 'cross-deposition':
            return self.correct_currents_cross()
        inv_dt = 1. / self.dt
        for m in range(self.Nm):
            s, t = self.spect[m], self.coef[m]
            kz, kr = self.kz[:, None], self.kr[m][None, :]
            if self.v_comoving is None:
                F = -self.inv_k2[m] * ((s['rho_next'] - s['rho_prev']) * inv_dt
                                       + 1.j * kz * s['Jz'] + kr * (s['Jp'] - s['Jm']))
            else:
                F = -self.inv_k2[m] * (t['T_cc'] * t['j_corr_coef'] * (s['rho_next'] - s['rho_prev'] * t['T_eb'])
                                       + 1.j * kz * s['Jz'] + kr * (s['Jp'] - s['Jm']))
            s['Jp'] += 0.5 * kr * F
            s['Jm'] += -0.5 * kr * F
            s['Jz'] += -1.j * kz * F

    def push_eb(self, use_true_rho=False):
        """numba_methods.py:119-186 (standard), :278-355 (comoving/Galilean); then push_rho."""
        dt = self.dt
        c2 = c**2
        for m in range(self.Nm):
            s, t = self.spect[m], self.coef[m]
            kz, kr = self.kz[:, None], self.kr[m][None, :]
            C, S_w, j_coef = t['C'], t['S_w'], t['j_coef']
            Ep, Em, Ez = s['Ep'].copy(), s['Em'].copy(), s['Ez'].copy()
            Bp, Bm, Bz = s['Bp'], s['Bm'], s['Bz']
            Jp, Jm, Jz = s['Jp'], s['Jm'], s['Jz']
            std = self.v_comoving is None
            if self.use_pml:
                # split components first, from the old Ez, Bz (numba_methods.py:189-214, 358-383)
                Tp = 1. if std else t['T_eb']
                for f in ('Ep_pml', 'Em_pml'):
                    s[f][:, :] = Tp * C * s[f] + c2 * Tp * S_w * (-1.j * 0.5 * kr * Bz)
                for f in ('Bp_pml', 'Bm_pml'):
                    s[f][:, :] = Tp * C * s[f] - Tp * S_w * (-1.j * 0.5 * kr * Ez)
            if use_true_rho:
                rho_diff = t['rho_next_coef'] * s['rho_next'] - t['rho_prev_coef'] * s['rho_prev']
            else:
                divE = kr * (Ep - Em) + 1.j * kz * Ez
                divJ = kr * (Jp - Jm) + 1.j * kz * Jz
                if std:
                    rho_diff = (t['rho_next_coef'] - t['rho_prev_coef']) * epsilon_0 * divE \
                        - t['rho_next_coef'] * dt * divJ
                else:
                    rho_diff = (t['T_eb'] * t['rho_next_coef'] - t['rho_prev_coef']) * epsilon_0 * divE \
                        + t['T_rho'] * t['rho_next_coef'] * divJ
            if std:
                Teb, Tcc, gal = 1., 1., 0.
            else:
                Teb, Tcc, gal = t['T_eb'], t['T_cc'], j_coef * 1.j * kz * self.v_comoving
            s['Ep'][:, :] = Teb * C * Ep + 0.5 * kr * rho_diff + gal * Jp \
                + c2 * Teb * S_w * (-1.j * 0.5 * kr * Bz + kz * Bp - mu_0 * Tcc * Jp)
            s['Em'][:, :] = Teb * C * Em - 0.5 * kr * rho_diff + gal * Jm \
                + c2 * Teb * S_w * (-1.j * 0.5 * kr * Bz - kz * Bm - mu_0 * Tcc * Jm)
            s['Ez'][:, :] = Teb * C * Ez - 1.j * kz * rho_diff + gal * Jz \
                + c2 * Teb * S_w * (1.j * kr * Bp + 1.j * kr * Bm - mu_0 * Tcc * Jz)
            nBp = Teb * C * Bp - Teb * S_w * (-1.j * 0.5 * kr * Ez + kz * Ep) \
                + j_coef * (-1.j * 0.5 * kr * Jz + kz * Jp)
            nBm = Teb * C * Bm - Teb * S_w * (-1.j * 0.5 * kr * Ez - kz * Em) \
                + j_coef * (-1.j * 0.5 * kr * Jz - kz * Jm)
            nBz = Teb * C * Bz - Teb * S_w * (1.j * kr * Ep + 1.j * kr * Em) \
                + j_coef * (1.j * kr * Jp + 1.j * kr * Jm)
            s['Bp'][:, :], s['Bm'][:, :], s['Bz'][:, :] = nBp, nBm, nBz
            s['rho_prev'][:, :] = s['rho_next']
            s['rho_next'][:, :] = 0.

    def shift_galilean(self, dt):
        self.zmin += self.v_comoving * dt
        self.zmax += self.v_comoving * dt

    # -- the PIC cycle
    def step(self, N=1, correct_currents=True, use_true_rho=False,
             move_positions=True, move_momenta=True):
        dt = self.dt
        self.interp2spect('E')
        self.interp2spect('B')
        if self.use_pml:
            self.interp2spect('E_pml')
            self.interp2spect('B_pml')
        for i_step in range(N):
            # single periodic domain: exchange_period == 1 (boundary_communicator.py:283-286)
            for sp in self.species:
                shift_periodic(sp['z'], self.zmin, self.zmax)
            self.deposit('rho_prev')
            if i_step == 0:
                self.deposit('J')
            for sp in self.species:
                self.gather(sp)
            if move_momenta:
                for sp in self.species:
                    self.push_p(sp)
            if move_positions:
                for sp in self.species:
                    self.push_x(sp, 0.5 * dt)
            if self.use_galilean:
                self.shift_galilean(0.5 * dt)
            self.deposit('J')
            if correct_currents and self.current_correction == 'cross-deposition':
                self.cross_deposit(move_positions)
            if move_positions:
                for sp in self.species:
                    self.push_x(sp, 0.5 * dt)
            if self.use_galilean:
                self.shift_galilean(0.5 * dt)
            self.deposit('rho_next')
            if correct_currents:
                self.correct_currents()
            self.push_eb(use_true_rho)
            # exchange_and_damp_EB (main.py:719-769) on one periodic domain: iFFT/FFT
            # round trip is an identity; only the final spect2interp remains.
            self.spect2interp('E')
            self.spect2interp('B')
            if self.use_pml:
                # full transforms both ways around the radial damping (main.py:732-761)
                self.spect2interp('E_pml')
                self.spect2interp('B_pml')
                self.damp_pml_EB()
                for ft in ('E', 'B', 'E_pml', 'B_pml'):
                    self.interp2spect(ft)
            self.time += dt
            self.iteration += 1
        self.spect2interp('J')
        self.spect2interp('rho_prev')

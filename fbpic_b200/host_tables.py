"""
Host-side (NumPy/SciPy, fp64) tables that feed the B200 hot path.

These tables are computed ONCE per simulation on the host and uploaded to HBM;
they are inputs of the CUDA kernels, not part of the per-step hot loop.
Every function cites the reference formula it restates (file:line relative to
the FBPIC source tree); none of this is executed per step.
"""
import numpy as np
from scipy.special import jn, jn_zeros, j1
from scipy.constants import c, mu_0, epsilon_0


# ---------------------------------------------------------------------------
# Discrete Hankel transform matrices
# ---------------------------------------------------------------------------
def bessel_alphas(m, Nr):
    """Spectral radial grid (Bessel zeros) of azimuthal mode `m`.

    For m != 0 the value 0 is prepended (kr=0 mode) and only Nr-1 zeros are
    used -- fbpic/fields/spectral_transform/hankel.py:74-81.
    """
    if m == 0:
        return jn_zeros(0, Nr)
    return np.concatenate(([0.], jn_zeros(m, Nr - 1)))


def hankel_matrices(p, m, Nr, rmax):
    """Return (M, invM, nu) of the order-`p` DHT used for azimuthal mode `m`.

    `transform(F) = F @ M`, `inverse(G) = G @ invM` on `[.., Nr]` rows
    (matrices are "transposed w.r.t. the paper", hankel.py:90-92, 207-243).
    invM[n, j] = J_p(alpha_n r_j / rmax) / (pi rmax^2 J_{p'}(alpha_n)^2),
    p' = p+1 if p == m else p (hankel.py:93-114); for m != 0 the first row is
    the kr=0 mode r^(m-1)/(pi rmax^(m+1)) when p == m-1, else 0 (:104-112).
    M = inv(invM), or pinv of rows 1.. with a zero first column (:117-122).
    """
    if m not in (p - 1, p, p + 1):
        raise ValueError('m must be either p-1, p or p+1')
    alphas = bessel_alphas(m, Nr)
    nu = 1. / (2 * np.pi * rmax) * alphas
    r = (rmax / Nr) * (np.arange(Nr) + 0.5)
    order_den = p + 1 if p == m else p
    den = np.pi * rmax**2 * jn(order_den, alphas)**2
    num = jn(p, 2 * np.pi * r[None, :] * nu[:, None])
    invM = np.empty((Nr, Nr))
    if m == 0:
        invM[:, :] = num / den[:, None]
    else:
        invM[1:, :] = num[1:, :] / den[1:, None]
        if p == m - 1:
            invM[0, :] = r**(m - 1) / (np.pi * rmax**(m + 1))
        else:
            invM[0, :] = 0.
    if m != 0 and p != m - 1:
        M = np.empty((Nr, Nr))
        M[:, 1:] = np.linalg.pinv(invM[1:, :])
        M[:, 0] = 0.
    else:
        M = np.linalg.inv(invM)
    return np.ascontiguousarray(M), np.ascontiguousarray(invM), nu


# ---------------------------------------------------------------------------
# Interpolation-grid tables: cell volumes and Ruyten coefficients
# ---------------------------------------------------------------------------
def cell_volumes(m, Nr, rmax, dz, use_modified_volume=True):
    """Cell volumes vol[ir] (fbpic/fields/interpolation_grid.py:88-97).

    Mode 0 uses the Hankel-corrected effective volume
    dz * sum_n M0[ir, n] * 2 / (alpha_n J1(alpha_n)); modes m>=1 use the
    geometric ring volume pi dz ((r+dr/2)^2 - (r-dr/2)^2).
    """
    dr = rmax / Nr
    if use_modified_volume and m == 0:
        alphas = jn_zeros(0, Nr)
        M0, _, _ = hankel_matrices(0, 0, Nr, rmax)
        wgt = alphas * j1(alphas)
        # row-by-row 1-D sums: keeps NumPy's pairwise summation order identical to the reference
        return dz * np.array([(M0[ir, :] * 2. / wgt).sum() for ir in range(Nr)])
    r = (0.5 + np.arange(Nr)) * dr
    return np.pi * dz * ((r + 0.5 * dr)**2 - (r - 0.5 * dr)**2)


def ruyten_coefs(vol, dr, dz, use_ruyten_shapes=True):
    """(linear, cubic) Ruyten coefficient arrays of length Nr+1, leading 0.

    interpolation_grid.py:107-138: beta_n = 6/(n+1) (cumsum(v) - (n+1)^2/2 - k)
    with k = 1/24 (linear), 1/8 (cubic, first entry special-cased).
    """
    Nr = len(vol)
    if not use_ruyten_shapes:
        z = np.zeros(Nr + 1)
        return z, z.copy()
    n1 = np.arange(Nr) + 1.
    vnorm = vol / (2 * np.pi * dr**2 * dz)
    cs = np.cumsum(vnorm)
    lin = 6. / n1 * (cs - 0.5 * n1**2 - 1. / 24)
    cub = 6. / n1 * (cs - 0.5 * n1**2 - 1. / 8)
    cub[0] = 6. * (vnorm[0] - 0.5 - 239. / (15 * 2**7))
    return np.concatenate(([0.], lin)), np.concatenate(([0.], cub))


# ---------------------------------------------------------------------------
# Spectral-grid tables
# ---------------------------------------------------------------------------
def modified_kz(kz_true, n_order, dz):
    """Finite-order modified wavenumber (fbpic/fields/utility_methods.py:11-66).

    [k] = sum_{n=1..m} a_n sin(n k dz)/(n dz),  a_0 = -2,
    a_n = -(m+1-n)/(m+n) a_{n-1},  m = n_order/2;  n_order = -1 -> k.
    """
    if n_order == -1:
        return kz_true
    if n_order % 2 == 1 or n_order <= 0:
        raise ValueError('Invalid n_order: %d' % n_order)
    half = n_order // 2
    a = np.zeros(half + 1)
    a[0] = -2.
    for n in range(1, half + 1):
        a[n] = -(half + 1 - n) * 1. / (half + n) * a[n - 1]
    n_arr = np.arange(1, half + 1)
    s = np.sin(kz_true[:, None] * n_arr[None, :] * dz) / (n_arr[None, :] * dz)
    return np.tensordot(s, a[1:], axes=(-1, -1))


def stencil_reach(Nz, dz, cdt, n_order, v_comoving, use_galilean):
    """Number of cells after which the PSATD stencil drops below 1e-16.

    utility_methods.py:69-185 (evaluated at kperp = 0.5); n_guard = reach+1
    (fbpic/boundaries/boundary_communicator.py:243-250).
    """
    kz = modified_kz(2 * np.pi * np.fft.fftfreq(Nz, d=dz), n_order, dz)
    kperp = 0.5
    k = np.sqrt(kz**2 + kperp**2)
    if use_galilean is True:
        theta2 = np.exp(1.j * np.abs(v_comoving) * kz * cdt / c / 2)**2
    else:
        theta2 = np.ones_like(kz)
    st_c = np.fft.ifft(theta2 * np.cos(k * cdt))
    sk = theta2 * np.sin(k * cdt) / np.where(k == 0, 1., k)
    st_z = np.fft.ifft(np.where(k == 0, kz, sk * kz))
    st_p = np.fft.ifft(np.where(k == 0, kperp, sk * kperp))
    alpha = np.sqrt(np.abs(st_c)**2 + np.abs(st_z)**2 + np.abs(st_p)**2)
    return int(np.where(alpha[:alpha.shape[0] // 2] < 1.e-16)[0][0])


def binomial_filters(kz_true, kr, dz, dr, n_passes=None, compensator=None):
    """(filter_z[Nz], filter_r[Nr]) of the binomial smoother.

    fbpic/fields/smoothing.py:57-94: (1 - sin^2(k d/2))^n [ * (1 + n sin^2) ].
    """
    n_passes = n_passes or {'z': 1, 'r': 1}
    compensator = compensator or {'z': False, 'r': False}
    out = []
    for k, d, ax in ((kz_true, dz, 'z'), (kr, dr, 'r')):
        s2 = np.sin(0.5 * k * d)**2
        n = n_passes[ax]
        f = (1. - s2)**n
        if compensator[ax]:
            f = f * (1. + n * s2)
        out.append(f)
    return out[0], out[1]


def inverse_k2(kz, kr):
    """1/(kz^2+kr^2) on the [Nz,Nr] mesh, 0 at k=0 (spectral_grid.py:116-120)."""
    k2 = kz[:, None]**2 + kr[None, :]**2
    zero = (k2 == 0)
    inv = 1. / np.where(zero, 1., k2)
    inv[zero] = 0.
    return inv


def psatd_coefficients(kz, kr, dt, V=None, use_galilean=False):
    """PSATD coefficient tables on the [Nz,Nr] mesh (fbpic/fields/psatd_coefs.py:66-163).

    Returns a dict with real tables C, S_w and (real if V is None, else complex)
    j_coef, rho_prev_coef, rho_next_coef; for the comoving/Galilean scheme also
    T_eb, T_cc, T_rho, j_corr_coef.  `kz`, `kr` are the 1-D (modified) axes.
    """
    KZ, KR = np.meshgrid(kz, kr, indexing='ij')
    w = c * np.sqrt(KZ**2 + KR**2)
    w0 = (w == 0)
    inv_w = 1. / np.where(w0, 1., w)
    inv_dt = 1. / dt
    t = {}
    t['C'] = np.cos(w * dt)
    S_w = np.sin(w * dt) * inv_w
    S_w[w0] = dt
    t['S_w'] = S_w
    comoving = (V is not None) and (V != 0.)
    if V is not None:
        T2 = np.exp(1.j * KZ * V * dt)
        if use_galilean:
            t['T_eb'] = T2
            t['T_cc'] = np.ones_like(T2)
        else:
            t['T_cc'] = np.exp(0.5j * KZ * V * dt)
            t['T_eb'] = np.ones_like(T2)
        if comoving:
            ikzV = 1.j * KZ * V
            ikzV[KZ == 0] = 1.
            t['T_rho'] = np.where(KZ == 0., -dt, (1. - T2) / (t['T_cc'] * ikzV))
            d = w**2 - KZ**2 * V**2
            inv_d = 1. / np.where(d == 0, 1., d)
            inv_1_T2 = 1. / np.where(T2 == 1, 1., 1 - T2)
            C = t['C']
            xi_1 = 1. / t['T_cc'] * inv_d * (1. - T2 * C + 1.j * KZ * V * T2 * S_w)
            xi_2 = np.where(
                KZ != 0,
                inv_d * (1. + 1.j * KZ * V * T2 * S_w * inv_1_T2
                         + KZ**2 * V**2 * inv_w**2 * T2 * inv_1_T2 * (1 - C)),
                inv_w**2 * (1. - S_w * inv_dt))
            xi_3 = np.where(
                KZ != 0,
                t['T_eb'] * inv_d * (C + 1.j * KZ * V * T2 * S_w * inv_1_T2
                                     + KZ**2 * V**2 * inv_w**2 * inv_1_T2 * (1 - C)),
                inv_w**2 * (C - S_w * inv_dt))
            t['j_corr_coef'] = np.where(KZ != 0, (-1.j * KZ * V) * inv_1_T2, inv_dt)
        else:
            t['T_rho'] = -dt * np.ones_like(KZ)
            t['j_corr_coef'] = inv_dt * np.ones_like(KZ)
    if comoving:
        j_coef = mu_0 * c**2 * xi_1
        rho_prev = c**2 / epsilon_0 * xi_3
        rho_next = c**2 / epsilon_0 * xi_2
    else:
        j_coef = mu_0 * c**2 * (1. - t['C']) * inv_w**2
        rho_prev = c**2 / epsilon_0 * (t['C'] - inv_dt * S_w) * inv_w**2
        rho_next = c**2 / epsilon_0 * (1 - inv_dt * S_w) * inv_w**2
    j_coef[w0] = mu_0 * c**2 * (0.5 * dt**2)
    rho_prev[w0] = c**2 / epsilon_0 * (-1. / 3 * dt**2)
    rho_next[w0] = c**2 / epsilon_0 * (1. / 6 * dt**2)
    t['j_coef'], t['rho_prev_coef'], t['rho_next_coef'] = j_coef, rho_prev, rho_next
    return t


# ---------------------------------------------------------------------------
# Boundary tables
# ---------------------------------------------------------------------------
def damp_array(n_guard, nz_damp, n_inject):
    """sin^2 damping profile of the open-z boundary, length ng+nd+ni.

    boundary_communicator.py:909-945: 0 below ng+ni, sin^2 ramp over nz_damp/2
    cells, then 1.
    """
    i = np.arange(n_guard + nz_damp + n_inject)
    edge = n_guard + n_inject
    d = np.where(i < edge + nz_damp / 2.,
                 np.sin((i - edge) * np.pi / (2 * nz_damp / 2.))**2, 1.)
    return np.where(i < edge, 0., d)


def pml_damp_array(n_pml, cdt_over_dr):
    """Radial PML damping profile (fbpic/boundaries/pml_damping.py:86-108)."""
    x_pml = np.arange(n_pml) * 1. / n_pml
    return np.exp(-4. * cdt_over_dr * x_pml**2)

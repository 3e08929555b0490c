/*
 * oracle/pic_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (OpenMP) restatement of the particle side of FBPIC's CPU (numba) hot
 * loop.  It is the parity checker for the CUDA kernels of fbpic_b200 and the
 * timed "cpu_baseline"/"--impl reference" leg of bench.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py's baseline legs may load it.
 * Parity pinned: checked against outputs of the unmodified reference generated
 * by oracle/gen_golden.py (npz files under tests/golden/, tests/test_oracle_golden.py).
 *
 * Each function cites the reference file:line whose arithmetic it follows
 * (paths relative to the FBPIC source tree).  Grids are complex128 [Nz][Nr]
 * row-major, passed as interleaved double pairs.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define C_LIGHT 299792458.0

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* parallel memset of the per-thread deposition copies (numba_erase_threading_buffer,
 * fbpic/fields/numba_methods.py:390-407) */
void orc_zero(double *a, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) a[i] = 0.;
}

/* ---- cell key: fbpic/particles/utilities/cuda_sorting.py:55-88 ---- */
void orc_cell_index(int64_t n, const double *x, const double *y, const double *z,
                    double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
                    int32_t *cell_idx) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double rj = sqrt(x[i] * x[i] + y[i] * y[i]);
        double r_cell = invdr * (rj - rmin) - 0.5;
        double z_cell = invdz * (z[i] - zmin) - 0.5;
        int ir_upper = (int)ceil(r_cell);
        int iz_upper = (int)ceil(z_cell);
        if (ir_upper > Nr) ir_upper = Nr;
        if (iz_upper < 0) iz_upper += Nz;
        else if (iz_upper > Nz - 1) iz_upper -= Nz;
        cell_idx[i] = ir_upper + iz_upper * (Nr + 1);
    }
}

/* ---- push: fbpic/particles/push/inline_functions.py:11-48, numba_methods.py:17-52 ---- */
void orc_push_p(int64_t n, double *ux, double *uy, double *uz, double *inv_gamma,
                const double *Ex, const double *Ey, const double *Ez,
                const double *Bx, const double *By, const double *Bz,
                double q, double m, double dt) {
    const double econst = q * dt / (m * C_LIGHT);
    const double bconst = 0.5 * q * dt / m;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double taux = bconst * Bx[i], tauy = bconst * By[i], tauz = bconst * Bz[i];
        double tau2 = taux * taux + tauy * tauy + tauz * tauz;
        double ig = inv_gamma[i];
        double uxp = ux[i] + econst * Ex[i] + ig * (uy[i] * tauz - uz[i] * tauy);
        double uyp = uy[i] + econst * Ey[i] + ig * (uz[i] * taux - ux[i] * tauz);
        double uzp = uz[i] + econst * Ez[i] + ig * (ux[i] * tauy - uy[i] * taux);
        double sigma = 1 + uxp * uxp + uyp * uyp + uzp * uzp - tau2;
        double utau = uxp * taux + uyp * tauy + uzp * tauz;
        double igf = sqrt(2. / (sigma + sqrt(sigma * sigma + 4 * (tau2 + utau * utau))));
        double tx = igf * taux, ty = igf * tauy, tz = igf * tauz, ut = igf * utau;
        double s = 1. / (1 + tau2 * igf * igf);
        ux[i] = s * (uxp + tx * ut + uyp * tz - uzp * ty);
        uy[i] = s * (uyp + ty * ut + uzp * tx - uxp * tz);
        uz[i] = s * (uzp + tz * ut + uxp * ty - uyp * tx);
        inv_gamma[i] = igf;
    }
}

void orc_push_x(int64_t n, double *x, double *y, double *z,
                const double *ux, const double *uy, const double *uz, const double *inv_gamma,
                double dt, double x_push, double y_push, double z_push) {
    const double chdt = C_LIGHT * dt;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        x[i] += chdt * inv_gamma[i] * x_push * ux[i];
        y[i] += chdt * inv_gamma[i] * y_push * uy[i];
        z[i] += chdt * inv_gamma[i] * z_push * uz[i];
    }
}

/* periodic wrap: fbpic/boundaries/particle_buffer_handling.py:537-560 */
void orc_shift_periodic(int64_t n, double *z, double zmin, double zmax) {
    double l_box = zmax - zmin;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        while (z[i] >= zmax) z[i] -= l_box;
        while (z[i] < zmin) z[i] += l_box;
    }
}

/* ---- gather: fbpic/particles/gathering/threading_methods.py:26-205 (linear), :208-367
 *      (cubic); inline_functions.py:9-91, 93-187.  `grids` holds 6*Nm pointers ordered
 *      [m][Er,Et,Ez,Br,Bt,Bz].  All modes are accumulated before the cylindrical->
 *      Cartesian rotation, as in the 2-mode kernels. ---- */
static inline void gather_mode_linear(int m, const double *Fr_g, const double *Ft_g, const double *Fz_g,
                                      int Nr, int iz_l, int iz_u, int ir_l, int ir_u,
                                      double S_ll, double S_lu, double S_lg, double S_ul, double S_uu, double S_ug,
                                      double e_re, double e_im, double *Fr, double *Ft, double *Fz) {
    const double *g[3] = {Fr_g, Ft_g, Fz_g};
    double acc_re[3], acc_im[3];
    for (int k = 0; k < 3; ++k) {
        const double *a = g[k];
        size_t ll = 2 * ((size_t)iz_l * Nr + ir_l), lu = 2 * ((size_t)iz_l * Nr + ir_u);
        size_t ul = 2 * ((size_t)iz_u * Nr + ir_l), uu = 2 * ((size_t)iz_u * Nr + ir_u);
        double re = 0., im = 0.;
        re += S_ll * a[ll];     im += S_ll * a[ll + 1];
        re += S_lu * a[lu];     im += S_lu * a[lu + 1];
        re += S_ul * a[ul];     im += S_ul * a[ul + 1];
        re += S_uu * a[uu];     im += S_uu * a[uu + 1];
        if (ir_l == 0 && ir_u == 0) {
            double flip = (m % 2 == 0) ? 1. : -1.;
            double sgn = (k == 2) ? flip : -flip;
            size_t l0 = 2 * ((size_t)iz_l * Nr), u0 = 2 * ((size_t)iz_u * Nr);
            re += sgn * S_lg * a[l0];   im += sgn * S_lg * a[l0 + 1];
            re += sgn * S_ug * a[u0];   im += sgn * S_ug * a[u0 + 1];
        }
        acc_re[k] = re; acc_im[k] = im;
    }
    double factor = (m == 0) ? 1. : 2.;
    *Fr += factor * (acc_re[0] * e_re - acc_im[0] * e_im);
    *Ft += factor * (acc_re[1] * e_re - acc_im[1] * e_im);
    *Fz += factor * (acc_re[2] * e_re - acc_im[2] * e_im);
}

static inline void gather_mode_cubic(int m, const double *Fr_g, const double *Ft_g, const double *Fz_g,
                                     int Nr, int Nz, int ir_lowest, int iz_lowest,
                                     const double *Sr, const double *Sz,
                                     double e_re, double e_im, double *Fr, double *Ft, double *Fz) {
    double r_re = 0, r_im = 0, t_re = 0, t_im = 0, z_re = 0, z_im = 0;
    double flip = (m % 2 == 0) ? 1. : -1.;
    for (int index_r = 0; index_r < 4; ++index_r) {
        int ir = ir_lowest + index_r;
        double Sr_long = Sr[index_r], Sr_perp = Sr[index_r];
        if (ir < 0) { Sr_long *= flip; Sr_perp *= -flip; }
        if (ir < 0) ir = abs(ir) - 1;
        else if (ir > Nr - 1) ir = Nr - 1;
        for (int index_z = 0; index_z < 4; ++index_z) {
            int iz = iz_lowest + index_z;
            double sz = Sz[index_z];
            if (iz < 0) iz += Nz;
            else if (iz > Nz - 1) iz -= Nz;
            size_t o = 2 * ((size_t)iz * Nr + ir);
            r_re += sz * Sr_perp * Fr_g[o]; r_im += sz * Sr_perp * Fr_g[o + 1];
            t_re += sz * Sr_perp * Ft_g[o]; t_im += sz * Sr_perp * Ft_g[o + 1];
            z_re += sz * Sr_long * Fz_g[o]; z_im += sz * Sr_long * Fz_g[o + 1];
        }
    }
    double factor = (m == 0) ? 1. : 2.;
    *Fr += factor * (r_re * e_re - r_im * e_im);
    *Ft += factor * (t_re * e_re - t_im * e_im);
    *Fz += factor * (z_re * e_re - z_im * e_im);
}

void orc_gather(int64_t n, const double *x, const double *y, const double *z,
                double rmax_gather, double invdz, double zmin, int Nz,
                double invdr, double rmin, int Nr, int Nm, const double *const *grids, int cubic,
                double *Ex, double *Ey, double *Ez, double *Bx, double *By, double *Bz) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double xj = x[i], yj = y[i], zj = z[i];
        double rj = sqrt(xj * xj + yj * yj);
        double cs, sn;
        if (rj != 0.) { double invr = 1. / rj; cs = xj * invr; sn = yj * invr; }
        else { cs = 1.; sn = 0.; }
        double r_cell = invdr * (rj - rmin) - 0.5;
        double z_cell = invdz * (zj - zmin) - 0.5;
        double F[2][3] = {{0, 0, 0}, {0, 0, 0}};
        if (rj < rmax_gather) {
            if (!cubic) {
                int ir_l = (int)floor(r_cell), ir_u = ir_l + 1;
                int iz_l = (int)floor(z_cell), iz_u = iz_l + 1;
                double Sr_l = ir_u - r_cell, Sr_u = r_cell - ir_l;
                double Sz_l = iz_u - z_cell, Sz_u = z_cell - iz_l;
                double Sr_g = 0.;
                if (ir_l < 0) { Sr_g = Sr_l; Sr_l = 0.; ir_l = 0; }
                if (ir_l > Nr - 1) ir_l = Nr - 1;
                if (ir_u > Nr - 1) ir_u = Nr - 1;
                if (iz_l < 0) iz_l += Nz;
                if (iz_u < 0) iz_u += Nz;
                if (iz_l > Nz - 1) iz_l -= Nz;
                if (iz_u > Nz - 1) iz_u -= Nz;
                double S_ll = Sz_l * Sr_l, S_lu = Sz_l * Sr_u, S_ul = Sz_u * Sr_l, S_uu = Sz_u * Sr_u;
                double S_lg = Sz_l * Sr_g, S_ug = Sz_u * Sr_g;
                double e_re = 1., e_im = 0.;
                for (int m = 0; m < Nm; ++m) {
                    for (int f = 0; f < 2; ++f)
                        gather_mode_linear(m, grids[6 * m + 3 * f], grids[6 * m + 3 * f + 1], grids[6 * m + 3 * f + 2],
                                           Nr, iz_l, iz_u, ir_l, ir_u, S_ll, S_lu, S_lg, S_ul, S_uu, S_ug,
                                           e_re, e_im, &F[f][0], &F[f][1], &F[f][2]);
                    /* exptheta_{m+1} = (cos - i sin) * exptheta_m */
                    double nr = e_re * cs + e_im * sn, ni = e_im * cs - e_re * sn;
                    e_re = nr; e_im = ni;
                }
            } else {
                double Sr[4], Sz[4];
                int ir_lowest = (int)floor(r_cell) - 1;
                double rl = r_cell - ir_lowest;
                Sr[0] = -1. / 6. * ((rl - 2.) * (rl - 2.) * (rl - 2.));
                Sr[1] = 1. / 6. * (3. * ((rl - 1.) * (rl - 1.) * (rl - 1.)) - 6. * ((rl - 1.) * (rl - 1.)) + 4.);
                Sr[2] = 1. / 6. * (3. * ((2. - rl) * (2. - rl) * (2. - rl)) - 6. * ((2. - rl) * (2. - rl)) + 4.);
                Sr[3] = -1. / 6. * ((1. - rl) * (1. - rl) * (1. - rl));
                int iz_lowest = (int)floor(z_cell) - 1;
                double zl = z_cell - iz_lowest;
                Sz[0] = -1. / 6. * ((zl - 2.) * (zl - 2.) * (zl - 2.));
                Sz[1] = 1. / 6. * (3. * ((zl - 1.) * (zl - 1.) * (zl - 1.)) - 6. * ((zl - 1.) * (zl - 1.)) + 4.);
                Sz[2] = 1. / 6. * (3. * ((2. - zl) * (2. - zl) * (2. - zl)) - 6. * ((2. - zl) * (2. - zl)) + 4.);
                Sz[3] = -1. / 6. * ((1. - zl) * (1. - zl) * (1. - zl));
                double e_re = 1., e_im = 0.;
                for (int m = 0; m < Nm; ++m) {
                    for (int f = 0; f < 2; ++f)
                        gather_mode_cubic(m, grids[6 * m + 3 * f], grids[6 * m + 3 * f + 1], grids[6 * m + 3 * f + 2],
                                          Nr, Nz, ir_lowest, iz_lowest, Sr, Sz, e_re, e_im,
                                          &F[f][0], &F[f][1], &F[f][2]);
                    double nr = e_re * cs + e_im * sn, ni = e_im * cs - e_re * sn;
                    e_re = nr; e_im = ni;
                }
            }
        }
        Ex[i] = cs * F[0][0] - sn * F[0][1];
        Ey[i] = sn * F[0][0] + cs * F[0][1];
        Ez[i] = F[0][2];
        Bx[i] = cs * F[1][0] - sn * F[1][1];
        By[i] = sn * F[1][0] + cs * F[1][1];
        Bz[i] = F[1][2];
    }
}

/* ---- deposition shapes: fbpic/particles/deposition/particle_shapes.py:17-80 ---- */
static inline double Sz_linear(double cp, int index) {
    double s = ceil(cp) - cp;
    if (index == 1) s = 1. - s;
    return s;
}
static inline double Sr_linear(double cp, int index, double flip, double beta_n) {
    int ir = (int)ceil(cp) - 1;
    double u = cp - ir;
    double s = (1. - u) + beta_n * (1. - u) * u;
    if (index == 1) s = 1. - s;
    if (index + ir < 0) s *= flip;
    return s;
}
static inline double cub(double a) { return a * a * a; }
static inline double Sz_cubic(double cp, int index) {
    int iz = (int)ceil(cp) - 2;
    double u = cp - iz - 1;
    double s = 0.;
    if (index == 0) s = (1. / 6.) * cub(1. - u);
    else if (index == 1) s = (1. / 6.) * (3. * cub(u) - 6. * (u * u) + 4.);
    else if (index == 2) s = (1. / 6.) * (3. * cub(1. - u) - 6. * ((1. - u) * (1. - u)) + 4.);
    else if (index == 3) s = (1. / 6.) * cub(u);
    return s;
}
static inline double Sr_cubic(double cp, int index, double flip, double beta_n) {
    int ir = (int)ceil(cp) - 2;
    double u = cp - ir - 1;
    double s = 0.;
    if (index == 0) s = (1. / 6.) * cub(1. - u);
    else if (index == 1) { s = (1. / 6.) * (3. * cub(u) - 6. * (u * u) + 4.); s += beta_n * (1. - u) * u; }
    else if (index == 2) { s = (1. / 6.) * (3. * cub(1. - u) - 6. * ((1. - u) * (1. - u)) + 4.); s -= beta_n * (1. - u) * u; }
    else if (index == 3) s = (1. / 6.) * cub(u);
    if (index + ir < 0) s *= flip;
    return s;
}

/* ---- deposition into per-thread guarded copies, fbpic/particles/deposition/
 *      threading_methods.py:28-148 (rho lin), :155-305 (J lin), :313-456 (rho cub), :459-650 (J cub).
 *      `glob` is [nthreads][ncomp][Nm][Nz+4][Nr+4] complex; comp order rho | Jr,Jt,Jz.
 *      what: 0 = rho, 1 = J. ---- */
void orc_deposit(int what, int64_t n, const double *x, const double *y, const double *z, const double *w, double q,
                 const double *ux, const double *uy, const double *uz, const double *inv_gamma,
                 double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm, int cubic,
                 const double *beta0, const double *beta_hi, int nthreads, double *glob) {
    const int ncomp = what ? 3 : 1;
    const size_t NZg = (size_t)Nz + 4, NRg = (size_t)Nr + 4;
    const size_t plane = NZg * NRg * 2;           /* doubles per (comp, mode) */
    const size_t per_thread = plane * Nm * ncomp;
    const int npts = cubic ? 4 : 2;
#pragma omp parallel for schedule(static, 1) num_threads(nthreads)
    for (int it = 0; it < nthreads; ++it) {
        /* chunking: fbpic/utils/threading.py get_chunk_indices (n/nthreads each, last takes the rest) */
        int64_t chunk = n / nthreads;
        int64_t p0 = it * chunk, p1 = (it == nthreads - 1) ? n : (it + 1) * chunk;
        double *G = glob + per_thread * it;
        double sc_re[3][16], sc_im[3][16];
        for (int64_t i = p0; i < p1; ++i) {
            double xj = x[i], yj = y[i], zj = z[i];
            double wj = q * w[i];
            double rj = sqrt(xj * xj + yj * yj);
            double cs, sn;
            if (rj != 0.) { double invr = 1. / rj; cs = xj * invr; sn = yj * invr; }
            else { cs = 1.; sn = 0.; }
            if (what == 0) {
                sc_re[0][0] = wj; sc_im[0][0] = 0.;
            } else {
                double ig = inv_gamma[i];
                sc_re[0][0] = wj * C_LIGHT * ig * (cs * ux[i] + sn * uy[i]); sc_im[0][0] = 0.;
                sc_re[1][0] = wj * C_LIGHT * ig * (cs * uy[i] - sn * ux[i]); sc_im[1][0] = 0.;
                sc_re[2][0] = wj * C_LIGHT * ig * uz[i];                      sc_im[2][0] = 0.;
            }
            for (int m = 1; m < Nm; ++m)
                for (int k = 0; k < ncomp; ++k) {
                    sc_re[k][m] = cs * sc_re[k][m - 1] - sn * sc_im[k][m - 1];
                    sc_im[k][m] = cs * sc_im[k][m - 1] + sn * sc_re[k][m - 1];
                }
            double r_cell = invdr * (rj - rmin) - 0.5;
            double z_cell = invdz * (zj - zmin) - 0.5;
            int ir_cell, iz_cell;
            if (!cubic) {
                ir_cell = (int)ceil(r_cell) + 1; if (ir_cell > Nr + 2) ir_cell = Nr + 2;
                iz_cell = (int)ceil(z_cell) + 1;
            } else {
                ir_cell = (int)ceil(r_cell); if (ir_cell > Nr) ir_cell = Nr;
                iz_cell = (int)ceil(z_cell);
            }
            int ir = (int)ceil(r_cell); if (ir > Nr) ir = Nr;
            for (int m = 0; m < Nm; ++m) {
                double bn = (m == 0) ? beta0[ir] : beta_hi[ir];
                double flip = (m % 2 == 0) ? 1. : -1.;
                for (int k = 0; k < ncomp; ++k) {
                    /* rho, Jz: flip=(-1)^m ; Jr, Jt: flip=-(-1)^m  (threading_methods.py:289-302) */
                    double fl = (what == 1 && k < 2) ? -flip : flip;
                    double *P = G + plane * ((size_t)k * Nm + m);
                    for (int a = 0; a < npts; ++a) {
                        double sz = cubic ? Sz_cubic(z_cell, a) : Sz_linear(z_cell, a);
                        for (int b = 0; b < npts; ++b) {
                            double sr = cubic ? Sr_cubic(r_cell, b, fl, bn) : Sr_linear(r_cell, b, fl, bn);
                            double s = sz * sr;
                            size_t o = 2 * ((size_t)(iz_cell + a) * NRg + (ir_cell + b));
                            P[o] += s * sc_re[k][m];
                            P[o + 1] += s * sc_im[k][m];
                        }
                    }
                }
            }
        }
    }
}

/* ---- fold of the guarded per-thread copies: fbpic/fields/numba_methods.py:410-461.
 *      Adds into `out` (complex [Nz][Nr]) the (comp,mode) plane `icm` of every thread. ---- */
static void reduce_slice(double *out, int Nr, int iz, const double *glob, size_t per_thread, size_t plane_off,
                         int nthreads, int iz_global) {
    const size_t NRg = (size_t)Nr + 4;
    double *o = out + 2 * (size_t)iz * Nr;
    for (int it = 0; it < nthreads; ++it) {
        const double *g = glob + per_thread * it + plane_off + 2 * (size_t)iz_global * NRg;
        o[2 * 1] += g[0];          o[2 * 1 + 1] += g[1];
        o[0] += g[2];              o[1] += g[3];
        for (int ir = 0; ir < Nr; ++ir) { o[2 * ir] += g[2 * (ir + 2)]; o[2 * ir + 1] += g[2 * (ir + 2) + 1]; }
        o[2 * (Nr - 1)] += g[2 * (Nr + 2)]; o[2 * (Nr - 1) + 1] += g[2 * (Nr + 2) + 1];
        o[2 * (Nr - 1)] += g[2 * (Nr + 3)]; o[2 * (Nr - 1) + 1] += g[2 * (Nr + 3) + 1];
    }
}
void orc_sum_reduce(const double *glob, int nthreads, int ncomp, int Nm, int Nz, int Nr, int icomp, int m, double *out) {
    const size_t plane = ((size_t)Nz + 4) * ((size_t)Nr + 4) * 2;
    const size_t per_thread = plane * Nm * ncomp;
    const size_t off = plane * ((size_t)icomp * Nm + m);
#pragma omp parallel for schedule(static)
    for (int iz = 0; iz < Nz; ++iz) reduce_slice(out, Nr, iz, glob, per_thread, off, nthreads, iz + 2);
    reduce_slice(out, Nr, Nz - 2, glob, per_thread, off, nthreads, 0);
    reduce_slice(out, Nr, Nz - 1, glob, per_thread, off, nthreads, 1);
    reduce_slice(out, Nr, 0, glob, per_thread, off, nthreads, Nz + 2);
    reduce_slice(out, Nr, 1, glob, per_thread, off, nthreads, Nz + 3);
}

// b2_particles.cu -- cell keys, stable cell sort + prefix sum, SoA permutation, field gather,
// Vay push, position push and the fused gather+push kernel.
#include "b2_common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <climits>

static inline unsigned grid1d(int64_t n, int tpb) { return (unsigned)((n + tpb - 1) / tpb); }

// =====================================================================================
// cell key (cuda_sorting.py:55-88)
// =====================================================================================
__global__ void k_cell_index(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                             const double *__restrict__ z, double invdz, double zmin, int Nz,
                             double invdr, double rmin, int Nr, int32_t *__restrict__ cell_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    B2Cyl c = b2_cyl(x[i], y[i], z[i], invdz, zmin, invdr, rmin);
    cell_idx[i] = b2_cell_of(c, Nz, Nr);
}

// =====================================================================================
// stable sort by cell + inclusive per-cell prefix sum (cuda_sorting.py:91-190)
// =====================================================================================
__global__ void k_iota(int64_t n, int32_t *v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}

// prefix_sum[c] = #particles with key <= c = upper_bound(keys_sorted, c): one thread per cell,
// cost independent of how the particles are distributed over the cells.
__global__ void k_prefix_sum(int64_t n, const int32_t *__restrict__ keys_sorted,
                             int32_t *__restrict__ prefix_sum, int ncells) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int64_t lo = 0, hi = n;                       // first position whose key is > c
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys_sorted + mid) <= c) lo = mid + 1;
        else hi = mid;
    }
    prefix_sum[c] = (int32_t)lo;
}

__global__ void k_finish_sort(int64_t n, const int32_t *__restrict__ keys_sorted,
                              const int32_t *__restrict__ idx32, int32_t *__restrict__ keys_out,
                              int64_t *__restrict__ sorted_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys_out[i] = keys_sorted[i];
    sorted_idx[i] = (int64_t)idx32[i];
}

struct B2Perm {
    const double *src[B2_MAX_ARRAYS];
    double *dst[B2_MAX_ARRAYS];
};
template <int NA>
__global__ void k_permute(int64_t n, const int64_t *__restrict__ sorted_idx, const int32_t *__restrict__ idx32,
                          B2Perm a, int n_arrays) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t j = sorted_idx ? sorted_idx[i] : (int64_t)idx32[i];
    if (NA > 0) {
#pragma unroll
        for (int k = 0; k < NA; ++k) a.dst[k][i] = __ldg(a.src[k] + j);
    } else {
        for (int k = 0; k < n_arrays; ++k) a.dst[k][i] = __ldg(a.src[k] + j);
    }
}

// =====================================================================================
// gather (gathering/cuda_methods.py:26-354, inline_functions.py:9-187), all modes in one pass
// =====================================================================================
struct B2Grids {
    const double2 *g[6 * B2_MAX_MODES];   // [m][Er,Et,Ez,Br,Bt,Bz]
};

template <int NM, bool CUBIC>
__device__ __forceinline__ void b2_gather_one(const B2Cyl &c, const B2Grids &G, double rmax_gather,
                                              int Nz, int Nr, double F[6]) {
    double Fc[2][3] = {{0., 0., 0.}, {0., 0., 0.}};   // [E|B][r,t,z]
    if (c.r < rmax_gather) {
        if (!CUBIC) {
            int ir_l = (int)floor(c.r_cell), ir_u = ir_l + 1;
            int iz_l = (int)floor(c.z_cell), iz_u = iz_l + 1;
            double Sr_l = ir_u - c.r_cell, Sr_u = c.r_cell - ir_l;
            double Sz_l = iz_u - c.z_cell, Sz_u = c.z_cell - iz_l;
            double Sr_g = 0.;
            if (ir_l < 0) { Sr_g = Sr_l; Sr_l = 0.; ir_l = 0; }
            if (ir_l > Nr - 1) ir_l = Nr - 1;
            if (ir_u > Nr - 1) ir_u = Nr - 1;
            if (iz_l < 0) iz_l += Nz;
            if (iz_u < 0) iz_u += Nz;
            if (iz_l > Nz - 1) iz_l -= Nz;
            if (iz_u > Nz - 1) iz_u -= Nz;
            const double S_ll = Sz_l * Sr_l, S_lu = Sz_l * Sr_u, S_ul = Sz_u * Sr_l, S_uu = Sz_u * Sr_u;
            const double S_lg = Sz_l * Sr_g, S_ug = Sz_u * Sr_g;
            const bool on_axis = (ir_l == 0 && ir_u == 0);
            const size_t o_ll = (size_t)iz_l * Nr + ir_l, o_lu = (size_t)iz_l * Nr + ir_u;
            const size_t o_ul = (size_t)iz_u * Nr + ir_l, o_uu = (size_t)iz_u * Nr + ir_u;
            const size_t o_l0 = (size_t)iz_l * Nr, o_u0 = (size_t)iz_u * Nr;
            double e_re = 1., e_im = 0.;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double flip = (m & 1) ? -1. : 1.;
                const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double2 *a = G.g[6 * m + 3 * f + k];
                        double2 v;
                        double re = 0., im = 0.;
                        v = __ldg(a + o_ll); re += S_ll * v.x; im += S_ll * v.y;
                        v = __ldg(a + o_lu); re += S_lu * v.x; im += S_lu * v.y;
                        v = __ldg(a + o_ul); re += S_ul * v.x; im += S_ul * v.y;
                        v = __ldg(a + o_uu); re += S_uu * v.x; im += S_uu * v.y;
                        if (on_axis) {
                            const double sgn = (k == 2) ? flip : -flip;
                            v = __ldg(a + o_l0); re += sgn * S_lg * v.x; im += sgn * S_lg * v.y;
                            v = __ldg(a + o_u0); re += sgn * S_ug * v.x; im += sgn * S_ug * v.y;
                        }
                        Fc[f][k] += factor * (re * e_re - im * e_im);
                    }
                }
                const double nr = e_re * c.cs + e_im * c.sn, ni = e_im * c.cs - e_re * c.sn;
                e_re = nr; e_im = ni;
            }
        } else {
            double Sr[4], Sz[4];
            const int ir_lowest = (int)floor(c.r_cell) - 1;
            const double rl = c.r_cell - ir_lowest;
            Sr[0] = -1. / 6. * ((rl - 2.) * (rl - 2.) * (rl - 2.));
            Sr[1] = 1. / 6. * (3. * ((rl - 1.) * (rl - 1.) * (rl - 1.)) - 6. * ((rl - 1.) * (rl - 1.)) + 4.);
            Sr[2] = 1. / 6. * (3. * ((2. - rl) * (2. - rl) * (2. - rl)) - 6. * ((2. - rl) * (2. - rl)) + 4.);
            Sr[3] = -1. / 6. * ((1. - rl) * (1. - rl) * (1. - rl));
            const int iz_lowest = (int)floor(c.z_cell) - 1;
            const double zl = c.z_cell - iz_lowest;
            Sz[0] = -1. / 6. * ((zl - 2.) * (zl - 2.) * (zl - 2.));
            Sz[1] = 1. / 6. * (3. * ((zl - 1.) * (zl - 1.) * (zl - 1.)) - 6. * ((zl - 1.) * (zl - 1.)) + 4.);
            Sz[2] = 1. / 6. * (3. * ((2. - zl) * (2. - zl) * (2. - zl)) - 6. * ((2. - zl) * (2. - zl)) + 4.);
            Sz[3] = -1. / 6. * ((1. - zl) * (1. - zl) * (1. - zl));
            int irs[4], izs[4];
            bool neg[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int ir = ir_lowest + b;
                neg[b] = (ir < 0);
                if (ir < 0) ir = -ir - 1;
                else if (ir > Nr - 1) ir = Nr - 1;
                irs[b] = ir;
                int iz = iz_lowest + b;
                if (iz < 0) iz += Nz;
                else if (iz > Nz - 1) iz -= Nz;
                izs[b] = iz;
            }
            double e_re = 1., e_im = 0.;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double flip = (m & 1) ? -1. : 1.;
                const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    double acc[3][2] = {{0., 0.}, {0., 0.}, {0., 0.}};
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double sr_long = neg[b] ? Sr[b] * flip : Sr[b];
                        const double sr_perp = neg[b] ? -Sr[b] * flip : Sr[b];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const size_t o = (size_t)izs[a] * Nr + irs[b];
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const double w = Sz[a] * ((k == 2) ? sr_long : sr_perp);
                                const double2 v = __ldg(G.g[6 * m + 3 * f + k] + o);
                                acc[k][0] += w * v.x;
                                acc[k][1] += w * v.y;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) Fc[f][k] += factor * (acc[k][0] * e_re - acc[k][1] * e_im);
                }
                const double nr = e_re * c.cs + e_im * c.sn, ni = e_im * c.cs - e_re * c.sn;
                e_re = nr; e_im = ni;
            }
        }
    }
    F[0] = c.cs * Fc[0][0] - c.sn * Fc[0][1];
    F[1] = c.sn * Fc[0][0] + c.cs * Fc[0][1];
    F[2] = Fc[0][2];
    F[3] = c.cs * Fc[1][0] - c.sn * Fc[1][1];
    F[4] = c.sn * Fc[1][0] + c.cs * Fc[1][1];
    F[5] = Fc[1][2];
}

template <int NM, bool CUBIC>
__global__ void __launch_bounds__(256)
k_gather(int64_t n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
         double rmax_gather, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, B2Grids G,
         double *__restrict__ Ex, double *__restrict__ Ey, double *__restrict__ Ez,
         double *__restrict__ Bx, double *__restrict__ By, double *__restrict__ Bz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    B2Cyl c = b2_cyl(x[i], y[i], z[i], invdz, zmin, invdr, rmin);
    double F[6];
    b2_gather_one<NM, CUBIC>(c, G, rmax_gather, Nz, Nr, F);
    Ex[i] = F[0]; Ey[i] = F[1]; Ez[i] = F[2];
    Bx[i] = F[3]; By[i] = F[4]; Bz[i] = F[5];
}

// =====================================================================================
// Vay pusher (push/inline_functions.py:11-48) and position push (push/cuda_methods.py:17-52)
// =====================================================================================
__global__ void k_push_p(int64_t n, double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
                         double *__restrict__ inv_gamma, const double *__restrict__ Ex,
                         const double *__restrict__ Ey, const double *__restrict__ Ez,
                         const double *__restrict__ Bx, const double *__restrict__ By,
                         const double *__restrict__ Bz, double econst, double bconst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double F[6] = {Ex[i], Ey[i], Ez[i], Bx[i], By[i], Bz[i]};
    double a = ux[i], b = uy[i], cz = uz[i], g = inv_gamma[i];
    b2_vay(a, b, cz, g, F, econst, bconst);
    ux[i] = a; uy[i] = b; uz[i] = cz; inv_gamma[i] = g;
}

__global__ void k_push_x(int64_t n, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                         const double *__restrict__ ux, const double *__restrict__ uy,
                         const double *__restrict__ uz, const double *__restrict__ inv_gamma,
                         double chdt, double xp, double yp, double zp) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double g = inv_gamma[i];
    x[i] += chdt * g * xp * ux[i];
    y[i] += chdt * g * yp * uy[i];
    z[i] += chdt * g * zp * uz[i];
}

// fused gather + push_p + push_x: one read and one write of the particle state per step.
// Optionally also emits the cell key of the NEW position (on a grid whose zmin is key_zmin: the
// Galilean scheme shifts the grid between the push and the deposition), saving the separate
// get_cell_idx_per_particle pass of the following sort.
template <int NM, bool CUBIC>
__global__ void __launch_bounds__(256)
k_gather_push(int64_t n, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
              double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
              double *__restrict__ inv_gamma, double rmax_gather, double invdz, double zmin, int Nz,
              double invdr, double rmin, int Nr, B2Grids G, double econst, double bconst, double chdt,
              int32_t *__restrict__ cell_idx, double key_zmin) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xj = x[i], yj = y[i], zj = z[i];
    B2Cyl c = b2_cyl(xj, yj, zj, invdz, zmin, invdr, rmin);
    double F[6];
    b2_gather_one<NM, CUBIC>(c, G, rmax_gather, Nz, Nr, F);
    double a = ux[i], b = uy[i], cz = uz[i], g = inv_gamma[i];
    b2_vay(a, b, cz, g, F, econst, bconst);
    ux[i] = a; uy[i] = b; uz[i] = cz; inv_gamma[i] = g;
    const double xn = xj + chdt * g * 1. * a;
    const double yn = yj + chdt * g * 1. * b;
    const double zn = zj + chdt * g * 1. * cz;
    x[i] = xn; y[i] = yn; z[i] = zn;
    if (cell_idx) cell_idx[i] = b2_cell_of(b2_cyl(xn, yn, zn, invdz, key_zmin, invdr, rmin), Nz, Nr);
}

// ---- tiled variant (linear shapes) -------------------------------------------------------------
// When the particles are cell-sorted, the 128 particles of a CTA touch a handful of neighbouring
// cells.  The CTA finds the bounding box of its stencils (integer min/max in shared memory), stages
// that (rows x cols) tile of all 6*NM field arrays in shared memory with coalesced 16-byte loads, and
// every thread then gathers from shared memory: ~50 dependent global loads per particle become a few
// coalesced tile loads per thread.  A CTA whose bounding box does not fit (unsorted particles, or a
// z-row boundary of the sorted order) falls back to global loads; results are identical either way.
#define GP_TPB 128
#define GP_TILE_CELLS 96
#ifndef GP_MIN_CTAS
#define GP_MIN_CTAS 8
#endif

// (Measured alternatives that lost: loading the momenta together with the positions -- 80 registers,
// 6 CTAs/SM, 0.79 vs 0.72 ms at C2 -- and prefetch.global.L2 of the momenta, 0.77 ms.)
template <int NM>
__global__ void __launch_bounds__(GP_TPB, GP_MIN_CTAS)
k_gather_push_tiled(int64_t n, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                    double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
                    double *__restrict__ inv_gamma, double rmax_gather, double invdz, double zmin, int Nz,
                    double invdr, double rmin, int Nr, B2Grids G, double econst, double bconst, double chdt,
                    int32_t *__restrict__ cell_idx, double key_zmin) {
    __shared__ double2 tile[6 * NM][GP_TILE_CELLS];
    __shared__ int4 s_boxw[GP_TPB / 32];     // per warp: min iz_l (unwrapped), max iz_u (unwrapped), min ir, max ir (clamped)
    const int tid = threadIdx.x;
    const int64_t i = blockIdx.x * (int64_t)GP_TPB + tid;
    const bool in_range = i < n;
    double xj = 0., yj = 0., zj = 0.;
    if (in_range) { xj = x[i]; yj = y[i]; zj = z[i]; }
    const B2Cyl c = b2_cyl(xj, yj, zj, invdz, zmin, invdr, rmin);
    const bool active = in_range && (c.r < rmax_gather);
    // stencil indices and weights exactly as gather_field_gpu_linear (gathering/cuda_methods.py:109-160)
    int ir_l = (int)floor(c.r_cell), ir_u = ir_l + 1;
    const int iz_l0 = (int)floor(c.z_cell);              // unwrapped, in [-1, Nz-1] for in-box particles
    double Sr_l = ir_u - c.r_cell, Sr_u = c.r_cell - ir_l;
    const double Sz_l = (iz_l0 + 1) - c.z_cell, Sz_u = c.z_cell - iz_l0;
    double Sr_g = 0.;
    if (ir_l < 0) { Sr_g = Sr_l; Sr_l = 0.; ir_l = 0; }
    if (ir_l > Nr - 1) ir_l = Nr - 1;
    if (ir_u > Nr - 1) ir_u = Nr - 1;
    // bounding box of the stencils of the particles that lie NEAR their warp's first particle (sorted particles all
    // do); a stray particle (moved far since the last sort) gathers from global memory on its own instead of blowing
    // up the tile for the whole CTA.  Warp-level integer reductions, one shared-memory slot per warp and one barrier:
    // the 4 x 128 shared atomics on four addresses plus the barrier for the CTA-wide anchor took as long as the
    // particle loads (ncu: barrier stall 4.7 cycles per issue).
    const int lane = tid & 31, warp_id = tid >> 5;
    const int a_z = __shfl_sync(0xffffffffu, iz_l0, 0), a_r = __shfl_sync(0xffffffffu, ir_l, 0);
    const bool near = active && abs(iz_l0 - a_z) <= 2 && ir_l >= a_r - 8 && ir_u <= a_r + 40;
    {
        const int b0 = __reduce_min_sync(0xffffffffu, near ? iz_l0 : INT_MAX);
        const int b1 = __reduce_max_sync(0xffffffffu, near ? iz_l0 + 1 : INT_MIN);
        const int b2 = __reduce_min_sync(0xffffffffu, near ? ir_l : INT_MAX);
        const int b3 = __reduce_max_sync(0xffffffffu, near ? ir_u : INT_MIN);
        if (lane == 0) s_boxw[warp_id] = make_int4(b0, b1, b2, b3);
    }
    __syncthreads();
    int4 box = s_boxw[0];
#pragma unroll
    for (int w = 1; w < GP_TPB / 32; ++w) {
        const int4 o = s_boxw[w];
        box.x = min(box.x, o.x); box.y = max(box.y, o.y); box.z = min(box.z, o.z); box.w = max(box.w, o.w);
    }
    const int z0 = box.x, r0 = box.z;
    const bool box_ok = (box.y >= box.x) && (box.w >= box.z);   // some particle is near
    const int nrow = box_ok ? box.y - z0 + 1 : 0, ncol = box_ok ? box.w - r0 + 1 : 0;
    const bool tile_ok = box_ok && (nrow * ncol <= GP_TILE_CELLS) && z0 >= -1 && box.y <= Nz;
    const bool use_tile = tile_ok && near;          // per thread
    if (tile_ok && tid < nrow * ncol) {
        // one thread per tile cell (GP_TILE_CELLS <= GP_TPB), 6*NM coalesced 16-byte loads each
        const int row = tid / ncol, col = tid - row * ncol;
        int iz = z0 + row;
        if (iz < 0) iz += Nz;
        if (iz > Nz - 1) iz -= Nz;
        const size_t o = (size_t)iz * Nr + r0 + col;
#pragma unroll
        for (int a = 0; a < 6 * NM; ++a) tile[a][tid] = __ldg(G.g[a] + o);
    }
    __syncthreads();
    if (!in_range) return;

    double Fc[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
    // Two copies of the stencil sum: when every gathering lane of the warp reads the tile (the usual case) the
    // warp runs the copy that has no global loads and none of their address arithmetic; a predicated
    // `tile ? LDS : LDG` select would issue both for every warp (48 dead LDG + ~100 address instructions per
    // warp, a fifth of the kernel -- profiles/r02_gather_pipe_vs_tiled.md).
    const unsigned live = __activemask();
    const bool warp_on_tile = __all_sync(live, use_tile || !active);
    // (the two guard-cell terms of a particle within half a cell of the axis: skipped as a whole by the warps --
    //  all but one in Nr -- that hold no such particle, instead of 24 predicated-off loads and 48 fp64 operations)
    const bool warp_near_axis = __any_sync(live, active && ir_l == 0 && ir_u == 0);
    if (active) {
        const double S_ll = Sz_l * Sr_l, S_lu = Sz_l * Sr_u, S_ul = Sz_u * Sr_l, S_uu = Sz_u * Sr_u;
        const double S_lg = Sz_l * Sr_g, S_ug = Sz_u * Sr_g;
        const bool on_axis = (ir_l == 0 && ir_u == 0);
        if (warp_on_tile) {
            const int t_ll = (iz_l0 - z0) * ncol + (ir_l - r0), t_lu = (iz_l0 - z0) * ncol + (ir_u - r0);
            const int t_ul = t_ll + ncol, t_uu = t_lu + ncol;
            const int t_l0 = (iz_l0 - z0) * ncol - r0, t_u0 = t_l0 + ncol;    // column 0 (only if on_axis: r0 == 0)
            // (warp-uniform choice: only the warps that hold a particle next to the axis carry the guard terms)
            if (warp_near_axis)
                b2_tile_sum<NM, true, GP_TILE_CELLS>(tile, t_ll, t_lu, t_ul, t_uu, t_l0, t_u0, S_ll, S_lu, S_ul, S_uu,
                                                      S_lg, S_ug, on_axis, c.cs, c.sn, Fc);
            else
                b2_tile_sum<NM, false, GP_TILE_CELLS>(tile, t_ll, t_lu, t_ul, t_uu, t_l0, t_u0, S_ll, S_lu, S_ul, S_uu,
                                                       S_lg, S_ug, false, c.cs, c.sn, Fc);
        } else {
            int iz_l = iz_l0, iz_u = iz_l0 + 1;
            if (iz_l < 0) iz_l += Nz;
            if (iz_u < 0) iz_u += Nz;
            if (iz_l > Nz - 1) iz_l -= Nz;
            if (iz_u > Nz - 1) iz_u -= Nz;
            // tile-relative cell offsets (used when use_tile) and global offsets (fallback)
            const int t_ll = (iz_l0 - z0) * ncol + (ir_l - r0), t_lu = (iz_l0 - z0) * ncol + (ir_u - r0);
            const int t_ul = t_ll + ncol, t_uu = t_lu + ncol;
            const int t_l0 = (iz_l0 - z0) * ncol - r0, t_u0 = t_l0 + ncol;
            const size_t o_ll = (size_t)iz_l * Nr + ir_l, o_lu = (size_t)iz_l * Nr + ir_u;
            const size_t o_ul = (size_t)iz_u * Nr + ir_l, o_uu = (size_t)iz_u * Nr + ir_u;
            const size_t o_l0 = (size_t)iz_l * Nr, o_u0 = (size_t)iz_u * Nr;
            double e_re = 1., e_im = 0.;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double flip = (m & 1) ? -1. : 1.;
                const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int a = 6 * m + 3 * f + k;
                        double2 v_ll, v_lu, v_ul, v_uu;
                        if (use_tile) {
                            v_ll = tile[a][t_ll]; v_lu = tile[a][t_lu]; v_ul = tile[a][t_ul]; v_uu = tile[a][t_uu];
                        } else {
                            const double2 *g = G.g[a];
                            v_ll = __ldg(g + o_ll); v_lu = __ldg(g + o_lu); v_ul = __ldg(g + o_ul); v_uu = __ldg(g + o_uu);
                        }
                        double re = 0., im = 0.;
                        re += S_ll * v_ll.x; im += S_ll * v_ll.y;
                        re += S_lu * v_lu.x; im += S_lu * v_lu.y;
                        re += S_ul * v_ul.x; im += S_ul * v_ul.y;
                        re += S_uu * v_uu.x; im += S_uu * v_uu.y;
                        if (on_axis) {
                            const double sgn = (k == 2) ? flip : -flip;
                            double2 v_l0, v_u0;
                            if (use_tile) { v_l0 = tile[a][t_l0]; v_u0 = tile[a][t_u0]; }
                            else { v_l0 = __ldg(G.g[a] + o_l0); v_u0 = __ldg(G.g[a] + o_u0); }
                            re += sgn * S_lg * v_l0.x; im += sgn * S_lg * v_l0.y;
                            re += sgn * S_ug * v_u0.x; im += sgn * S_ug * v_u0.y;
                        }
                        Fc[f][k] += factor * (re * e_re - im * e_im);
                    }
                }
                const double nr = e_re * c.cs + e_im * c.sn, ni = e_im * c.cs - e_re * c.sn;
                e_re = nr; e_im = ni;
            }
        }
    }
    double F[6];
    F[0] = c.cs * Fc[0][0] - c.sn * Fc[0][1];
    F[1] = c.sn * Fc[0][0] + c.cs * Fc[0][1];
    F[2] = Fc[0][2];
    F[3] = c.cs * Fc[1][0] - c.sn * Fc[1][1];
    F[4] = c.sn * Fc[1][0] + c.cs * Fc[1][1];
    F[5] = Fc[1][2];
    double a = ux[i], b = uy[i], cz = uz[i], g = inv_gamma[i];
    b2_vay(a, b, cz, g, F, econst, bconst);
    ux[i] = a; uy[i] = b; uz[i] = cz; inv_gamma[i] = g;
    const double xn = xj + chdt * g * 1. * a;
    const double yn = yj + chdt * g * 1. * b;
    const double zn = zj + chdt * g * 1. * cz;
    x[i] = xn; y[i] = yn; z[i] = zn;
    if (cell_idx) cell_idx[i] = b2_cell_of(b2_cyl(xn, yn, zn, invdz, key_zmin, invdr, rmin), Nz, Nr);
}

// push_x + optional periodic wrap of z into [wrap_zmin, wrap_zmax) + optional cell key of the new
// position (push/cuda_methods.py:17-52, particle_buffer_handling.py:637-658, cuda_sorting.py:22-88)
__global__ void k_push_x_key(int64_t n, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                             const double *__restrict__ ux, const double *__restrict__ uy,
                             const double *__restrict__ uz, const double *__restrict__ inv_gamma,
                             double chdt, int wrap, double wrap_zmin, double wrap_zmax,
                             double invdz, double key_zmin, int Nz, double invdr, double rmin, int Nr,
                             int32_t *__restrict__ cell_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double g = inv_gamma[i];
    const double xn = x[i] + chdt * g * 1. * ux[i];
    const double yn = y[i] + chdt * g * 1. * uy[i];
    double zn = z[i] + chdt * g * 1. * uz[i];
    if (wrap) {
        const double l_box = wrap_zmax - wrap_zmin;
        while (zn >= wrap_zmax) zn -= l_box;
        while (zn < wrap_zmin) zn += l_box;
    }
    x[i] = xn; y[i] = yn; z[i] = zn;
    if (cell_idx) cell_idx[i] = b2_cell_of(b2_cyl(xn, yn, zn, invdz, key_zmin, invdr, rmin), Nz, Nr);
}

__global__ void k_shift_periodic(int64_t n, double *__restrict__ z, double zmin, double zmax) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double l_box = zmax - zmin;
    double zi = z[i];
    while (zi >= zmax) zi -= l_box;
    while (zi < zmin) zi += l_box;
    z[i] = zi;
}

__global__ void k_add_scalar(int64_t n, double *__restrict__ v, double a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] += a;
}

// =====================================================================================
// particle exchange: 3-way stable partition by z (remove_particles_cpu semantics,
// fbpic/boundaries/particle_buffer_handling.py:58-175: left if z < zlo, right if z > zhi, the rest
// stays, relative order preserved)
// =====================================================================================
__global__ void k_classify_z(int64_t n, const double *__restrict__ z, double zlo, double zhi,
                             int32_t *__restrict__ f_stay, int32_t *__restrict__ f_left,
                             int32_t *__restrict__ f_right) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    const int l = zi < zlo, r = zi > zhi;
    f_left[i] = l; f_right[i] = r; f_stay[i] = !(l || r);
}
struct B2Part {
    const double *src[B2_MAX_ARRAYS];
    double *stay[B2_MAX_ARRAYS], *left[B2_MAX_ARRAYS], *right[B2_MAX_ARRAYS];
};
// after the exclusive scans p_* hold the destination index inside each class; the class of i is
// recovered from the scans themselves (p[i+1] - p[i], or the known total for the last element)
__global__ void k_partition_scatter(int64_t n, const int32_t *__restrict__ p_stay,
                                    const int32_t *__restrict__ p_left, const int32_t *__restrict__ p_right,
                                    const double *__restrict__ z, double zlo, double zhi, B2Part a, int n_arrays) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zi = z[i];
    const int l = zi < zlo, r = zi > zhi;
    for (int k = 0; k < n_arrays; ++k) {
        const double v = a.src[k][i];
        if (l) { if (a.left[k]) a.left[k][p_left[i]] = v; }
        else if (r) { if (a.right[k]) a.right[k][p_right[i]] = v; }
        else a.stay[k][p_stay[i]] = v;
    }
}

// =====================================================================================
// C ABI
// =====================================================================================

// ---- cubic shapes: the same tiling for the 4 x 4 stencil -------------------------------------------------------
// (gather_field_gpu_cubic, gathering/cuda_methods.py:209-354: the reference walks the 16 points of every mode array
//  in global memory per particle.)  One CTA = 128 consecutive (sorted) particles; the bounding box of their stencils
//  -- rows iz_lowest .. iz_lowest + 3 unwrapped, columns after the reflection at the axis and the clamp at Nr - 1 --
//  goes to shared memory once, every mode array; a CTA whose box does not fit GPC_TILE_CELLS cells (it straddles a
//  z-row boundary of the sorted order) or a stray particle walks global memory like b2_gather_one.
template <int NM>
__global__ void __launch_bounds__(GP_TPB, 4)
k_gather_push_tiled_cubic(int64_t n, double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                          double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
                          double *__restrict__ inv_gamma, double rmax_gather, double invdz, double zmin, int Nz,
                          double invdr, double rmin, int Nr, B2Grids G, double econst, double bconst, double chdt,
                          int32_t *__restrict__ cell_idx, double key_zmin) {
    constexpr int GPC_TILE_CELLS = NM <= 3 ? 128 : 120;       // (48 KB of static shared memory at Nm = 4)
    static_assert(GPC_TILE_CELLS <= GP_TPB, "one thread loads one tile cell");
    __shared__ double2 tile[6 * NM][GPC_TILE_CELLS];
    __shared__ int4 s_boxw[GP_TPB / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp_id = tid >> 5;
    const int64_t i = blockIdx.x * (int64_t)GP_TPB + tid;
    const bool in_range = i < n;
    double xj = 0., yj = 0., zj = 0.;
    if (in_range) { xj = x[i]; yj = y[i]; zj = z[i]; }
    const B2Cyl c = b2_cyl(xj, yj, zj, invdz, zmin, invdr, rmin);
    const bool active = in_range && (c.r < rmax_gather);
    const int ir_lowest = (int)floor(c.r_cell) - 1, iz_lowest = (int)floor(c.z_cell) - 1;
    int irs[4];
    bool neg[4];
    int ir_min = INT_MAX, ir_max = INT_MIN;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int ir = ir_lowest + b;
        neg[b] = (ir < 0);
        if (ir < 0) ir = -ir - 1;
        else if (ir > Nr - 1) ir = Nr - 1;
        irs[b] = ir;
        ir_min = min(ir_min, ir); ir_max = max(ir_max, ir);
    }
    const int a_z = __shfl_sync(0xffffffffu, iz_lowest, 0), a_r = __shfl_sync(0xffffffffu, ir_min, 0);
    const bool near = active && abs(iz_lowest - a_z) <= 2 && ir_min >= a_r - 8 && ir_max <= a_r + 40;
    {
        const int b0 = __reduce_min_sync(0xffffffffu, near ? iz_lowest : INT_MAX);
        const int b1 = __reduce_max_sync(0xffffffffu, near ? iz_lowest + 3 : INT_MIN);
        const int b2 = __reduce_min_sync(0xffffffffu, near ? ir_min : INT_MAX);
        const int b3 = __reduce_max_sync(0xffffffffu, near ? ir_max : INT_MIN);
        if (lane == 0) s_boxw[warp_id] = make_int4(b0, b1, b2, b3);
    }
    __syncthreads();
    int4 box = s_boxw[0];
#pragma unroll
    for (int w = 1; w < GP_TPB / 32; ++w) {
        const int4 o = s_boxw[w];
        box.x = min(box.x, o.x); box.y = max(box.y, o.y); box.z = min(box.z, o.z); box.w = max(box.w, o.w);
    }
    const int z0 = box.x, r0 = box.z;
    const bool box_ok = (box.y >= box.x) && (box.w >= box.z);
    const int nrow = box_ok ? box.y - z0 + 1 : 0, ncol = box_ok ? box.w - r0 + 1 : 0;
    // (rows may hang over either end of the periodic box by the stencil reach: wrapped once by the loader)
    const bool tile_ok = box_ok && (nrow * ncol <= GPC_TILE_CELLS) && z0 >= -Nz && box.y < 2 * Nz;
    const bool use_tile = tile_ok && near;
    if (tile_ok && tid < nrow * ncol) {
        const int row = tid / ncol, col = tid - row * ncol;
        int iz = z0 + row;
        if (iz < 0) iz += Nz;
        if (iz > Nz - 1) iz -= Nz;
        const size_t o = (size_t)iz * Nr + r0 + col;
#pragma unroll
        for (int a = 0; a < 6 * NM; ++a) tile[a][tid] = __ldg(G.g[a] + o);
    }
    __syncthreads();
    if (!in_range) return;
    double F[6];
    const unsigned live = __activemask();
    const bool warp_on_tile = __all_sync(live, use_tile || !active);
    if (warp_on_tile) {
        double Fc[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
        if (active) {
            double Sr[4], Sz[4];
            const double rl = c.r_cell - ir_lowest;
            Sr[0] = -1. / 6. * ((rl - 2.) * (rl - 2.) * (rl - 2.));
            Sr[1] = 1. / 6. * (3. * ((rl - 1.) * (rl - 1.) * (rl - 1.)) - 6. * ((rl - 1.) * (rl - 1.)) + 4.);
            Sr[2] = 1. / 6. * (3. * ((2. - rl) * (2. - rl) * (2. - rl)) - 6. * ((2. - rl) * (2. - rl)) + 4.);
            Sr[3] = -1. / 6. * ((1. - rl) * (1. - rl) * (1. - rl));
            const double zl = c.z_cell - iz_lowest;
            Sz[0] = -1. / 6. * ((zl - 2.) * (zl - 2.) * (zl - 2.));
            Sz[1] = 1. / 6. * (3. * ((zl - 1.) * (zl - 1.) * (zl - 1.)) - 6. * ((zl - 1.) * (zl - 1.)) + 4.);
            Sz[2] = 1. / 6. * (3. * ((2. - zl) * (2. - zl) * (2. - zl)) - 6. * ((2. - zl) * (2. - zl)) + 4.);
            Sz[3] = -1. / 6. * ((1. - zl) * (1. - zl) * (1. - zl));
            const int t0 = (iz_lowest - z0) * ncol - r0;       // tile index of (row a, column ir) = t0 + a*ncol + ir
            double e_re = 1., e_im = 0.;
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double flip = (m & 1) ? -1. : 1.;
                const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    double acc[3][2] = {{0., 0.}, {0., 0.}, {0., 0.}};
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const double sr_long = neg[b] ? Sr[b] * flip : Sr[b];
                        const double sr_perp = neg[b] ? -Sr[b] * flip : Sr[b];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const int t = t0 + a * ncol + irs[b];
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const double w = Sz[a] * ((k == 2) ? sr_long : sr_perp);
                                const double2 v = tile[6 * m + 3 * f + k][t];
                                acc[k][0] += w * v.x;
                                acc[k][1] += w * v.y;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) Fc[f][k] += factor * (acc[k][0] * e_re - acc[k][1] * e_im);
                }
                const double nr = e_re * c.cs + e_im * c.sn, ni = e_im * c.cs - e_re * c.sn;
                e_re = nr; e_im = ni;
            }
        }
        F[0] = c.cs * Fc[0][0] - c.sn * Fc[0][1];
        F[1] = c.sn * Fc[0][0] + c.cs * Fc[0][1];
        F[2] = Fc[0][2];
        F[3] = c.cs * Fc[1][0] - c.sn * Fc[1][1];
        F[4] = c.sn * Fc[1][0] + c.cs * Fc[1][1];
        F[5] = Fc[1][2];
    } else {
        b2_gather_one<NM, true>(c, G, rmax_gather, Nz, Nr, F);
    }
    double a = ux[i], b = uy[i], cz = uz[i], g = inv_gamma[i];
    b2_vay(a, b, cz, g, F, econst, bconst);
    ux[i] = a; uy[i] = b; uz[i] = cz; inv_gamma[i] = g;
    const double xn = xj + chdt * g * 1. * a;
    const double yn = yj + chdt * g * 1. * b;
    const double zn = zj + chdt * g * 1. * cz;
    x[i] = xn; y[i] = yn; z[i] = zn;
    if (cell_idx) cell_idx[i] = b2_cell_of(b2_cyl(xn, yn, zn, invdz, key_zmin, invdr, rmin), Nz, Nr);
}

template <int NM>
static void launch_gather(bool cubic, unsigned g, cudaStream_t s, int64_t n, const double *x, const double *y,
                          const double *z, double rg, double invdz, double zmin, int Nz, double invdr, double rmin,
                          int Nr, const B2Grids &G, double *Ex, double *Ey, double *Ez, double *Bx, double *By,
                          double *Bz) {
    if (cubic) k_gather<NM, true><<<g, 256, 0, s>>>(n, x, y, z, rg, invdz, zmin, Nz, invdr, rmin, Nr, G, Ex, Ey, Ez, Bx, By, Bz);
    else k_gather<NM, false><<<g, 256, 0, s>>>(n, x, y, z, rg, invdz, zmin, Nz, invdr, rmin, Nr, G, Ex, Ey, Ez, Bx, By, Bz);
}
template <int NM>
static void launch_gather_push(bool cubic, unsigned g, cudaStream_t s, int64_t n, double *x, double *y, double *z,
                               double *ux, double *uy, double *uz, double *ig, double rg, double invdz, double zmin,
                               int Nz, double invdr, double rmin, int Nr, const B2Grids &G, double ec, double bc,
                               double chdt, int32_t *cell_idx, double key_zmin) {
    // cubic shapes: tiled as well (B2_GATHER_CUBIC=plain: the thread-per-particle walk through global memory)
    static const bool plain = []() { const char *e = getenv("B2_GATHER_CUBIC"); return e && !strcmp(e, "plain"); }();
    if (cubic && plain) k_gather_push<NM, true><<<g, 256, 0, s>>>(n, x, y, z, ux, uy, uz, ig, rg, invdz, zmin, Nz, invdr, rmin, Nr, G, ec, bc, chdt, cell_idx, key_zmin);
    else if (cubic) k_gather_push_tiled_cubic<NM><<<grid1d(n, GP_TPB), GP_TPB, 0, s>>>(n, x, y, z, ux, uy, uz, ig, rg, invdz, zmin, Nz, invdr, rmin, Nr, G, ec, bc, chdt, cell_idx, key_zmin);
    else k_gather_push_tiled<NM><<<grid1d(n, GP_TPB), GP_TPB, 0, s>>>(n, x, y, z, ux, uy, uz, ig, rg, invdz, zmin, Nz, invdr, rmin, Nr, G, ec, bc, chdt, cell_idx, key_zmin);
}

extern "C" {

int b2_cell_index(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, double invdz,
                  double zmin, int Nz, double invdr, double rmin, int Nr, int32_t *cell_idx, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_CELL_INDEX, b2_stream_of(ctx, stream));
    k_cell_index<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, x, y, z, invdz, zmin, Nz, invdr, rmin, Nr, cell_idx);
    B2_LAUNCHED();
    return 0;
}

int b2_sort_cells(b2_ctx *ctx, int64_t n, int32_t *cell_idx, int64_t *sorted_idx, int32_t *prefix_sum,
                  int Nz, int Nr, void *stream) {
    cudaStream_t s = b2_stream_of(ctx, stream);
    const int ncells = Nz * (Nr + 1);
    if (n <= 0) {
        // an empty species still counts as sorted (cuda_sorting.py:106,181 skip the launches, the
        // prefix sum stays all-zero): record it so that a following permute / deposit_permute matches
        B2_CUDA(cudaMemsetAsync(prefix_sum, 0, sizeof(int32_t) * (size_t)ncells, s));
        ctx->last_sort_n = 0;
        ctx->last_sort_prefix = prefix_sum;
        return 0;
    }
    // 32-bit particle indices (CUB num_items, the idx32 permutation, the int32 prefix sums): one species on one
    // GPU may hold up to 2^31 - 1 macroparticles (137 GB of SoA state -- more than fits next to the sort buffers)
    if (n > (int64_t)INT32_MAX) return b2_fail(-3, "b2_sort_cells: more than 2^31-1 particles in one species on one GPU", __FILE__, __LINE__);
    B2Prof prof_(B2P_SORT, s);
    int end_bit = 1;
    while ((1LL << end_bit) < (long long)ncells && end_bit < 31) ++end_bit;
    size_t temp_bytes = 0;
    int32_t *nul = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, nul, nul, nul, nul, (int)n, 0, end_bit, s);
    // scratch 0: [keys_sorted | idx_in | idx_sorted | cub temp]
    const size_t na = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255;
    void *base;
    int rc = b2_scratch(ctx, 0, 3 * na + temp_bytes, &base);
    if (rc) return rc;
    int32_t *keys_sorted = (int32_t *)base;
    int32_t *idx_in = (int32_t *)((char *)base + na);
    int32_t *idx_sorted = (int32_t *)((char *)base + 2 * na);
    void *temp = (char *)base + 3 * na;
    k_iota<<<grid1d(n, 256), 256, 0, s>>>(n, idx_in);
    B2_LAUNCHED();
    B2_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const int32_t *)cell_idx, keys_sorted,
                                            (const int32_t *)idx_in, idx_sorted, (int)n, 0, end_bit, s));
    g_b2_launches.fetch_add(4);
    ctx->last_idx32 = idx_sorted; ctx->last_keys_sorted = keys_sorted; ctx->last_sort_n = n;
    ctx->last_sort_prefix = prefix_sum;
    if (sorted_idx) {     // materialise the API-visible int64 permutation and the sorted keys
        k_finish_sort<<<grid1d(n, 256), 256, 0, s>>>(n, keys_sorted, idx_sorted, cell_idx, sorted_idx);
        B2_LAUNCHED();
    }
    k_prefix_sum<<<grid1d(ncells, 256), 256, 0, s>>>(n, keys_sorted, prefix_sum, ncells);
    B2_LAUNCHED();
    return 0;
}

int b2_permute(b2_ctx *ctx, int64_t n, const int64_t *sorted_idx, int n_arrays, const double *const *src,
               double *const *dst, void *stream) {
    if (n <= 0 || n_arrays <= 0) return 0;
    B2Prof prof_(B2P_PERMUTE, b2_stream_of(ctx, stream));
    if (n_arrays > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    B2Perm a;
    for (int k = 0; k < n_arrays; ++k) { a.src[k] = src[k]; a.dst[k] = dst[k]; }
    cudaStream_t s = b2_stream_of(ctx, stream);
    const int32_t *i32 = nullptr;
    if (!sorted_idx) {
        if (!ctx->last_idx32 || ctx->last_sort_n != n)
            return b2_fail(-4, "b2_permute: no sorted_idx given and no matching sort in this context", __FILE__, __LINE__);
        i32 = ctx->last_idx32;
    }
    if (n_arrays == 8) k_permute<8><<<grid1d(n, 256), 256, 0, s>>>(n, sorted_idx, i32, a, n_arrays);
    else if (n_arrays == 14) k_permute<14><<<grid1d(n, 256), 256, 0, s>>>(n, sorted_idx, i32, a, n_arrays);
    else k_permute<0><<<grid1d(n, 256), 256, 0, s>>>(n, sorted_idx, i32, a, n_arrays);
    B2_LAUNCHED();
    return 0;
}

int b2_gather(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, double rmax_gather,
              double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
              const void *const *grids, int cubic, double *Ex, double *Ey, double *Ez, double *Bx, double *By,
              double *Bz, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_GATHER, b2_stream_of(ctx, stream));
    if (Nm < 1 || Nm > 4) return b2_fail(-3, "b2_gather: Nm must be in 1..4", __FILE__, __LINE__);
    B2Grids G;
    for (int k = 0; k < 6 * Nm; ++k) G.g[k] = (const double2 *)grids[k];
    cudaStream_t s = b2_stream_of(ctx, stream);
    unsigned g = grid1d(n, 256);
#define B2_ARGS cubic != 0, g, s, n, x, y, z, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, G, Ex, Ey, Ez, Bx, By, Bz
    switch (Nm) {
        case 1: launch_gather<1>(B2_ARGS); break;
        case 2: launch_gather<2>(B2_ARGS); break;
        case 3: launch_gather<3>(B2_ARGS); break;
        default: launch_gather<4>(B2_ARGS); break;
    }
#undef B2_ARGS
    B2_LAUNCHED();
    return 0;
}

int b2_push_p(b2_ctx *ctx, int64_t n, double *ux, double *uy, double *uz, double *inv_gamma, const double *Ex,
              const double *Ey, const double *Ez, const double *Bx, const double *By, const double *Bz, double q,
              double m, double dt, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_PUSH, b2_stream_of(ctx, stream));
    const double econst = q * dt / (m * B2_C_LIGHT), bconst = 0.5 * q * dt / m;
    k_push_p<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, ux, uy, uz, inv_gamma, Ex, Ey, Ez, Bx, By, Bz, econst, bconst);
    B2_LAUNCHED();
    return 0;
}

int b2_push_x(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, const double *ux, const double *uy,
              const double *uz, const double *inv_gamma, double dt, double xp, double yp, double zp, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_PUSH, b2_stream_of(ctx, stream));
    k_push_x<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, x, y, z, ux, uy, uz, inv_gamma, B2_C_LIGHT * dt, xp, yp, zp);
    B2_LAUNCHED();
    return 0;
}

int b2_gather_push(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, double *ux, double *uy, double *uz,
                   double *inv_gamma, double rmax_gather, double invdz, double zmin, int Nz, double invdr,
                   double rmin, int Nr, int Nm, const void *const *grids, int cubic, double q, double m,
                   double dt_p, double dt_x, int32_t *cell_idx, double key_zmin, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_GATHER_PUSH, b2_stream_of(ctx, stream));
    if (Nm < 1 || Nm > 4) return b2_fail(-3, "b2_gather_push: Nm must be in 1..4", __FILE__, __LINE__);
    B2Grids G;
    for (int k = 0; k < 6 * Nm; ++k) G.g[k] = (const double2 *)grids[k];
    cudaStream_t s = b2_stream_of(ctx, stream);
    const double ec = q * dt_p / (m * B2_C_LIGHT), bc = 0.5 * q * dt_p / m, chdt = B2_C_LIGHT * dt_x;
    if (!cubic) {
        // linear shapes, B2_GATHER_IMPL=pipe only: the persistent kernel with TMA-staged field tiles
        // (b2_gather_pipe.cu) takes the full 128-particle chunks and reports how many it did; what is left (n % 128
        // particles -- or everything, by default) goes to the tiled kernels of this file
        int64_t done = 0;
        int rc = b2_gather_push_pipe(ctx, n, x, y, z, ux, uy, uz, inv_gamma, rmax_gather, invdz, zmin, Nz, invdr, rmin,
                                     Nr, Nm, grids, ec, bc, chdt, cell_idx, key_zmin, s, &done);
        if (rc) return rc;
        if (done >= n) return 0;
        x += done; y += done; z += done; ux += done; uy += done; uz += done; inv_gamma += done;
        if (cell_idx) cell_idx += done;
        n -= done;
    }
    unsigned g = grid1d(n, 256);
#define B2_ARGS cubic != 0, g, s, n, x, y, z, ux, uy, uz, inv_gamma, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, G, ec, bc, chdt, cell_idx, key_zmin
    switch (Nm) {
        case 1: launch_gather_push<1>(B2_ARGS); break;
        case 2: launch_gather_push<2>(B2_ARGS); break;
        case 3: launch_gather_push<3>(B2_ARGS); break;
        default: launch_gather_push<4>(B2_ARGS); break;
    }
#undef B2_ARGS
    B2_LAUNCHED();
    return 0;
}

int b2_push_x_key(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, const double *ux, const double *uy,
                  const double *uz, const double *inv_gamma, double dt, int wrap, double wrap_zmin, double wrap_zmax,
                  double invdz, double key_zmin, int Nz, double invdr, double rmin, int Nr, int32_t *cell_idx,
                  void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_PUSH, b2_stream_of(ctx, stream));
    k_push_x_key<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, x, y, z, ux, uy, uz, inv_gamma,
        B2_C_LIGHT * dt, wrap, wrap_zmin, wrap_zmax, invdz, key_zmin, Nz, invdr, rmin, Nr, cell_idx);
    B2_LAUNCHED();
    return 0;
}

int b2_exchange_classify(b2_ctx *ctx, int64_t n, const double *z, double zlo, double zhi, int64_t *h_counts3,
                         void *stream) {
    h_counts3[0] = h_counts3[1] = h_counts3[2] = 0;
    if (n <= 0) return 0;
    if (n > (int64_t)INT32_MAX) return b2_fail(-3, "b2_exchange_classify: more than 2^31-1 particles in one species on one GPU", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    // scratch 1: [f_stay | f_left | f_right | cub temp]; the flags are scanned in place
    const size_t na = ((size_t)(n + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    size_t temp_bytes = 0;
    int32_t *nul = nullptr;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, nul, nul, (int)n, s);
    void *base;
    int rc = b2_scratch(ctx, 1, 3 * na + temp_bytes + 64, &base);
    if (rc) return rc;
    int32_t *f[3] = {(int32_t *)base, (int32_t *)((char *)base + na), (int32_t *)((char *)base + 2 * na)};
    void *temp = (char *)base + 3 * na;
    int32_t last_flag[3], last_pos[3];
    k_classify_z<<<grid1d(n, 256), 256, 0, s>>>(n, z, zlo, zhi, f[0], f[1], f[2]);
    B2_LAUNCHED();
    for (int k = 0; k < 3; ++k)
        B2_CUDA(cudaMemcpyAsync(&last_flag[k], f[k] + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    for (int k = 0; k < 3; ++k) {
        B2_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, f[k], f[k], (int)n, s));
        B2_CUDA(cudaMemcpyAsync(&last_pos[k], f[k] + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    g_b2_launches.fetch_add(3);
    B2_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) h_counts3[k] = (int64_t)last_pos[k] + last_flag[k];
    ctx->part_n = n;
    return 0;
}

int b2_exchange_scatter(b2_ctx *ctx, int64_t n, const double *z, double zlo, double zhi, int n_arrays,
                        const double *const *src, double *const *stay, double *const *left, double *const *right,
                        void *stream) {
    if (n <= 0) return 0;
    if (ctx->part_n != n || !ctx->scratch[1])
        return b2_fail(-4, "b2_exchange_scatter: call b2_exchange_classify first", __FILE__, __LINE__);
    if (n_arrays > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    const size_t na = ((size_t)(n + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    char *base = (char *)ctx->scratch[1];
    B2Part a;
    for (int k = 0; k < n_arrays; ++k) {
        a.src[k] = src[k]; a.stay[k] = stay[k];
        a.left[k] = left ? left[k] : nullptr; a.right[k] = right ? right[k] : nullptr;
    }
    k_partition_scatter<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(
        n, (const int32_t *)base, (const int32_t *)(base + na), (const int32_t *)(base + 2 * na), z, zlo, zhi, a, n_arrays);
    B2_LAUNCHED();
    return 0;
}

int b2_add_scalar(b2_ctx *ctx, int64_t n, double *v, double value, void *stream) {
    if (n <= 0) return 0;
    k_add_scalar<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, v, value);
    B2_LAUNCHED();
    return 0;
}

int b2_shift_periodic(b2_ctx *ctx, int64_t n, double *z, double zmin, double zmax, void *stream) {
    if (n <= 0) return 0;
    B2Prof prof_(B2P_PUSH, b2_stream_of(ctx, stream));
    k_shift_periodic<<<grid1d(n, 256), 256, 0, b2_stream_of(ctx, stream)>>>(n, z, zmin, zmax);
    B2_LAUNCHED();
    return 0;
}

}  // extern "C"

// b2_common.cuh -- internal helpers shared by the kernels of libfbpic_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <atomic>
#include "../../include/fbpic_b200.h"

#define B2_C_LIGHT 299792458.0
#define B2_FFT_MAX_LANES 8

struct b2_ctx {
    int device;
    cudaStream_t stream;
    // grow-only scratch (sort temp storage, permutation indices, ...)
    void *scratch[4];
    size_t scratch_bytes[4];
    // cuFFT plans keyed by (Nz, Nr, lane): one plan (and work area) per concurrent lane
    std::map<uint64_t, cufftHandle> fft_plans;
    // fork/join lanes of b2_fft_z_multi: independent latency-bound transforms overlap
    cudaStream_t fft_lane[B2_FFT_MAX_LANES];
    cudaEvent_t fft_fork, fft_join[B2_FFT_MAX_LANES];
    int fft_lanes;             // 0 = not initialised yet
    // result of the last b2_sort_cells on this context (lives in scratch slot 0)
    const int32_t *last_idx32;
    const int32_t *last_keys_sorted;
    int64_t last_sort_n;
    const void *last_sort_prefix;   // prefix_sum array of that sort: identifies the species it belongs to
    int64_t part_n;            // size of the last b2_exchange_classify (scratch slot 1)
    void *nccl_comm;
    int nccl_rank, nccl_size;
    int sm_count;
};

extern std::atomic<uint64_t> g_b2_launches;
extern thread_local char g_b2_err[512];

int b2_fail(int code, const char *what, const char *file, int line);

#define B2_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) return b2_fail((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define B2_LAUNCHED()                                                              \
    do {                                                                           \
        g_b2_launches.fetch_add(1, std::memory_order_relaxed);                     \
        cudaError_t _e = cudaPeekAtLastError();                                    \
        if (_e != cudaSuccess) return b2_fail((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static inline cudaStream_t b2_stream_of(b2_ctx *ctx, void *stream) {
    return stream ? (cudaStream_t)stream : (ctx ? ctx->stream : (cudaStream_t)0);
}

int b2_scratch(b2_ctx *ctx, int slot, size_t nbytes, void **ptr);
// b2_fft.cu: two-pass z-FFT; rc 1 = no plan for this length (cuFFT takes it)
int b2_fft_own(b2_ctx *ctx, int na, const void *const *in, void *const *out, int Nz, int Nr, int inverse,
               cudaStream_t s);
// b2_dht_tma.cu: drop the cached packed Hankel matrix / tensor maps of a buffer that is freed or overwritten
void b2_dht_forget(const void *p);

// ---- optional per-kernel-family device timing (CUDA events on the launching stream) ----
enum B2ProfSlot {
    B2P_CELL_INDEX = 0, B2P_SORT, B2P_PERMUTE, B2P_GATHER, B2P_PUSH, B2P_GATHER_PUSH, B2P_DEPOSIT_RHO,
    B2P_DEPOSIT_J, B2P_FFT, B2P_DHT, B2P_SPECTRAL, B2P_ELEMENTWISE, B2P_COMM, B2P_NSLOTS
};
extern bool g_b2_prof_on;
struct B2Prof {
    int idx;
    cudaStream_t s;
    B2Prof(int slot, cudaStream_t stream);
    ~B2Prof();
};

// pointer bundle passed by value to multi-array kernels
struct B2Ptrs {
    void *p[B2_MAX_ARRAYS];
};

// ---- device helpers -------------------------------------------------------------------
// (r_cell, z_cell) in cell units and cos/sin of the azimuth: the same arithmetic for the cell
// key, the gather and the deposition (SURVEY Appendix A notation).
struct B2Cyl {
    double r, cs, sn, r_cell, z_cell;
};
__device__ __forceinline__ B2Cyl b2_cyl(double xj, double yj, double zj,
                                        double invdz, double zmin, double invdr, double rmin) {
    B2Cyl c;
    // explicit round-to-nearest mul/add (never FMA-contracted): the cell key derived from
    // r_cell / z_cell must be bit-identical to the CPU formula (cuda_sorting.py:64-72).
    c.r = sqrt(__dadd_rn(__dmul_rn(xj, xj), __dmul_rn(yj, yj)));
    if (c.r != 0.) {
        double invr = 1. / c.r;
        c.cs = xj * invr;
        c.sn = yj * invr;
    } else {
        c.cs = 1.;
        c.sn = 0.;
    }
    c.r_cell = __dadd_rn(__dmul_rn(invdr, __dsub_rn(c.r, rmin)), -0.5);
    c.z_cell = __dadd_rn(__dmul_rn(invdz, __dsub_rn(zj, zmin)), -0.5);
    return c;
}

// cell key (cuda_sorting.py:55-88)
__device__ __forceinline__ int b2_cell_of(const B2Cyl &c, int Nz, int Nr) {
    int ir_upper = (int)ceil(c.r_cell);
    int iz_upper = (int)ceil(c.z_cell);
    if (ir_upper > Nr) ir_upper = Nr;
    if (iz_upper < 0) iz_upper += Nz;
    else if (iz_upper > Nz - 1) iz_upper -= Nz;
    return ir_upper + iz_upper * (Nr + 1);
}


// Vay pusher (push/inline_functions.py:11-48)
__device__ __forceinline__ void b2_vay(double &ux, double &uy, double &uz, double &inv_gamma,
                                       const double F[6], double econst, double bconst) {
    const double taux = bconst * F[3], tauy = bconst * F[4], tauz = bconst * F[5];
    const double tau2 = taux * taux + tauy * tauy + tauz * tauz;
    const double uxp = ux + econst * F[0] + inv_gamma * (uy * tauz - uz * tauy);
    const double uyp = uy + econst * F[1] + inv_gamma * (uz * taux - ux * tauz);
    const double uzp = uz + econst * F[2] + inv_gamma * (ux * tauy - uy * taux);
    const double sigma = 1 + uxp * uxp + uyp * uyp + uzp * uzp - tau2;
    const double utau = uxp * taux + uyp * tauy + uzp * tauz;
    const double igf = sqrt(2. / (sigma + sqrt(sigma * sigma + 4 * (tau2 + utau * utau))));
    const double tx = igf * taux, ty = igf * tauy, tz = igf * tauz, ut = igf * utau;
    const double s = 1. / (1 + tau2 * igf * igf);
    ux = s * (uxp + tx * ut + uyp * tz - uzp * ty);
    uy = s * (uyp + ty * ut + uzp * tx - uxp * tz);
    uz = s * (uzp + tz * ut + uxp * ty - uyp * tx);
    inv_gamma = igf;
}


// stencil sum of one particle from a shared-memory tile of all 6*NM mode arrays (row pitch PITCH cells); AXIS: with
// the two guard-cell terms of a particle within half a cell of the axis (gathering/cuda_methods.py:126-160).
// F = sum_m fac_m Re[(sum_pt S_pt F_m[pt]) e^{-i m theta}] is evaluated with the phase folded into the weights,
// w_re = fac_m S_pt Re(e), w_im = -fac_m S_pt Im(e):  F += sum_pt (w_re Re F_m[pt] + w_im Im F_m[pt]) -- 8 fused
// multiply-adds per array instead of 8 + a 3-operation combine, and for m = 0 (e = 1) only the real parts are
// read (4 multiply-adds, 8-byte loads).
template <int NM, bool AXIS, int PITCH>
__device__ __forceinline__ void b2_tile_sum(const double2 (*tile)[PITCH], int t_ll, int t_lu, int t_ul, int t_uu,
                                            int t_l0, int t_u0, double S_ll, double S_lu, double S_ul, double S_uu,
                                            double S_lg, double S_ug, bool on_axis, double cs, double sn,
                                            double (&Fc)[2][3]) {
    double e_re = 1., e_im = 0.;
    if (NM > 2) {
        // (three and four modes: the 8 folded weights per mode cost more registers than they save operations --
        //  measured at C4, 4.04 ms plain against 4.22 ms folded -- so the sum is combined after the fact)
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const double flip = (m & 1) ? -1. : 1.;
            const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
            for (int f = 0; f < 2; ++f) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int a = 6 * m + 3 * f + k;
                    const double2 v_ll = tile[a][t_ll], v_lu = tile[a][t_lu], v_ul = tile[a][t_ul], v_uu = tile[a][t_uu];
                    double re = 0., im = 0.;
                    re += S_ll * v_ll.x; im += S_ll * v_ll.y;
                    re += S_lu * v_lu.x; im += S_lu * v_lu.y;
                    re += S_ul * v_ul.x; im += S_ul * v_ul.y;
                    re += S_uu * v_uu.x; im += S_uu * v_uu.y;
                    if (AXIS && on_axis) {
                        const double sgn = (k == 2) ? flip : -flip;
                        const double2 v_l0 = tile[a][t_l0], v_u0 = tile[a][t_u0];
                        re += sgn * S_lg * v_l0.x; im += sgn * S_lg * v_l0.y;
                        re += sgn * S_ug * v_u0.x; im += sgn * S_ug * v_u0.y;
                    }
                    Fc[f][k] += factor * (re * e_re - im * e_im);
                }
            }
            const double nr = e_re * cs + e_im * sn, ni = e_im * cs - e_re * sn;
            e_re = nr; e_im = ni;
        }
        return;
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const double flip = (m & 1) ? -1. : 1.;
        const double factor = (m == 0) ? 1. : 2.;
        const double a_re = factor * e_re, a_im = -factor * e_im;
        const double r_ll = S_ll * a_re, r_lu = S_lu * a_re, r_ul = S_ul * a_re, r_uu = S_uu * a_re;
        const double i_ll = S_ll * a_im, i_lu = S_lu * a_im, i_ul = S_ul * a_im, i_uu = S_uu * a_im;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int a = 6 * m + 3 * f + k;
                double acc;
                if (m == 0) {
                    acc = r_ll * tile[a][t_ll].x;
                    acc += r_lu * tile[a][t_lu].x;
                    acc += r_ul * tile[a][t_ul].x;
                    acc += r_uu * tile[a][t_uu].x;
                } else {
                    const double2 v_ll = tile[a][t_ll], v_lu = tile[a][t_lu], v_ul = tile[a][t_ul], v_uu = tile[a][t_uu];
                    acc = r_ll * v_ll.x;  acc += i_ll * v_ll.y;
                    acc += r_lu * v_lu.x; acc += i_lu * v_lu.y;
                    acc += r_ul * v_ul.x; acc += i_ul * v_ul.y;
                    acc += r_uu * v_uu.x; acc += i_uu * v_uu.y;
                }
                if (AXIS && on_axis) {
                    const double sgn = (k == 2) ? flip : -flip;
                    const double2 v_l0 = tile[a][t_l0], v_u0 = tile[a][t_u0];
                    acc += sgn * (S_lg * (a_re * v_l0.x + a_im * v_l0.y) + S_ug * (a_re * v_u0.x + a_im * v_u0.y));
                }
                Fc[f][k] += acc;
            }
        }
        const double nr = e_re * cs + e_im * sn, ni = e_im * cs - e_re * sn;
        e_re = nr; e_im = ni;
    }
}

// b2_gather_pipe.cu: persistent gather + push kernel with TMA-staged field tiles (linear shapes).  Processes the
// first *done particles (a multiple of 128, possibly 0 when the path is unavailable); rc != 0 is an error.
int b2_gather_push_pipe(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, double *ux, double *uy, double *uz,
                        double *inv_gamma, double rmax_gather, double invdz, double zmin, int Nz, double invdr,
                        double rmin, int Nr, int Nm, const void *const *grids, double econst, double bconst, double chdt,
                        int32_t *cell_idx, double key_zmin, cudaStream_t s, int64_t *done);

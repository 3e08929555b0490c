// b2_gather_pipe.cu -- fused field gather + Vay push + half position push, linear shapes, as a persistent,
// software-pipelined kernel with TMA-staged field tiles.
//
// Replaces (like k_gather_push_tiled of b2_particles.cu) gather_field_gpu_linear + push_p_gpu + push_x_gpu
// (fbpic/particles/gathering/cuda_methods.py:26-205, push/cuda_methods.py:17-100): one read and one write of the
// particle state per step, gathered fields never leave the registers.  What changes is HOW the operands arrive:
// k_gather_push_tiled runs three dependent DRAM round trips per 128-particle CTA (positions -> bounding box ->
// field tile -> momenta) and hides them only by occupancy (55 % issue utilisation measured).  Here a CTA is
// persistent, walks chunks of 128 cell-sorted particles, and keeps three chunks in flight:
//   chunk c+2 : its 7 SoA slices (x,y,z,ux,uy,uz,inv_gamma; 7 x 1 KB) stream into shared memory by cp.async.bulk;
//   chunk c+1 : positions are in shared memory -> cylindrical coordinates, stencil bounding box of the chunk
//               (warp min/max + 4 shared atomics per warp), then ONE thread issues 6*Nm cp.async.bulk.tensor.2d
//               boxes (3 rows x 16 cells of every E/B mode array, the TMA-staged field tile);
//   chunk c   : tile and particle data are resident -> gather from shared memory, push, coalesced stores.
// All global->shared traffic is asynchronous (TMA unit, mbarrier completion); the threads only compute.
// A chunk whose bounding box does not fit the 3 x 16 box (unsorted particles, a z-row boundary of the sorted
// order, the periodic seam) gathers from global memory instead -- per particle, results are identical either way.
#include "b2_common.cuh"
#include <cuda.h>
#include <climits>
#include <cstring>
#include <mutex>

#define GQ_TPB 128
#define GQ_ROWS 3
#define GQ_COLS 16
#define GQ_PSLOTS 3                       // particle-data ring
#define GQ_TSLOTS 2                       // field-tile ring
#define GQ_NPA 12                         // per-particle doubles in a ring slot: 7 loaded + r_cell, z_cell, cs, sn, r

int b2_tma_field_map(const void *A, int Nz, int Nr, int box_d, int box_rows, int swizzle128, CUtensorMap *out);
std::mutex &b2_tma_mutex();

template <int NM> struct GqParams {
    CUtensorMap map[6 * NM];              // [m][Er,Et,Ez,Br,Bt,Bz]
    const double2 *g[6 * NM];
    double *x, *y, *z, *ux, *uy, *uz, *ig;
    int32_t *cell_idx;
    long long nchunks;
    double rmax_gather, invdz, zmin, invdr, rmin, econst, bconst, chdt, key_zmin;
    int Nz, Nr;
};

__device__ __forceinline__ uint32_t gq_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gq_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gq_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void gq_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gq_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gq_tma_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void gq_bulk_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared memory of one CTA
template <int NM> struct GqSmem {
    double2 tile[GQ_TSLOTS][6 * NM][GQ_ROWS * GQ_COLS];      // TMA destination: 128-byte aligned (768 B per array)
    double part[GQ_PSLOTS][GQ_NPA][GQ_TPB];
    unsigned long long pfull[GQ_PSLOTS], tfull[GQ_TSLOTS];
    int box[GQ_TSLOTS][4];                                     // min iz_l (unwrapped), max iz_u, min ir, max ir
    int anchor[GQ_TSLOTS][2];
};

template <int NM>
__global__ void __launch_bounds__(GQ_TPB, 4)
k_gather_push_pipe(const __grid_constant__ GqParams<NM> P) {
    extern __shared__ __align__(128) unsigned char gq_raw[];
    GqSmem<NM> &S = *reinterpret_cast<GqSmem<NM> *>(gq_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const long long G = gridDim.x;
    const long long c0 = blockIdx.x;
    const long long my_chunks = (P.nchunks - c0 + G - 1) / G;          // chunks c0, c0+G, ...

    if (tid == 0) {
        for (int s = 0; s < GQ_PSLOTS; ++s) gq_mbar_init(gq_smem_u32(&S.pfull[s]), 1);
        for (int s = 0; s < GQ_TSLOTS; ++s) gq_mbar_init(gq_smem_u32(&S.tfull[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // particle slices of local chunk k -> ring slot k % 3 (one thread, 7 bulk copies of 1 KB)
    auto issue_particles = [&](long long k) {
        const int slot = (int)(k % GQ_PSLOTS);
        const size_t off = (size_t)(c0 + k * G) * GQ_TPB;
        const uint32_t bar = gq_smem_u32(&S.pfull[slot]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        gq_mbar_expect_tx(bar, 7 * GQ_TPB * 8);
        const double *src[7] = {P.x, P.y, P.z, P.ux, P.uy, P.uz, P.ig};
#pragma unroll
        for (int a = 0; a < 7; ++a) gq_bulk_1d(gq_smem_u32(&S.part[slot][a][0]), src[a] + off, GQ_TPB * 8, bar);
    };

    // stage 2 of local chunk k: cylindrical coordinates -> ring slot, bounding box -> tile request
    auto prepare_tile = [&](long long k) {
        const int ps = (int)(k % GQ_PSLOTS), ts = (int)(k % GQ_TSLOTS);
        gq_mbar_wait(gq_smem_u32(&S.pfull[ps]), (uint32_t)((k / GQ_PSLOTS) & 1));
        const double xj = S.part[ps][0][tid], yj = S.part[ps][1][tid], zj = S.part[ps][2][tid];
        const B2Cyl c = b2_cyl(xj, yj, zj, P.invdz, P.zmin, P.invdr, P.rmin);
        S.part[ps][7][tid] = c.r_cell; S.part[ps][8][tid] = c.z_cell;
        S.part[ps][9][tid] = c.cs;     S.part[ps][10][tid] = c.sn;   S.part[ps][11][tid] = c.r;
        int ir_l = (int)floor(c.r_cell), ir_u = ir_l + 1;
        const int iz_l0 = (int)floor(c.z_cell);
        if (ir_l < 0) ir_l = 0;
        if (ir_l > P.Nr - 1) ir_l = P.Nr - 1;
        if (ir_u > P.Nr - 1) ir_u = P.Nr - 1;
        if (tid == 0) {
            S.box[ts][0] = INT_MAX; S.box[ts][1] = INT_MIN; S.box[ts][2] = INT_MAX; S.box[ts][3] = INT_MIN;
            S.anchor[ts][0] = iz_l0; S.anchor[ts][1] = ir_l;
        }
        __syncthreads();
        const bool active = c.r < P.rmax_gather;
        // window around the chunk's first particle that always fits the box: rows az-1 .. az+1, 16 columns from
        // ar-2 (sorted particles sit at or right of the first one); anything else gathers from global memory
        const int az = S.anchor[ts][0], ar = S.anchor[ts][1];
        const bool near = active && (iz_l0 == az || iz_l0 == az - 1) && ir_l >= ar - 2 && ir_u <= ar + (GQ_COLS - 3);
        int v0 = near ? iz_l0 : INT_MAX, v1 = near ? iz_l0 + 1 : INT_MIN;
        int v2 = near ? ir_l : INT_MAX, v3 = near ? ir_u : INT_MIN;
        v0 = __reduce_min_sync(0xffffffffu, v0); v1 = __reduce_max_sync(0xffffffffu, v1);
        v2 = __reduce_min_sync(0xffffffffu, v2); v3 = __reduce_max_sync(0xffffffffu, v3);
        if (lane == 0) {
            atomicMin(&S.box[ts][0], v0); atomicMax(&S.box[ts][1], v1);
            atomicMin(&S.box[ts][2], v2); atomicMax(&S.box[ts][3], v3);
        }
        __syncthreads();
        if (tid == 0) {
            const int z0 = S.box[ts][0], z1 = S.box[ts][1], r0 = S.box[ts][2], r1 = S.box[ts][3];
            // rows beyond Nz-1 / columns beyond Nr-1 of the box are zero-filled by the TMA unit and never read;
            // a box that crosses the periodic seam in z is not staged
            const bool ok = (z1 >= z0) && (r1 >= r0) && (z1 - z0 < GQ_ROWS) && (r1 - r0 < GQ_COLS)
                            && z0 >= 0 && z1 <= P.Nz - 1;
            const uint32_t bar = gq_smem_u32(&S.tfull[ts]);
            if (ok) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                gq_mbar_expect_tx(bar, 6 * NM * GQ_ROWS * GQ_COLS * 16);
#pragma unroll
                for (int a = 0; a < 6 * NM; ++a)
                    gq_tma_2d(gq_smem_u32(&S.tile[ts][a][0]), &P.map[a], bar, 2 * r0, z0);
            } else {
                S.box[ts][1] = INT_MIN;          // marks "no tile" for the consumers (box[1] < box[0])
                gq_mbar_arrive(bar);
            }
        }
    };

    if (my_chunks > 0 && tid == 0) {
        issue_particles(0);
        if (my_chunks > 1) issue_particles(1);
    }
    if (my_chunks > 0) prepare_tile(0);

    for (long long k = 0; k < my_chunks; ++k) {
        const int ps = (int)(k % GQ_PSLOTS), ts = (int)(k % GQ_TSLOTS);
        // slot (k+2)%3 held chunk k-1, whose consumers passed the barrier at the end of the previous iteration
        if (tid == 0 && k + 2 < my_chunks) issue_particles(k + 2);
        if (k + 1 < my_chunks) prepare_tile(k + 1);
        __syncthreads();                                   // tile request of k+1 posted; box of chunk k is final
        gq_mbar_wait(gq_smem_u32(&S.tfull[ts]), (uint32_t)((k / GQ_TSLOTS) & 1));

        // ---------------------------------------------------------------- chunk k: gather + push
        const long long i = (c0 + k * G) * GQ_TPB + tid;
        const double xj = S.part[ps][0][tid], yj = S.part[ps][1][tid], zj = S.part[ps][2][tid];
        const double r_cell = S.part[ps][7][tid], z_cell = S.part[ps][8][tid];
        const double cs = S.part[ps][9][tid], sn = S.part[ps][10][tid];
        const int z0 = S.box[ts][0], r0 = S.box[ts][2];
        const bool tile_ok = S.box[ts][1] >= z0;
        // stencil indices and weights exactly as gather_field_gpu_linear (gathering/cuda_methods.py:109-160)
        int ir_l = (int)floor(r_cell), ir_u = ir_l + 1;
        const int iz_l0 = (int)floor(z_cell);
        double Sr_l = ir_u - r_cell, Sr_u = r_cell - ir_l;
        const double Sz_l = (iz_l0 + 1) - z_cell, Sz_u = z_cell - iz_l0;
        double Sr_g = 0.;
        if (ir_l < 0) { Sr_g = Sr_l; Sr_l = 0.; ir_l = 0; }
        if (ir_l > P.Nr - 1) ir_l = P.Nr - 1;
        if (ir_u > P.Nr - 1) ir_u = P.Nr - 1;
        const bool active = S.part[ps][11][tid] < P.rmax_gather;
        const bool use_tile = tile_ok && active && iz_l0 >= z0 && iz_l0 + 1 < z0 + GQ_ROWS && ir_l >= r0
                              && ir_u < r0 + GQ_COLS;
        double Fc[2][3] = {{0., 0., 0.}, {0., 0., 0.}};
        // tile-only copy of the stencil sum when every gathering lane of the warp reads the tile (no dead
        // predicated LDG and address arithmetic), mixed copy otherwise
        const bool warp_on_tile = __all_sync(0xffffffffu, use_tile || !active);
        // (guard-cell terms near the axis: skipped as a whole by the warps that hold no such particle)
        const bool warp_near_axis = __any_sync(0xffffffffu, active && ir_l == 0 && ir_u == 0);
        if (active) {
            const double S_ll = Sz_l * Sr_l, S_lu = Sz_l * Sr_u, S_ul = Sz_u * Sr_l, S_uu = Sz_u * Sr_u;
            const double S_lg = Sz_l * Sr_g, S_ug = Sz_u * Sr_g;
            const bool on_axis = (ir_l == 0 && ir_u == 0);
            const int t_ll = (iz_l0 - z0) * GQ_COLS + (ir_l - r0), t_lu = (iz_l0 - z0) * GQ_COLS + (ir_u - r0);
            const int t_ul = t_ll + GQ_COLS, t_uu = t_lu + GQ_COLS;
            const int t_l0 = (iz_l0 - z0) * GQ_COLS - r0, t_u0 = t_l0 + GQ_COLS;   // column 0 (on_axis: r0 == 0)
            double e_re = 1., e_im = 0.;
            if (warp_on_tile) {
                if (warp_near_axis)
                    b2_tile_sum<NM, true, GQ_ROWS * GQ_COLS>(S.tile[ts], t_ll, t_lu, t_ul, t_uu, t_l0, t_u0, S_ll, S_lu,
                                                              S_ul, S_uu, S_lg, S_ug, on_axis, cs, sn, Fc);
                else
                    b2_tile_sum<NM, false, GQ_ROWS * GQ_COLS>(S.tile[ts], t_ll, t_lu, t_ul, t_uu, t_l0, t_u0, S_ll, S_lu,
                                                               S_ul, S_uu, S_lg, S_ug, false, cs, sn, Fc);
            } else {
                int iz_l = iz_l0, iz_u = iz_l0 + 1;
                if (iz_l < 0) iz_l += P.Nz;
                if (iz_u < 0) iz_u += P.Nz;
                if (iz_l > P.Nz - 1) iz_l -= P.Nz;
                if (iz_u > P.Nz - 1) iz_u -= P.Nz;
                const size_t o_ll = (size_t)iz_l * P.Nr + ir_l, o_lu = (size_t)iz_l * P.Nr + ir_u;
                const size_t o_ul = (size_t)iz_u * P.Nr + ir_l, o_uu = (size_t)iz_u * P.Nr + ir_u;
                const size_t o_l0 = (size_t)iz_l * P.Nr, o_u0 = (size_t)iz_u * P.Nr;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    const double flip = (m & 1) ? -1. : 1.;
                    const double factor = (m == 0) ? 1. : 2.;
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const int a = 6 * m + 3 * f + q;
                            double2 v_ll, v_lu, v_ul, v_uu;
                            if (use_tile) {
                                const double2 *T = S.tile[ts][a];
                                v_ll = T[t_ll]; v_lu = T[t_lu]; v_ul = T[t_ul]; v_uu = T[t_uu];
                            } else {
                                const double2 *g = P.g[a];
                                v_ll = __ldg(g + o_ll); v_lu = __ldg(g + o_lu); v_ul = __ldg(g + o_ul); v_uu = __ldg(g + o_uu);
                            }
                            double re = 0., im = 0.;
                            re += S_ll * v_ll.x; im += S_ll * v_ll.y;
                            re += S_lu * v_lu.x; im += S_lu * v_lu.y;
                            re += S_ul * v_ul.x; im += S_ul * v_ul.y;
                            re += S_uu * v_uu.x; im += S_uu * v_uu.y;
                            if (on_axis) {
                                const double sgn = (q == 2) ? flip : -flip;
                                double2 v_l0, v_u0;
                                if (use_tile) { v_l0 = S.tile[ts][a][t_l0]; v_u0 = S.tile[ts][a][t_u0]; }
                                else { v_l0 = __ldg(P.g[a] + o_l0); v_u0 = __ldg(P.g[a] + o_u0); }
                                re += sgn * S_lg * v_l0.x; im += sgn * S_lg * v_l0.y;
                                re += sgn * S_ug * v_u0.x; im += sgn * S_ug * v_u0.y;
                            }
                            Fc[f][q] += factor * (re * e_re - im * e_im);
                        }
                    }
                    const double nr = e_re * cs + e_im * sn, ni = e_im * cs - e_re * sn;
                    e_re = nr; e_im = ni;
                }
            }
        }
        double F[6];
        F[0] = cs * Fc[0][0] - sn * Fc[0][1];
        F[1] = sn * Fc[0][0] + cs * Fc[0][1];
        F[2] = Fc[0][2];
        F[3] = cs * Fc[1][0] - sn * Fc[1][1];
        F[4] = sn * Fc[1][0] + cs * Fc[1][1];
        F[5] = Fc[1][2];
        double a = S.part[ps][3][tid], b = S.part[ps][4][tid], cz = S.part[ps][5][tid], gam = S.part[ps][6][tid];
        b2_vay(a, b, cz, gam, F, P.econst, P.bconst);
        P.ux[i] = a; P.uy[i] = b; P.uz[i] = cz; P.ig[i] = gam;
        const double xn = xj + P.chdt * gam * 1. * a;
        const double yn = yj + P.chdt * gam * 1. * b;
        const double zn = zj + P.chdt * gam * 1. * cz;
        P.x[i] = xn; P.y[i] = yn; P.z[i] = zn;
        if (P.cell_idx)
            P.cell_idx[i] = b2_cell_of(b2_cyl(xn, yn, zn, P.invdz, P.key_zmin, P.invdr, P.rmin), P.Nz, P.Nr);
        __syncthreads();       // chunk k is consumed: its particle slot and tile slot may be refilled
    }
}

template <int NM>
static int gq_launch(b2_ctx *ctx, int64_t nchunks, double *x, double *y, double *z, double *ux, double *uy, double *uz,
                     double *ig, double rmax_gather, double invdz, double zmin, int Nz, double invdr, double rmin,
                     int Nr, const void *const *grids, double ec, double bc, double chdt, int32_t *cell_idx,
                     double key_zmin, cudaStream_t s) {
    GqParams<NM> P;
    memset(&P, 0, sizeof(P));
    {
        std::lock_guard<std::mutex> lk(b2_tma_mutex());
        for (int a = 0; a < 6 * NM; ++a) {
            int rc = b2_tma_field_map(grids[a], Nz, Nr, 2 * GQ_COLS, GQ_ROWS, 0, &P.map[a]);
            if (rc) return rc;                   // 1: TMA descriptors unavailable
            P.g[a] = (const double2 *)grids[a];
        }
    }
    P.x = x; P.y = y; P.z = z; P.ux = ux; P.uy = uy; P.uz = uz; P.ig = ig; P.cell_idx = cell_idx;
    P.nchunks = nchunks;
    P.rmax_gather = rmax_gather; P.invdz = invdz; P.zmin = zmin; P.invdr = invdr; P.rmin = rmin;
    P.econst = ec; P.bconst = bc; P.chdt = chdt; P.key_zmin = key_zmin; P.Nz = Nz; P.Nr = Nr;
    const int smem = (int)sizeof(GqSmem<NM>);
    static bool attr_set = false;
    if (!attr_set) {
        B2_CUDA(cudaFuncSetAttribute(k_gather_push_pipe<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    int per_sm = (227 * 1024) / (smem + 1024);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)(ctx && ctx->sm_count > 0 ? ctx->sm_count : 148) * per_sm;
    if (grid > nchunks) grid = nchunks;
    k_gather_push_pipe<NM><<<(unsigned)grid, GQ_TPB, smem, s>>>(P);
    B2_LAUNCHED();
    return 0;
}

int b2_gather_push_pipe(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, double *ux, double *uy, double *uz,
                        double *inv_gamma, double rmax_gather, double invdz, double zmin, int Nz, double invdr,
                        double rmin, int Nr, int Nm, const void *const *grids, double econst, double bconst, double chdt,
                        int32_t *cell_idx, double key_zmin, cudaStream_t s, int64_t *done) {
    *done = 0;
    // opt-in: B2_GATHER_IMPL=pipe (every Nm).  Measured on a B200 (profiles/r02_gather_pipe_vs_tiled.md): at C2
    // 0.620 ms here against 0.596 ms of k_gather_push_tiled, at C4 (Nm = 4, 36 KB of tile per slot: 3 CTAs per SM)
    // 4.5 against 4.0 ms -- the tiled kernel is the default.
    static const bool on = []() { const char *e = getenv("B2_GATHER_IMPL"); return e && !strcmp(e, "pipe"); }();
    const bool off = !on;
    const int64_t nchunks = n / GQ_TPB;
    if (off || nchunks == 0 || Nz < GQ_ROWS || Nr < GQ_COLS) return 0;
    // cp.async.bulk needs 16-byte aligned sources: chunk starts are multiples of 1 KB from the array base
    const uintptr_t al = (uintptr_t)x | (uintptr_t)y | (uintptr_t)z | (uintptr_t)ux | (uintptr_t)uy | (uintptr_t)uz
                         | (uintptr_t)inv_gamma;
    if (al & 15) return 0;
    int rc;
#define GQ_ARGS ctx, nchunks, x, y, z, ux, uy, uz, inv_gamma, rmax_gather, invdz, zmin, Nz, invdr, rmin, Nr, grids, econst, bconst, chdt, cell_idx, key_zmin, s
    switch (Nm) {
        case 1: rc = gq_launch<1>(GQ_ARGS); break;
        case 2: rc = gq_launch<2>(GQ_ARGS); break;
        case 3: rc = gq_launch<3>(GQ_ARGS); break;
        default: rc = gq_launch<4>(GQ_ARGS); break;
    }
#undef GQ_ARGS
    if (rc == 1) return 0;              // descriptor API unavailable: the caller's kernels take everything
    if (rc) return rc;
    *done = nchunks * GQ_TPB;
    return 0;
}

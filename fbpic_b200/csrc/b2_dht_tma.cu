// b2_dht_tma.cu -- discrete Hankel transform: TMA-fed, persistent, warp-specialised fp64 DMMA GEMM.
//
// Same contract as k_dht (b2_dht.cu; replaces DHT.transform / inverse_transform, hankel.py:182-243, and the
// (r,t)<->(p,m) passes spectral_transform/cuda_methods.py:120-158):
//     out[iz, n] = rowscale[iz] * sum_j A[iz, j] * M[j, n]          (complex A, real M)
// but the operands reach shared memory through the TMA unit instead of LDG -> registers -> STS:
//   * A (field array, complex [Nz,Nr] as it lies in HBM) : cp.async.bulk.tensor.2d boxes of 64 rows x 8 complex
//     (128-byte rows, SWIZZLE_128B); the DMMA row g of an 8-row block is fed from tile row rho(g) =
//     ((g&1)<<2)|(g>>1), which makes the 8 lanes of every LDS.128 quarter-warp hit 8 different 16-byte chunks
//     of the swizzled layout (conflict-free without padding);
//   * M (Hankel matrix)  : re-tiled once on the device into the fragment order [k/4][n/4][k%4][n%4]
//     (b2_dht_pack, cached per matrix) so that a K-chunk of a column strip is one contiguous cp.async.bulk and
//     every half-warp of an LDS.64 reads 128 contiguous bytes.
// One producer warp runs a STAGES-deep mbarrier ring (full/empty) ahead of 8 consumer warps (2 x 4, warp tile
// 32 rows x 32 columns x re/im = 64 fp64 accumulators per thread: 8 shared-memory loads feed 32 DMMAs);
// consumers never meet at a CTA barrier.  Persistent grid (one CTA per SM): the (job, column strip, 16-row unit) space is cut
// into equal contiguous shares, so the last wave is as full as the first (the tile grid of k_dht lost 13 % to
// wave quantisation at 6 jobs x 4096 x 256).  Tiles are up to 64 rows x 128 columns (scalar jobs, one product) or
// 64 rows x 64 columns of two products (vector jobs: forward (r,t)->(p,m) with the combination applied to the A
// fragments, inverse (p,m)->(r,t) with the combination in the epilogue); ONE launch carries a mixed list of all
// three job kinds (the tile kind switches per tile), so a whole batch of a PIC step pays one ramp-up and one tail.
// fp64 only: sm_100a has no tcgen05 f64 kind; the f64 tensor path is mma.sync m8n8k4 (DMMA.8x8x4).
#include "b2_common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#define DT_BM 64                 // rows of iz per full tile
#define DT_BK 16                 // K per stage (2 TMA boxes of 8 complex per A operand)
#define DT_CONSUMER_WARPS 8
#define DT_THREADS ((DT_CONSUMER_WARPS + 4) * 32)   // + one producer warpgroup (register donor)
#define DT_MAX_JOBS 16

enum { DT_SCALAR = 0, DT_INVERSE = 1, DT_FORWARD = 2 };    // tile kinds: 1 product | 2 products, mix in the epilogue | on A
struct DtJob {
    CUtensorMap mapA[2];         // scalar: in1 ; vector: (r,t) or (p,m)
    const double *B[2];          // packed matrices (k_dht_pack layout)
    double2 *out[2];
    const double *rowscale;
    int kind;                    // DT_SCALAR / DT_INVERSE / DT_FORWARD
    int n_strips;                // column strips of this job: ceil(Nr / 128) (scalar) or ceil(Nr / 64) (vector)
    int unit_begin, pad_;        // first 16-row unit of this job in the launch-wide unit space
};
struct DtParams {
    DtJob job[DT_MAX_JOBS];
    int njobs, Nz, Nr, Np;       // Np: padded matrix width (multiple of 128)
    int KT;                      // K stages = ceil(Nr / 16)
    int blocks64;                // ceil(Nz / 64): 64-row blocks per column strip
    int total_units, producer_sleep;
    unsigned stagger_ns, pad3_;
    long long *dbg;              // -DDT_DEBUG builds: per consumer warp [t_total, t_kloop, t_epilogue, tiles] (cycles)
};

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool dt_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void dt_mbar_wait(uint32_t bar, uint32_t parity) {
    while (!dt_mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void dt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dt_tma_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dt_bulk_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory loads by 32-bit shared-window address (a pointer derived from the aligned-up base would be
// generic: LD.E through the address translation instead of LDS)
__device__ __forceinline__ double2 dt_lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double dt_lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dt_dmma(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define DT_STAGES 4
#define DT_STAGE_BYTES (48 * 1024)      // ring slot: fits the two-product stage (32 KB of A + 16 KB of M)
#define DT_SMEM_BYTES (DT_STAGES * DT_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/)
template <int NPROD> struct DtCfg {
    static constexpr int BN = 128 / NPROD;                       // output columns per product per tile
    static constexpr int A_BOX_BYTES = DT_BM * 128;              // 64 rows x 8 complex
    static constexpr int A_BYTES = NPROD * 2 * A_BOX_BYTES;      // per stage
    static constexpr int B_KB_BYTES = BN * 32;                   // one k4 block of the strip: [BN/4][4][4] doubles
    static constexpr int B_BYTES = NPROD * 4 * B_KB_BYTES;       // per stage
    static constexpr int TX_BYTES = A_BYTES + B_BYTES;           // bytes the TMA unit delivers per stage
    static constexpr int NI = BN / 4 / 8;                        // 8-column MMA blocks per warp per product
};

// the tile walk shared by the producer and the consumers: contiguous share of the 16-row units
struct DtWalk {
    int u, u_end, first;
    __device__ __forceinline__ DtWalk(const DtParams &P) {
        const long long G = gridDim.x, b = blockIdx.x;
        u = (int)((long long)P.total_units * b / G);
        u_end = (int)((long long)P.total_units * (b + 1) / G);
        // the first tile of CTA b has 1 + b%4 units: the CTAs reach their epilogues (a 128 KB burst of stores
        // each) at four different phases of the tile period instead of all at once
        first = (u_end - u >= 8) ? 1 + (int)(b & 3) : DT_BM / 16;
    }
    // next tile: job, column strip, first row, number of 16-row units (1..4); false when done.
    // Unit order inside a job: (64-row block, column strip, 16-row unit of the block) -- the strips of one row
    // block follow each other in the walk, so the A rows a CTA has just streamed for strip s are still in L2
    // (mostly in flight on the same SM) when strip s+1 asks for them: DRAM reads every field array once
    // (strip-major order re-read it once per strip: 2.1x the algorithmic bytes in the r02 capture).
    __device__ __forceinline__ bool next(const DtParams &P, int &job, int &strip, int &m0, int &cnt) {
        while (u < u_end) {
            int j = 0;
            while (j + 1 < P.njobs && P.job[j + 1].unit_begin <= u) ++j;
            const int v = u - P.job[j].unit_begin;
            const int sub = v & 3, q = v >> 2;
            const int ns = P.job[j].n_strips;
            const int mb = q / ns;
            int c = DT_BM / 16 - sub;
            if (c > first) c = first;
            first = DT_BM / 16;
            if (c > u_end - u) c = u_end - u;
            u += c;
            const int row0 = mb * DT_BM + sub * 16;
            if (row0 >= P.Nz) continue;               // padding units of the last row block
            job = j; strip = q - mb * ns; m0 = row0; cnt = c;
            return true;
        }
        return false;
    }
};

// fragments of one k4 step (4 values of K) of a consumer warp: NB 8-row blocks of A (re, im), NI column blocks of M
template <int NPROD, int NB> struct DtFrag {
    double2 a[NPROD][NB > 0 ? NB : 1];
    double b[NPROD][DtCfg<NPROD>::NI];
};
template <int NPROD, bool MIX, int NB>
__device__ __forceinline__ void dt_load_frag(DtFrag<NPROD, NB> &f, uint32_t sa, uint32_t sb, int k4, uint32_t a_off0,
                                             uint32_t a_off1) {
    using C = DtCfg<NPROD>;
#pragma unroll
    for (int p = 0; p < NPROD; ++p)
#pragma unroll
        for (int n = 0; n < C::NI; ++n) f.b[p][n] = dt_lds64(sb + (p * 4 + k4) * C::B_KB_BYTES + n * 256);
#pragma unroll
    for (int i = 0; i < NB; ++i) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p)
            f.a[p][i] = dt_lds128(sa + (p * 2 + (k4 >> 1)) * C::A_BOX_BYTES + i * 1024 + ((k4 & 1) ? a_off1 : a_off0));
    }
}
// (r,t) -> (p,m) on the A fragments of one 8-row block: p = r - i t, m = r + i t (the factor 1/2 is applied
// in the epilogue)
template <int NPROD, int NB>
__device__ __forceinline__ void dt_mix_block(DtFrag<NPROD, NB> &f, int i) {
    const double2 r = f.a[0][i], tt = f.a[NPROD - 1][i];
    f.a[0][i] = make_double2(r.x + tt.y, r.y - tt.x);
    f.a[NPROD - 1][i] = make_double2(r.x - tt.y, r.y + tt.x);
}
// DMMAs of one k4 step; when MIX, the combination of the NEXT step's fragments (loaded before this call) is
// interleaved block by block, so that the DADDs neither wait for their loads nor sit in front of the DMMAs
template <int NPROD, bool MIX, int NB>
__device__ __forceinline__ void dt_mma_frag(double (&acc)[NPROD][4][DtCfg<NPROD>::NI][2][2],
                                            const DtFrag<NPROD, NB> &f, DtFrag<NPROD, NB> &next, bool mix_next) {
    using C = DtCfg<NPROD>;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p)
#pragma unroll
            for (int n = 0; n < C::NI; ++n) {
                dt_dmma(acc[p][i][n][0][0], acc[p][i][n][0][1], f.a[p][i].x, f.b[p][n]);
                dt_dmma(acc[p][i][n][1][0], acc[p][i][n][1][1], f.a[p][i].y, f.b[p][n]);
            }
        if (NPROD == 2 && MIX && mix_next) dt_mix_block<NPROD, NB>(next, i);
    }
}

// K loop of one tile for a consumer warp that owns NB 8-row blocks of it.  Software pipelined across the
// stages of the ring: the fragments of k4 step q+1 (possibly the first of the next stage, after its full
// barrier) are loaded into the other register buffer before the DMMAs of step q are issued.
template <int NPROD, bool MIX, int NB>
__device__ __forceinline__ void dt_kloop(double (&acc)[NPROD][4][DtCfg<NPROD>::NI][2][2], int KT, uint32_t &s,
                                         uint32_t &ph, uint32_t smem_base, uint32_t bar_base, uint32_t a_off0,
                                         uint32_t a_off1, uint32_t b_off, int lane) {
    using C = DtCfg<NPROD>;
    if (NB == 0) {          // no rows of this warp in the tile: keep the ring turning
        for (int kt = 0; kt < KT; ++kt) {
            dt_mbar_wait(bar_base + 8 * s, ph);
            __syncwarp();
            if (lane == 0) dt_mbar_arrive(bar_base + 8 * (DT_STAGES + s));
            if (++s == DT_STAGES) { s = 0; ph ^= 1u; }
        }
        return;
    }
    DtFrag<NPROD, NB> f0, f1;
    dt_mbar_wait(bar_base + 8 * s, ph);
    uint32_t sa = smem_base + s * DT_STAGE_BYTES, sb = sa + C::A_BYTES + b_off;
    dt_load_frag<NPROD, MIX, NB>(f0, sa, sb, 0, a_off0, a_off1);
    if (NPROD == 2 && MIX) {
#pragma unroll
        for (int i = 0; i < NB; ++i) dt_mix_block<NPROD, NB>(f0, i);
    }
    for (int kt = 0; kt < KT; ++kt) {
        const uint32_t empty = bar_base + 8 * (DT_STAGES + s);
        dt_load_frag<NPROD, MIX, NB>(f1, sa, sb, 1, a_off0, a_off1);
        dt_mma_frag<NPROD, MIX, NB>(acc, f0, f1, true);
        dt_load_frag<NPROD, MIX, NB>(f0, sa, sb, 2, a_off0, a_off1);
        dt_mma_frag<NPROD, MIX, NB>(acc, f1, f0, true);
        dt_load_frag<NPROD, MIX, NB>(f1, sa, sb, 3, a_off0, a_off1);
        dt_mma_frag<NPROD, MIX, NB>(acc, f0, f1, true);
        // next stage: its first fragments are loaded before this stage is released
        if (++s == DT_STAGES) { s = 0; ph ^= 1u; }
        const bool more = kt + 1 < KT;
        if (more) {
            dt_mbar_wait(bar_base + 8 * s, ph);
            sa = smem_base + s * DT_STAGE_BYTES; sb = sa + C::A_BYTES + b_off;
            dt_load_frag<NPROD, MIX, NB>(f0, sa, sb, 0, a_off0, a_off1);
        }
        dt_mma_frag<NPROD, MIX, NB>(acc, f1, f0, more);
        __syncwarp();
        if (lane == 0) dt_mbar_arrive(empty);
    }
}

// producer side of one tile: KT stages of A boxes + packed M blocks into the ring
template <int NPROD>
__device__ __forceinline__ void dt_produce_tile(const DtParams &P, const DtJob &J, int strip, int m0, uint32_t &s,
                                                uint32_t &ph, uint32_t smem_base, uint32_t bar_base) {
    using C = DtCfg<NPROD>;
    const int n0 = strip * C::BN;
    for (int kt = 0; kt < P.KT; ++kt) {
        const uint32_t full = bar_base + 8 * s, empty = bar_base + 8 * (DT_STAGES + s);
        if (P.producer_sleep) { while (!dt_mbar_try_wait(empty, ph ^ 1u)) __nanosleep(200); }
        else dt_mbar_wait(empty, ph ^ 1u);
        dt_mbar_expect_tx(full, C::TX_BYTES);
        const uint32_t a_dst = smem_base + s * DT_STAGE_BYTES;
        const uint32_t b_dst = a_dst + C::A_BYTES;
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
#pragma unroll
            for (int bx = 0; bx < 2; ++bx)
                dt_tma_2d(a_dst + (p * 2 + bx) * C::A_BOX_BYTES, &J.mapA[p], full, 2 * (kt * DT_BK + bx * 8), m0);
            const double *Bp = J.B[p] + ((size_t)(kt * 4) * (P.Np / 4) + n0 / 4) * 16;
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
                dt_bulk_1d(b_dst + (p * 4 + kb) * C::B_KB_BYTES, Bp + (size_t)kb * (P.Np / 4) * 16, C::B_KB_BYTES, full);
        }
        if (++s == DT_STAGES) { s = 0; ph ^= 1u; }
    }
}

// consumer side of one tile: K loop + epilogue.  Lane (g = lane/4, t = lane%4) of warp (wm, wn) ends up with
// rows iz = m0 + wm*32 + i*8 + rho(g), columns n0 + wn*(BN/4) + n*8 + 2t + {0,1}.
template <int NPROD, bool MIX>
__device__ __forceinline__ void dt_consume_tile(const DtParams &P, const DtJob &J, int strip, int m0, int cnt,
                                                uint32_t &s, uint32_t &ph, uint32_t smem_base, uint32_t bar_base,
                                                int lane, int warp, long long *t_k) {
    using C = DtCfg<NPROD>;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int rho = ((g & 1) << 2) | (g >> 1);              // tile row (within an 8-row block) that feeds MMA row g
    // A fragment: byte offset inside a box of row (wm*32 + i*8 + rho), chunk ((k4&1)*4 + t) ^ (row & 7):
    // block i adds i*1024 bytes, the upper half of the box (k4 odd) flips bit 6 of the chunk field
    const uint32_t a_off0 = (uint32_t)((wm * 32 + rho) * 128 + ((t ^ rho) << 4));
    const uint32_t a_off1 = a_off0 ^ 64u;
    // B fragment: byte offset inside a k4 block of the strip: ((nb*4 + t)*4 + c)*8, nb = col/4 + (g>>2), c = g&3;
    // MMA block n adds 8 columns = 256 bytes
    const uint32_t b_off = (uint32_t)(((((wn * (C::BN / 4)) >> 2) + (g >> 2)) * 4 + t) * 4 + (g & 3)) * 8u;

    int nblk = cnt * 2 - wm * 4;                             // 8-row blocks of this warp inside the tile (<= 4)
    if (nblk > 4) nblk = 4;
    double acc[NPROD][4][C::NI][2][2];                       // [prod][i][n][re|im][2]
#pragma unroll
    for (int p = 0; p < NPROD; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int n = 0; n < C::NI; ++n)
                acc[p][i][n][0][0] = acc[p][i][n][0][1] = acc[p][i][n][1][0] = acc[p][i][n][1][1] = 0.;

    // three copies of the K loop (4, 2 or 0 of this warp's 8-row blocks lie inside the tile): the partial
    // tiles at the ends of a CTA's share cost their share of DMMAs, nothing is predicated off
    if (nblk >= 4)
        dt_kloop<NPROD, MIX, 4>(acc, P.KT, s, ph, smem_base, bar_base, a_off0, a_off1, b_off, lane);
    else if (nblk >= 2)
        dt_kloop<NPROD, MIX, 2>(acc, P.KT, s, ph, smem_base, bar_base, a_off0, a_off1, b_off, lane);
    else
        dt_kloop<NPROD, MIX, 0>(acc, P.KT, s, ph, smem_base, bar_base, a_off0, a_off1, b_off, lane);
#ifdef DT_DEBUG
    if (t_k) *t_k = clock64();
#endif

    const int n0 = strip * C::BN;
    const bool wide = (P.Nr & 1) == 0 && (((uintptr_t)J.out[0] | (uintptr_t)J.out[NPROD - 1]) & 31) == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int iz = m0 + wm * 32 + i * 8 + rho;
        if (i >= nblk || iz >= P.Nz) continue;
        double rs = J.rowscale ? __ldg(J.rowscale + iz) : 1.;
        if (NPROD == 2 && MIX) rs *= 0.5;
#pragma unroll
        for (int n = 0; n < C::NI; ++n) {
            const int col = n0 + wn * (C::BN / 4) + n * 8 + 2 * t;      // this lane: columns col, col + 1
            if (col >= P.Nr) continue;
            const size_t o = (size_t)iz * P.Nr + col;
            double2 v0[NPROD], v1[NPROD];                                  // [output array] at col, col + 1
            if (NPROD == 1) {
                v0[0] = make_double2(rs * acc[0][i][n][0][0], rs * acc[0][i][n][1][0]);
                v1[0] = make_double2(rs * acc[0][i][n][0][1], rs * acc[0][i][n][1][1]);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double pr = acc[0][i][n][0][h], pi = acc[0][i][n][1][h];
                    const double qr = acc[NPROD - 1][i][n][0][h], qi = acc[NPROD - 1][i][n][1][h];
                    double2 x, y;
                    if (MIX) {
                        x = make_double2(rs * pr, rs * pi);
                        y = make_double2(rs * qr, rs * qi);
                    } else {
                        x = make_double2(rs * (pr + qr), rs * (pi + qi));            // r = P + Q
                        y = make_double2(rs * -(pi - qi), rs * (pr - qr));           // t = i (P - Q)
                    }
                    if (h == 0) { v0[0] = x; v0[NPROD - 1] = y; } else { v1[0] = x; v1[NPROD - 1] = y; }
                }
            }
#pragma unroll
            for (int p = 0; p < NPROD; ++p) {
                if (wide) {          // 32-byte store: the 4 lanes of a row write one full 128-byte line
                    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(J.out[p] + o), "d"(v0[p].x),
                                 "d"(v0[p].y), "d"(v1[p].x), "d"(v1[p].y) : "memory");
                } else {
                    J.out[p][o] = v0[p];
                    if (col + 1 < P.Nr) J.out[p][o + 1] = v1[p];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(DT_THREADS, 1)
k_dht_tma(const __grid_constant__ DtParams P) {
    extern __shared__ unsigned char dt_smem_raw[];
    // 1024-byte alignment: required by SWIZZLE_128B
    unsigned char *smem = (unsigned char *)(((uintptr_t)dt_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = dt_smem_u32(smem);
    const uint32_t bar_base = smem_base + DT_STAGES * DT_STAGE_BYTES;      // full[STAGES] | empty[STAGES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < DT_STAGES; ++s) {
            dt_mbar_init(bar_base + 8 * s, 1);
            dt_mbar_init(bar_base + 8 * (DT_STAGES + s), DT_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= DT_CONSUMER_WARPS) {
        // =========================================================== producer warpgroup: hands its registers to the
        // consumers (setmaxnreg is warpgroup-wide), then one elected lane feeds the ring
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == DT_CONSUMER_WARPS && lane == 0) {
            DtWalk walk(P);
            int job, strip, m0, cnt;
            uint32_t s = 0, ph = 0;
            while (walk.next(P, job, strip, m0, cnt)) {
                const DtJob &J = P.job[job];
                if (J.kind == DT_SCALAR) dt_produce_tile<1>(P, J, strip, m0, s, ph, smem_base, bar_base);
                else dt_produce_tile<2>(P, J, strip, m0, s, ph, smem_base, bar_base);
            }
        }
        return;
    }

    // =============================================================== consumer warps (2 warpgroups)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // the two consumer warpgroups (rows 0-31 / 32-63 of every tile) run about one ring stage out of step: while one
    // group is blocked issuing its epilogue stores (128 KB per tile through the SM's store path) the other still has
    // K stages to multiply, so the DMMA pipe does not drain at tile boundaries.  The offset is set once here; the
    // ring (empty barriers need both groups) bounds it, and it is self-restoring: a group alone on the pipe runs
    // at twice the speed.
    if (P.stagger_ns > 0 && warp >= DT_CONSUMER_WARPS / 2) __nanosleep(P.stagger_ns);
    DtWalk walk(P);
    int job, strip, m0, cnt;
    uint32_t s = 0, ph = 0;
#ifdef DT_DEBUG
    long long t_begin = 0, t_k = 0, t_e = 0, n_tiles = 0;
    if (P.dbg) t_begin = clock64();
#endif
    while (walk.next(P, job, strip, m0, cnt)) {
        const DtJob &J = P.job[job];
        long long *tkp = nullptr;
#ifdef DT_DEBUG
        long long t0 = 0, t1 = 0;
        if (P.dbg) { t0 = clock64(); tkp = &t1; }
#endif
        if (J.kind == DT_SCALAR)
            dt_consume_tile<1, false>(P, J, strip, m0, cnt, s, ph, smem_base, bar_base, lane, warp, tkp);
        else if (J.kind == DT_INVERSE)
            dt_consume_tile<2, false>(P, J, strip, m0, cnt, s, ph, smem_base, bar_base, lane, warp, tkp);
        else
            dt_consume_tile<2, true>(P, J, strip, m0, cnt, s, ph, smem_base, bar_base, lane, warp, tkp);
#ifdef DT_DEBUG
        if (P.dbg) { t_k += t1 - t0; t_e += clock64() - t1; ++n_tiles; }
#endif
    }
#ifdef DT_DEBUG
    if (P.dbg && lane == 0) {
        long long *d = P.dbg + ((size_t)blockIdx.x * DT_CONSUMER_WARPS + warp) * 4;
        d[0] = clock64() - t_begin; d[1] = t_k; d[2] = t_e; d[3] = n_tiles;
    }
#endif
}

// M [Nr,Nr] row-major -> fragment order [Kp/4][Np/4][4][4], zero padded
__global__ void k_dht_pack(const double *__restrict__ M, double *__restrict__ Pk, int Nr, int Kp, int Np) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Kp * Np) return;
    const int c = idx & 3, t = (idx >> 2) & 3;
    const int nb = (idx >> 4) % (Np / 4), kb = (idx >> 4) / (Np / 4);
    const int k = 4 * kb + t, n = 4 * nb + c;
    Pk[idx] = (k < Nr && n < Nr) ? M[(size_t)k * Nr + n] : 0.;
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*dt_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static dt_encode_fn g_dt_encode = nullptr;
static int g_dt_state = 0;          // 0 unknown, 1 usable, -1 unavailable
static std::mutex g_dt_mutex;
struct DtPacked { double *ptr; int Nr; };
static std::map<const void *, DtPacked> g_dt_packed;                 // keyed by the row-major matrix pointer
struct DtMapKey {
    const void *p; int Nz, Nr, box_d, box_rows, swz;
    bool operator<(const DtMapKey &o) const {
        if (p != o.p) return p < o.p;
        if (Nz != o.Nz) return Nz < o.Nz;
        if (Nr != o.Nr) return Nr < o.Nr;
        if (box_d != o.box_d) return box_d < o.box_d;
        if (box_rows != o.box_rows) return box_rows < o.box_rows;
        return swz < o.swz;
    }
};
static std::map<DtMapKey, CUtensorMap> g_dt_maps;

static int dt_resolve() {
    if (g_dt_state) return g_dt_state;
    const char *impl = getenv("B2_DHT_IMPL");
    if (impl && !strcmp(impl, "legacy")) { g_dt_state = -1; return g_dt_state; }
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        fprintf(stderr, "fbpic_b200: cuTensorMapEncodeTiled not available, Hankel transform uses the LDG-staged kernel\n");
        g_dt_state = -1;
        return g_dt_state;
    }
    g_dt_encode = (dt_encode_fn)fn;
    g_dt_state = 1;
    return g_dt_state;
}

// called by b2_free / b2_memcpy_* / b2_memset (b2_runtime.cu): the contents behind `p` changed or died
void b2_dht_forget(const void *p) {
    std::lock_guard<std::mutex> lk(g_dt_mutex);
    auto it = g_dt_packed.find(p);
    if (it != g_dt_packed.end()) { cudaFree(it->second.ptr); g_dt_packed.erase(it); }
    for (auto m = g_dt_maps.begin(); m != g_dt_maps.end();) {
        if (m->first.p == p) m = g_dt_maps.erase(m); else ++m;
    }
}

static inline int dt_round_up(int v, int q) { return (v + q - 1) / q * q; }

static int dt_packed_matrix(const double *M, int Nr, cudaStream_t s, const double **out) {
    auto it = g_dt_packed.find(M);
    if (it != g_dt_packed.end() && it->second.Nr == Nr) { *out = it->second.ptr; return 0; }
    if (it != g_dt_packed.end()) { cudaFree(it->second.ptr); g_dt_packed.erase(it); }
    const int Kp = dt_round_up(Nr, DT_BK), Np = dt_round_up(Nr, 128);
    double *pk = nullptr;
    B2_CUDA(cudaMalloc(&pk, sizeof(double) * (size_t)Kp * Np));
    k_dht_pack<<<(Kp * Np + 255) / 256, 256, 0, s>>>(M, pk, Nr, Kp, Np);
    B2_LAUNCHED();
    g_dt_packed[M] = DtPacked{pk, Nr};
    *out = pk;
    return 0;
}

// Tensor map of a complex [Nz,Nr] field array seen as doubles [Nz][2*Nr]: boxes of box_rows rows x box_d doubles
// (cached per array / shape / box).  Shared with b2_gather_pipe.cu.  rc 1: the TMA descriptor API is unavailable.
int b2_tma_field_map(const void *A, int Nz, int Nr, int box_d, int box_rows, int swizzle128, CUtensorMap *out) {
    if (dt_resolve() < 0) return 1;
    const DtMapKey key{A, Nz, Nr, box_d, box_rows, swizzle128};
    auto it = g_dt_maps.find(key);
    if (it != g_dt_maps.end()) { *out = it->second; return 0; }
    if (g_dt_maps.size() > 4096) g_dt_maps.clear();
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * Nr, (cuuint64_t)Nz};       // doubles per row, rows
    const cuuint64_t strides[1] = {(cuuint64_t)Nr * 16};                   // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)box_d, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_dt_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(A), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return b2_fail((int)r, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    g_dt_maps[key] = m;
    *out = m;
    return 0;
}
std::mutex &b2_tma_mutex() { return g_dt_mutex; }
static int dt_tensor_map(const void *A, int Nz, int Nr, CUtensorMap *out) {
    return b2_tma_field_map(A, Nz, Nr, 16, DT_BM, 1, out);       // 8 complex x 64 rows, SWIZZLE_128B
}

struct DtHostJob {           // one job in terms of raw pointers
    const void *in[2];
    void *out[2];
    const double *M[2];
    const double *rowscale;
    int kind;                // DT_SCALAR / DT_INVERSE / DT_FORWARD
};

static int dt_launch(b2_ctx *ctx, const DtHostJob *jobs, int njobs, int Nz, int Nr, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        B2_CUDA(cudaFuncSetAttribute(k_dht_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES));
        attr_set = true;
    }
    DtParams P;
    memset(&P, 0, sizeof(P));
    P.njobs = njobs; P.Nz = Nz; P.Nr = Nr;
    P.Np = dt_round_up(Nr, 128);
    P.KT = dt_round_up(Nr, DT_BK) / DT_BK;
    P.blocks64 = (Nz + DT_BM - 1) / DT_BM;
    long long units = 0;
    double products = 0.;
    for (int k = 0; k < njobs; ++k) {
        const int nprod = jobs[k].kind == DT_SCALAR ? 1 : 2;
        for (int p = 0; p < nprod; ++p) {
            int rc = dt_tensor_map(jobs[k].in[p], Nz, Nr, &P.job[k].mapA[p]);
            if (rc) return rc;
            rc = dt_packed_matrix(jobs[k].M[p], Nr, s, &P.job[k].B[p]);
            if (rc) return rc;
            P.job[k].out[p] = (double2 *)jobs[k].out[p];
        }
        P.job[k].rowscale = jobs[k].rowscale;
        P.job[k].kind = jobs[k].kind;
        const int bn = 128 / nprod;
        P.job[k].n_strips = (Nr + bn - 1) / bn;
        P.job[k].unit_begin = (int)units;
        units += (long long)P.job[k].n_strips * P.blocks64 * (DT_BM / 16);
        if (units > 0x7fffffffLL) return b2_fail(-3, "b2_dht: grid too large", __FILE__, __LINE__);
        products += nprod;
    }
    P.total_units = (int)units;
    long long grid = ctx && ctx->sm_count > 0 ? ctx->sm_count : 148;
    if (grid > units) grid = units;
#ifdef DT_DEBUG
    static const bool debug = getenv("B2_DHT_DEBUG") != nullptr;
#else
    const bool debug = false;     // the cycle breakdown needs a -DDT_DEBUG build of this file
#endif
    P.dbg = nullptr;
    static const int psleep = getenv("B2_DHT_PSLEEP") ? atoi(getenv("B2_DHT_PSLEEP")) : 1;
    P.producer_sleep = psleep;
    static const int stagger = getenv("B2_DHT_STAGGER_NS") ? atoi(getenv("B2_DHT_STAGGER_NS")) : 0;   // measured: no gain (profiles/README)
    P.stagger_ns = (unsigned)(stagger > 0 ? stagger : 0);
    if (debug) {
        B2_CUDA(cudaMalloc(&P.dbg, sizeof(long long) * grid * DT_CONSUMER_WARPS * 4));
        B2_CUDA(cudaMemsetAsync(P.dbg, 0, sizeof(long long) * grid * DT_CONSUMER_WARPS * 4, s));
    }
    k_dht_tma<<<(unsigned)grid, DT_THREADS, DT_SMEM_BYTES, s>>>(P);
    B2_LAUNCHED();
    if (debug) {       // cycle breakdown of the consumer warps (diagnostic build, synchronises)
        std::vector<long long> h((size_t)grid * DT_CONSUMER_WARPS * 4);
        B2_CUDA(cudaStreamSynchronize(s));
        B2_CUDA(cudaMemcpy(h.data(), P.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(P.dbg);
        double tot = 0, tk = 0, te = 0, nt = 0, tmax = 0;
        for (size_t w = 0; w < h.size() / 4; ++w) {
            tot += h[4 * w]; tk += h[4 * w + 1]; te += h[4 * w + 2]; nt += h[4 * w + 3];
            if (h[4 * w] > tmax) tmax = (double)h[4 * w];
        }
        const double nw = (double)h.size() / 4;
        // ideal K-loop cycles of a warp: its share of DMMAs x 16 cycles x 2 warps per scheduler
        const double dmma_per_warp = products * 2. * Nz * ((Nr + 7) / 8) * ((Nr + 3) / 4) / 8. / nw;
        fprintf(stderr, "[dht debug] jobs=%d products=%.0f grid=%lld: per warp avg total %.0f (max %.0f) kloop %.0f "
                        "epilogue %.0f cycles, %.2f tiles; ideal kloop %.0f (DMMA x16 x2)\n",
                njobs, products, grid, tot / nw, tmax, tk / nw, te / nw, nt / nw, dmma_per_warp * 32.);
    }
    return 0;
}

// entry used by b2_dht.cu: a mixed job list (kinds as B2_DHT_* of the header) in ONE launch; returns 1 if the TMA
// path is not usable (caller falls back to k_dht)
int b2_dht_tma_run(b2_ctx *ctx, const b2_dht_job *jobs, int njobs, int Nz, int Nr, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_dt_mutex);
    if (dt_resolve() < 0 || Nr < 8) return 1;
    if (njobs > DT_MAX_JOBS) return 1;
    DtHostJob hj[DT_MAX_JOBS];
    for (int k = 0; k < njobs; ++k) {
        const b2_dht_job &u = jobs[k];
        DtHostJob &h = hj[k];
        h.in[0] = u.in1; h.in[1] = u.in2; h.out[0] = u.out1; h.out[1] = u.out2;
        h.M[0] = u.M1; h.M[1] = u.M2; h.rowscale = u.rowscale;
        if (u.kind == B2_DHT_SCALAR) h.kind = DT_SCALAR;
        else if (u.kind == B2_DHT_PM_TO_RT) h.kind = DT_INVERSE;
        else if (u.kind == B2_DHT_RT_TO_PM) h.kind = DT_FORWARD;
        else return b2_fail(-3, "b2_dht_batch: unknown job kind", __FILE__, __LINE__);
    }
    return dt_launch(ctx, hj, njobs, Nz, Nr, s);
}

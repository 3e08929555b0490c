// b2_deposit.cu -- charge / current deposition on cell-sorted particles.
//
// Replaces deposit_{rho,J}_gpu_{linear,cubic}[_one_mode] (fbpic/particles/deposition/
// cuda_methods.py:28,202,466,751 and cuda_methods_one_mode.py) -- one thread per cell, TPB 8,
// uncoalesced per-particle loads, one pass per azimuthal mode for Nm != 2 -- and, in its PERMUTE
// flavour, also write_sorting_buffer / rearrange_particle_arrays (cuda_sorting.py:193-213,
// particles.py:510-555).
//
// B200 design: a CTA owns DEP_TPB consecutive cells of the sorted order, i.e. ONE contiguous range
// of sorted particles [prefix_sum[c0-1], prefix_sum[c0+TPB-1]).  The range is streamed through
// shared memory in chunks with a 2-stage cp.async pipeline (the next chunk lands while the current
// one is reduced).  Each thread reduces the particles of its own cell into registers -- every
// particle of a cell scatters to the same stencil points, so the cell's sums are formed first: no
// atomics, no shuffles -- visiting them in a lane-rotated order so that the 32 lanes hit different
// shared-memory banks.  All azimuthal modes are handled in the same pass.  One fp64 RED
// (red.global.add.f64, measured ~590 G/s on B200) per cell, stencil point and real value flushes
// the sums: REDs per CELL instead of per particle.
//
//   PERMUTE   : the chunk is gathered through the sort permutation (idx32) from the UNSORTED
//               attribute arrays and also written back, coalesced, to the sorted arrays: the
//               permutation pass and the deposition share one read of the particle data.
//   DISPLACED : (rho, linear) the particles were sorted half a step ago and have moved a little since:
//               a particle that is still in the cell it was sorted into goes into the cell's register
//               sums, one that crossed a cell boundary issues its own 4 x (2Nm-1) REDs.
//               Saves the second sort of the PIC cycle (particles.py:866-871 sorts before every
//               deposit; main.py:511,528).
//
// Boundary folds follow fbpic/fields/numba_methods.py:410-461 / cuda_methods.py:167-177,670-691:
// z periodic; cells below the axis fold to -(1+ir) with the flip sign (-1)^m (rho, Jz) or
// -(-1)^m (Jr, Jt) (particle_shapes.py:33-36,76-79); cells beyond Nr-1 clamp to Nr-1.
#include "b2_common.cuh"

#define DEP_TPB 128
#define DEP_CHUNK 256

struct B2DepGrids {
    double2 *g[3 * B2_MAX_MODES];   // rho: [m] ; J: [m][Jr,Jt,Jz]
};
struct B2DepPtrs {
    const double *src[8];           // x,y,z,w,ux,uy,uz,inv_gamma (sorted, or unsorted when PERMUTE)
    double *dst[8];                 // PERMUTE: sorted destination arrays
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// linear shape factors of one particle (particle_shapes.py:17-37); flips are applied at flush time
__device__ __forceinline__ void lin_shapes(const B2Cyl &c, double beta0, double beta_hi,
                                           double sz[2], double sr0[2], double sr1[2]) {
    sz[0] = ceil(c.z_cell) - c.z_cell;
    sz[1] = 1. - sz[0];
    const double u = c.r_cell - (ceil(c.r_cell) - 1.);
    const double base = 1. - u, t = (1. - u) * u;
    sr0[0] = base + beta0 * t;   sr0[1] = 1. - sr0[0];
    sr1[0] = base + beta_hi * t; sr1[1] = 1. - sr1[0];
}

template <int NM, int NVM, int OZ, int OR>
__device__ __forceinline__ void add_footprint(double (&acc)[4][4][NVM], const double sz[2], const double sr0[2],
                                              const double sr1[2], const double V[NVM]) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double w0 = sz[a] * sr0[b], w1 = sz[a] * sr1[b];
            acc[OZ + a][OR + b][0] += w0 * V[0];
#pragma unroll
            for (int v = 1; v < NVM; ++v) acc[OZ + a][OR + b][v] += w1 * V[v];
        }
}

// NM modes; IS_J: J (x,y,z,w,ux,uy,uz,inv_gamma) or rho (x,y,z,w); NPT = 2 linear, 4 cubic;
// components [C0, C0+NC) of (rho) or (Jr,Jt,Jz) are deposited by this launch.
template <int NM, bool IS_J, int NPT, int C0, int NC, bool PERMUTE, bool DISPLACED>
__global__ void __launch_bounds__(DEP_TPB)
k_deposit(B2DepPtrs P, const int32_t *__restrict__ idx32, double q,
          double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, B2DepGrids G,
          const int32_t *__restrict__ prefix_sum, const double *__restrict__ ruyten0,
          const double *__restrict__ ruyten_hi, int ncells) {
    constexpr int NNEED = IS_J ? 8 : 4;                 // attributes the deposition reads
    constexpr int NATTR = PERMUTE ? 8 : NNEED;          // attributes staged (PERMUTE moves all 8)
    constexpr int NVM = 2 * NM - 1;                     // real values per component
    constexpr int NV = NC * NVM;
    constexpr int FP = NPT;                             // register footprint (points per dimension)
    static_assert(!DISPLACED || (!IS_J && NPT == 2 && NC == 1), "DISPLACED: rho, linear only");
    __shared__ double sm[2][NATTR][DEP_CHUNK];

    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * DEP_TPB;
    const int cell = c0 + tid;
    const bool valid = cell < ncells;
    const int c_last = min(c0 + DEP_TPB, ncells) - 1;
    const int pb = (c0 == 0) ? 0 : prefix_sum[c0 - 1];
    const int pe = prefix_sum[c_last];
    if (pe == pb) return;                               // no particle in this CTA's cells (uniform exit)
    int s = 0, e = 0;
    if (valid) {
        s = (cell == 0) ? 0 : prefix_sum[cell - 1];
        e = prefix_sum[cell];
    }
    const int iz_u = valid ? cell / (Nr + 1) : 0;
    const int ir_u = valid ? cell - iz_u * (Nr + 1) : 0;
    const double beta0_c = ruyten0[ir_u], beta_hi_c = ruyten_hi[ir_u];

    double acc[FP][FP][NV];
#pragma unroll
    for (int a = 0; a < FP; ++a)
#pragma unroll
        for (int b = 0; b < FP; ++b)
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[a][b][v] = 0.;

    const int nchunk = (pe - pb + DEP_CHUNK - 1) / DEP_CHUNK;
    // source index of the elements this thread stages for the chunk that will be issued next
    // (PERMUTE: the sort permutation, loaded one chunk ahead so that its latency is hidden)
    size_t jn[DEP_CHUNK / DEP_TPB];
    auto load_idx = [&](int ch) {
        const int q0 = pb + ch * DEP_CHUNK;
#pragma unroll
        for (int r = 0; r < DEP_CHUNK / DEP_TPB; ++r) {
            const int i = tid + r * DEP_TPB;
            jn[r] = (PERMUTE && q0 + i < pe) ? (size_t)__ldg(idx32 + q0 + i) : (size_t)(q0 + i);
        }
    };
    // issue the asynchronous copies of chunk `ch` into stage `st` (uses jn)
    auto issue = [&](int ch, int st) {
        const int q0 = pb + ch * DEP_CHUNK;
        const int cnt = min(DEP_CHUNK, pe - q0);
#pragma unroll
        for (int r = 0; r < DEP_CHUNK / DEP_TPB; ++r) {
            const int i = tid + r * DEP_TPB;
            if (i < cnt) {
#pragma unroll
                for (int k = 0; k < NATTR; ++k) cp_async8(&sm[st][k][i], P.src[k] + jn[r]);
            }
        }
        cp_async_commit();
    };

    load_idx(0);
    issue(0, 0);
    if (nchunk > 1) load_idx(1);
    for (int ch = 0; ch < nchunk; ++ch) {
        const int st = ch & 1;
        const int q0 = pb + ch * DEP_CHUNK;
        const int q1 = min(q0 + DEP_CHUNK, pe);
        if (ch + 1 < nchunk) {
            issue(ch + 1, st ^ 1);
            if (ch + 2 < nchunk) load_idx(ch + 2);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (PERMUTE) {          // coalesced write-out of the sorted chunk
#pragma unroll
            for (int r = 0; r < DEP_CHUNK / DEP_TPB; ++r) {
                const int i = tid + r * DEP_TPB;
                if (i < q1 - q0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) P.dst[k][q0 + i] = sm[st][k][i];
                }
            }
        }
        // ---- reduce my cell's particles that lie in this chunk ----
        const int lo = max(s, q0), hi = min(e, q1);
        const int cnt = hi - lo;
        if (cnt > 0) {
            int k = lo + (int)(tid & 31) % cnt;          // lane-rotated start: spreads smem banks
            for (int jj = 0; jj < cnt; ++jj) {
                const int i = k - q0;
                const double xj = sm[st][0][i], yj = sm[st][1][i], zj = sm[st][2][i];
                const double wj = q * sm[st][3][i];
                const B2Cyl c = b2_cyl(xj, yj, zj, invdz, zmin, invdr, rmin);
                // per-particle values: comp-major, then [m=0 | Re m=1, Im m=1 | ...]
                double V[NV];
                if (!IS_J) {
                    V[0] = wj;
                } else {
                    const double f = wj * B2_C_LIGHT * sm[st][NNEED - 1][i];
                    const double uxj = sm[st][4][i], uyj = sm[st][5][i], uzj = sm[st][6][i];
                    const double j3[3] = {f * (c.cs * uxj + c.sn * uyj), f * (c.cs * uyj - c.sn * uxj), f * uzj};
#pragma unroll
                    for (int kc = 0; kc < NC; ++kc) V[kc * NVM] = j3[C0 + kc];
                }
#pragma unroll
                for (int kc = 0; kc < NC; ++kc) {
                    double re = V[kc * NVM], im = 0.;
#pragma unroll
                    for (int m = 1; m < NM; ++m) {
                        const double nre = c.cs * re - c.sn * im, nim = c.cs * im + c.sn * re;
                        re = nre; im = nim;
                        V[kc * NVM + 2 * m - 1] = re;
                        V[kc * NVM + 2 * m] = im;
                    }
                }
                if constexpr (DISPLACED) {
                    // where is this particle now, relative to the cell it was sorted into?
                    int irp = (int)ceil(c.r_cell), izp = (int)ceil(c.z_cell);
                    if (irp > Nr) irp = Nr;
                    if (izp < 0) izp += Nz; else if (izp > Nz - 1) izp -= Nz;
                    int dz = izp - iz_u, dr = irp - ir_u;
                    if (dz > Nz / 2) dz -= Nz; else if (dz < -(Nz / 2)) dz += Nz;
                    double sz[2], sr0[2], sr1[2];
                    lin_shapes(c, __ldg(ruyten0 + irp), __ldg(ruyten_hi + irp), sz, sr0, sr1);
                    if (dz == 0 && dr == 0) {
                        // still in the cell it was sorted into: register sums, as in the sorted kernel
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                const double w0 = sz[a] * sr0[b], w1 = sz[a] * sr1[b];
                                acc[a][b][0] += w0 * V[0];
#pragma unroll
                                for (int v = 1; v < NVM; ++v) acc[a][b][v] += w1 * V[v];
                            }
                    } else {
                        // left its cell since the sort: per-particle REDs (fp64 REDG is cheap on B200)
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            int ir = irp - 1 + b;
                            const bool below = ir < 0;
                            if (below) ir = -(1 + ir);
                            if (ir > Nr - 1) ir = Nr - 1;
#pragma unroll
                            for (int a = 0; a < 2; ++a) {
                                int iz = izp - 1 + a;
                                if (iz < 0) iz += Nz;
                                const size_t o = (size_t)iz * Nr + ir;
#pragma unroll
                                for (int m = 0; m < NM; ++m) {
                                    const double sgn = (below && (m & 1)) ? -1. : 1.;
                                    double *p = (double *)(G.g[m] + o);
                                    if (m == 0) atomicAdd(p, sgn * sz[a] * sr0[b] * V[0]);
                                    else {
                                        atomicAdd(p, sgn * sz[a] * sr1[b] * V[2 * m - 1]);
                                        atomicAdd(p + 1, sgn * sz[a] * sr1[b] * V[2 * m]);
                                    }
                                }
                            }
                        }
                    }
                } else {
                    // shape factors (particle_shapes.py:17-80), flip applied once per cell at flush
                    double sz[NPT], sr0[NPT], sr1[NPT];
                    if (NPT == 2) {
                        lin_shapes(c, beta0_c, beta_hi_c, sz, sr0, sr1);
                    } else {
                        const double uz_ = c.z_cell - (ceil(c.z_cell) - 2.) - 1.;
                        const double vz = 1. - uz_;
                        sz[0] = (1. / 6.) * (vz * vz * vz);
                        sz[1] = (1. / 6.) * (3. * (uz_ * uz_ * uz_) - 6. * (uz_ * uz_) + 4.);
                        sz[NPT - 2] = (1. / 6.) * (3. * (vz * vz * vz) - 6. * (vz * vz) + 4.);
                        sz[NPT - 1] = (1. / 6.) * (uz_ * uz_ * uz_);
                        const double u = c.r_cell - (ceil(c.r_cell) - 2.) - 1.;
                        const double v = 1. - u, t = (1. - u) * u;
                        const double s0 = (1. / 6.) * (v * v * v);
                        const double s1 = (1. / 6.) * (3. * (u * u * u) - 6. * (u * u) + 4.);
                        const double s2 = (1. / 6.) * (3. * (v * v * v) - 6. * (v * v) + 4.);
                        const double s3 = (1. / 6.) * (u * u * u);
                        sr0[0] = s0; sr0[1] = s1 + beta0_c * t;   sr0[NPT - 2] = s2 - beta0_c * t;   sr0[NPT - 1] = s3;
                        sr1[0] = s0; sr1[1] = s1 + beta_hi_c * t; sr1[NPT - 2] = s2 - beta_hi_c * t; sr1[NPT - 1] = s3;
                    }
#pragma unroll
                    for (int a = 0; a < NPT; ++a)
#pragma unroll
                        for (int b = 0; b < NPT; ++b) {
                            const double w0 = sz[a] * sr0[b], w1 = sz[a] * sr1[b];
#pragma unroll
                            for (int kc = 0; kc < NC; ++kc) {
                                acc[a][b][kc * NVM] += w0 * V[kc * NVM];
#pragma unroll
                                for (int v = 1; v < NVM; ++v) acc[a][b][kc * NVM + v] += w1 * V[kc * NVM + v];
                            }
                        }
                }
                ++k;
                if (k == hi) k = lo;
            }
        }
        __syncthreads();
    }
    if (!valid || e == s) return;

    // ---- flush: one RED per (stencil point, component, real value) of this cell ----
    constexpr int OFF = NPT / 2;                        // footprint origin relative to (iz_u, ir_u)
#pragma unroll
    for (int b = 0; b < FP; ++b) {
        int ir = ir_u - OFF + b;
        const bool below = ir < 0;
        if (below) ir = -(1 + ir);
        if (ir > Nr - 1) ir = Nr - 1;
#pragma unroll
        for (int a = 0; a < FP; ++a) {
            int iz = iz_u - OFF + a;
            if (iz < 0) iz += Nz;
            if (iz > Nz - 1) iz -= Nz;
            const size_t o = (size_t)iz * Nr + ir;
#pragma unroll
            for (int kc = 0; kc < NC; ++kc) {
                const int comp = C0 + kc;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    double sgn = 1.;                    // flip of contributions folded from below the axis
                    if (below) {
                        sgn = (m & 1) ? -1. : 1.;
                        if (IS_J && comp < 2) sgn = -sgn;
                    }
                    double *p = (double *)(G.g[IS_J ? (3 * m + comp) : m] + o);
                    if (m == 0) {
                        const double v = acc[a][b][kc * NVM];
                        if (!DISPLACED || v != 0.) atomicAdd(p, sgn * v);
                    } else {
                        const double vr = acc[a][b][kc * NVM + 2 * m - 1], vi = acc[a][b][kc * NVM + 2 * m];
                        if (!DISPLACED || vr != 0. || vi != 0.) {
                            atomicAdd(p, sgn * vr);
                            atomicAdd(p + 1, sgn * vi);
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct DepArgs {
    B2DepPtrs P;
    const int32_t *idx32;
    double q, invdz, zmin, invdr, rmin;
    int Nz, Nr;
    B2DepGrids G;
    const int32_t *prefix;
    const double *r0, *rh;
};

template <int NM, bool IS_J, int NPT, int C0, int NC, bool PERMUTE, bool DISPLACED>
static void launch_dep(cudaStream_t s, const DepArgs &A) {
    const int ncells = A.Nz * (A.Nr + 1);
    unsigned grid = (unsigned)((ncells + DEP_TPB - 1) / DEP_TPB);
    k_deposit<NM, IS_J, NPT, C0, NC, PERMUTE, DISPLACED><<<grid, DEP_TPB, 0, s>>>(
        A.P, A.idx32, A.q, A.invdz, A.zmin, A.Nz, A.invdr, A.rmin, A.Nr, A.G, A.prefix, A.r0, A.rh, ncells);
}

template <int NM>
static int dispatch_dep(bool is_J, bool cubic, bool permute, bool displaced, cudaStream_t s, const DepArgs &A) {
    if (displaced) {
        if (is_J || cubic || permute) return b2_fail(-3, "displaced deposition: rho/linear only", __FILE__, __LINE__);
        launch_dep<NM, false, 2, 0, 1, false, true>(s, A);
        return 0;
    }
    if (!is_J) {
        if (!cubic) { if (permute) launch_dep<NM, false, 2, 0, 1, true, false>(s, A); else launch_dep<NM, false, 2, 0, 1, false, false>(s, A); }
        else { if (permute) launch_dep<NM, false, 4, 0, 1, true, false>(s, A); else launch_dep<NM, false, 4, 0, 1, false, false>(s, A); }
    } else {
        if (!cubic) { if (permute) launch_dep<NM, true, 2, 0, 3, true, false>(s, A); else launch_dep<NM, true, 2, 0, 3, false, false>(s, A); }
        else {   // 16 stencil points: one component per launch keeps the sums in registers
            DepArgs B = A;
            if (permute) {
                launch_dep<NM, true, 4, 0, 1, true, false>(s, A);
                for (int k = 0; k < 8; ++k) B.P.src[k] = A.P.dst[k];      // now sorted
            } else {
                launch_dep<NM, true, 4, 0, 1, false, false>(s, A);
            }
            launch_dep<NM, true, 4, 1, 1, false, false>(s, B);
            launch_dep<NM, true, 4, 2, 1, false, false>(s, B);
            g_b2_launches.fetch_add(2);
        }
    }
    return 0;
}

static int deposit_any(b2_ctx *ctx, bool is_J, int64_t n, const double *const *src8, double *const *dst8,
                       const int32_t *idx32, double q, double invdz, double zmin, int Nz, double invdr, double rmin,
                       int Nr, int Nm, void *const *grids, const int32_t *prefix, const double *r0, const double *rh,
                       int cubic, bool displaced, void *stream) {
    if (n <= 0) return 0;
    if (Nm < 1 || Nm > 4) return b2_fail(-3, "deposit: Nm must be in 1..4", __FILE__, __LINE__);
    DepArgs A;
    const bool permute = (dst8 != nullptr);
    for (int k = 0; k < 8; ++k) {
        A.P.src[k] = (k < (is_J || permute ? 8 : 4)) ? src8[k] : nullptr;
        A.P.dst[k] = permute ? dst8[k] : nullptr;
    }
    A.idx32 = idx32;
    A.q = q; A.invdz = invdz; A.zmin = zmin; A.invdr = invdr; A.rmin = rmin; A.Nz = Nz; A.Nr = Nr;
    const int ng = is_J ? 3 * Nm : Nm;
    for (int k = 0; k < ng; ++k) A.G.g[k] = (double2 *)grids[k];
    A.prefix = prefix; A.r0 = r0; A.rh = rh;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(is_J ? B2P_DEPOSIT_J : B2P_DEPOSIT_RHO, s);
    int rc;
    switch (Nm) {
        case 1: rc = dispatch_dep<1>(is_J, cubic != 0, permute, displaced, s, A); break;
        case 2: rc = dispatch_dep<2>(is_J, cubic != 0, permute, displaced, s, A); break;
        case 3: rc = dispatch_dep<3>(is_J, cubic != 0, permute, displaced, s, A); break;
        default: rc = dispatch_dep<4>(is_J, cubic != 0, permute, displaced, s, A); break;
    }
    if (rc) return rc;
    B2_LAUNCHED();
    return 0;
}

extern "C" {

int b2_deposit_rho(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                   double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                   void *const *grids, const int32_t *prefix, const double *r0, const double *rh, int cubic,
                   void *stream) {
    const double *src[8] = {x, y, z, w, nullptr, nullptr, nullptr, nullptr};
    return deposit_any(ctx, false, n, src, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, prefix,
                       r0, rh, cubic, false, stream);
}

int b2_deposit_J(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                 double q, const double *ux, const double *uy, const double *uz, const double *ig, double invdz,
                 double zmin, int Nz, double invdr, double rmin, int Nr, int Nm, void *const *grids,
                 const int32_t *prefix, const double *r0, const double *rh, int cubic, void *stream) {
    const double *src[8] = {x, y, z, w, ux, uy, uz, ig};
    return deposit_any(ctx, true, n, src, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, prefix,
                       r0, rh, cubic, false, stream);
}

int b2_deposit_permute(b2_ctx *ctx, int what, int64_t n, const double *const *src8, double *const *dst8,
                       double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                       void *const *grids, const int32_t *prefix, const double *r0, const double *rh, int cubic,
                       void *stream) {
    if (!ctx->last_idx32 || ctx->last_sort_n != n)
        return b2_fail(-4, "b2_deposit_permute: no matching b2_sort_cells result in this context", __FILE__, __LINE__);
    return deposit_any(ctx, what != 0, n, src8, dst8, ctx->last_idx32, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                       prefix, r0, rh, cubic, false, stream);
}

int b2_deposit_rho_displaced(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z,
                             const double *w, double q, double invdz, double zmin, int Nz, double invdr, double rmin,
                             int Nr, int Nm, void *const *grids, const int32_t *prefix, const double *r0,
                             const double *rh, void *stream) {
    const double *src[8] = {x, y, z, w, nullptr, nullptr, nullptr, nullptr};
    return deposit_any(ctx, false, n, src, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids, prefix,
                       r0, rh, 0, true, stream);
}

}  // extern "C"

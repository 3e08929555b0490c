// b2_deposit.cu -- charge / current deposition on cell-sorted particles.
//
// Replaces deposit_{rho,J}_gpu_{linear,cubic}[_one_mode] (fbpic/particles/deposition/
// cuda_methods.py:28,202,466,751 and cuda_methods_one_mode.py) -- one thread per cell, TPB 8,
// uncoalesced per-particle loads, one pass per azimuthal mode for Nm != 2.
//
// B200 design: a CTA owns TPB consecutive cells of the sorted order, i.e. ONE contiguous
// particle range [prefix_sum[c0-1], prefix_sum[c0+TPB-1]).  The range is streamed through
// shared memory in coalesced chunks (SoA, 8 B/attribute/particle); each thread then reduces the
// particles of its own cell into registers (no atomics, no shuffles: every particle of a cell
// scatters to the same stencil points, so the cell's sums are formed first), visiting them in a
// lane-rotated order so that the 32 lanes hit 32 different shared-memory banks.  All azimuthal
// modes are handled in the same pass.  One fp64 RED (red.global.add.f64) per cell, stencil point
// and real component flushes the sums: 2*ncomp*(2Nm-1)*npts^2 REDs per CELL instead of per
// particle (measured REDG.F64 rate on B200: ~590 G/s, profiles/r01_microbench.txt).
//
// Boundary folds follow fbpic/fields/numba_methods.py:410-461 / cuda_methods.py:167-177,670-691:
// z periodic; cells below the axis fold to -(1+ir) with the flip sign (-1)^m (rho, Jz) or
// -(-1)^m (Jr, Jt) (particle_shapes.py:33-36,76-79); cells beyond Nr-1 clamp to Nr-1.
#include "b2_common.cuh"

#define DEP_TPB 128
#define DEP_CHUNK 512

struct B2DepGrids {
    double2 *g[3 * B2_MAX_MODES];   // rho: [m] ; J: [m][Jr,Jt,Jz]
};

// NM modes; NATTR = 4 (rho: x,y,z,w) or 8 (J: + ux,uy,uz,inv_gamma); NPT = 2 linear, 4 cubic;
// components [C0, C0+NC) of (rho) or (Jr,Jt,Jz) are deposited by this launch.
template <int NM, bool IS_J, int NPT, int C0, int NC>
__global__ void __launch_bounds__(DEP_TPB)
k_deposit(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
          const double *__restrict__ w, const double *__restrict__ ux, const double *__restrict__ uy,
          const double *__restrict__ uz, const double *__restrict__ inv_gamma, double q,
          double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, B2DepGrids G,
          const int32_t *__restrict__ prefix_sum, const double *__restrict__ ruyten0,
          const double *__restrict__ ruyten_hi, int ncells) {
    constexpr int NATTR = IS_J ? 8 : 4;
    constexpr int NVM = 2 * NM - 1;          // real values per component: m=0 real, m>=1 complex
    constexpr int NV = NC * NVM;
    __shared__ double sm[NATTR][DEP_CHUNK];

    const int c0 = blockIdx.x * DEP_TPB;
    const int cell = c0 + threadIdx.x;
    const bool valid = cell < ncells;
    const int c_last = min(c0 + DEP_TPB, ncells) - 1;
    const int pb = (c0 == 0) ? 0 : prefix_sum[c0 - 1];
    const int pe = prefix_sum[c_last];
    if (pe == pb) return;                    // no particle in this CTA's cells (uniform exit)
    int s = 0, e = 0;
    if (valid) {
        s = (cell == 0) ? 0 : prefix_sum[cell - 1];
        e = prefix_sum[cell];
    }
    const int iz_u = valid ? cell / (Nr + 1) : 0;
    const int ir_u = valid ? cell - iz_u * (Nr + 1) : 0;
    const double beta0 = ruyten0[ir_u], beta_hi = ruyten_hi[ir_u];

    double acc[NPT][NPT][NV];
#pragma unroll
    for (int a = 0; a < NPT; ++a)
#pragma unroll
        for (int b = 0; b < NPT; ++b)
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[a][b][v] = 0.;

    const double *src[8] = {x, y, z, w, ux, uy, uz, inv_gamma};

    for (int q0 = pb; q0 < pe; q0 += DEP_CHUNK) {
        const int q1 = min(q0 + DEP_CHUNK, pe);
        // ---- stage the chunk (coalesced) ----
#pragma unroll
        for (int k = 0; k < NATTR; ++k)
            for (int i = threadIdx.x; i < q1 - q0; i += DEP_TPB) sm[k][i] = __ldg(src[k] + q0 + i);
        __syncthreads();
        // ---- reduce my cell's particles that lie in this chunk ----
        const int lo = max(s, q0), hi = min(e, q1);
        const int cnt = hi - lo;
        if (cnt > 0) {
            int k = lo + (int)(threadIdx.x & 31) % cnt;    // lane-rotated start: spreads smem banks
            for (int j = 0; j < cnt; ++j) {
                const int i = k - q0;
                const double xj = sm[0][i], yj = sm[1][i], zj = sm[2][i];
                const double wj = q * sm[3][i];
                const B2Cyl c = b2_cyl(xj, yj, zj, invdz, zmin, invdr, rmin);
                // per-particle values: comp-major, then [m=0 | Re m=1, Im m=1 | ...]
                double V[NV];
                if (!IS_J) {
                    V[0] = wj;
                } else {
                    const double f = wj * B2_C_LIGHT * sm[7][i];
                    const double uxj = sm[4][i], uyj = sm[5][i], uzj = sm[6][i];
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
                // shape factors (particle_shapes.py:17-80), flip applied once per cell at flush
                double sz[NPT], sr0[NPT], sr1[NPT];
                if (NPT == 2) {
                    sz[0] = ceil(c.z_cell) - c.z_cell;
                    sz[1] = 1. - sz[0];
                    const double u = c.r_cell - (ceil(c.r_cell) - 1.);
                    const double base = 1. - u, t = (1. - u) * u;
                    sr0[0] = base + beta0 * t;   sr0[1] = 1. - sr0[0];
                    sr1[0] = base + beta_hi * t; sr1[1] = 1. - sr1[0];
                } else {
                    const double uz_ = c.z_cell - (ceil(c.z_cell) - 2.) - 1.;
                    const double vz = 1. - uz_;
                    sz[0] = (1. / 6.) * (vz * vz * vz);
                    sz[1] = (1. / 6.) * (3. * (uz_ * uz_ * uz_) - 6. * (uz_ * uz_) + 4.);
                    sz[2] = (1. / 6.) * (3. * (vz * vz * vz) - 6. * (vz * vz) + 4.);
                    sz[3] = (1. / 6.) * (uz_ * uz_ * uz_);
                    const double u = c.r_cell - (ceil(c.r_cell) - 2.) - 1.;
                    const double v = 1. - u, t = (1. - u) * u;
                    const double s0 = (1. / 6.) * (v * v * v);
                    const double s1 = (1. / 6.) * (3. * (u * u * u) - 6. * (u * u) + 4.);
                    const double s2 = (1. / 6.) * (3. * (v * v * v) - 6. * (v * v) + 4.);
                    const double s3 = (1. / 6.) * (u * u * u);
                    sr0[0] = s0; sr0[1] = s1 + beta0 * t;   sr0[2] = s2 - beta0 * t;   sr0[3] = s3;
                    sr1[0] = s0; sr1[1] = s1 + beta_hi * t; sr1[2] = s2 - beta_hi * t; sr1[3] = s3;
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
                ++k;
                if (k == hi) k = lo;
            }
        }
        __syncthreads();
    }
    if (!valid || e == s) return;

    // ---- flush: one RED per (stencil point, component, real value) of this cell ----
#pragma unroll
    for (int b = 0; b < NPT; ++b) {
        int ir = ir_u - NPT / 2 + b;
        const bool below = ir < 0;
        if (below) ir = -(1 + ir);
        if (ir > Nr - 1) ir = Nr - 1;
#pragma unroll
        for (int a = 0; a < NPT; ++a) {
            int iz = iz_u - NPT / 2 + a;
            if (iz < 0) iz += Nz;
            if (iz > Nz - 1) iz -= Nz;
            const size_t o = (size_t)iz * Nr + ir;
#pragma unroll
            for (int kc = 0; kc < NC; ++kc) {
                const int comp = C0 + kc;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    // flip sign for contributions folded from below the axis
                    double sgn = 1.;
                    if (below) {
                        sgn = (m & 1) ? -1. : 1.;
                        if (IS_J && comp < 2) sgn = -sgn;
                    }
                    double *p = (double *)(G.g[IS_J ? (3 * m + comp) : m] + o);
                    if (m == 0) {
                        atomicAdd(p, sgn * acc[a][b][kc * NVM]);
                    } else {
                        atomicAdd(p, sgn * acc[a][b][kc * NVM + 2 * m - 1]);
                        atomicAdd(p + 1, sgn * acc[a][b][kc * NVM + 2 * m]);
                    }
                }
            }
        }
    }
}

template <int NM, bool IS_J, int NPT, int C0, int NC>
static void launch_dep(cudaStream_t s, int ncells, const double *x, const double *y, const double *z,
                       const double *w, const double *ux, const double *uy, const double *uz, const double *ig,
                       double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
                       const B2DepGrids &G, const int32_t *prefix, const double *r0, const double *rh) {
    unsigned grid = (unsigned)((ncells + DEP_TPB - 1) / DEP_TPB);
    k_deposit<NM, IS_J, NPT, C0, NC><<<grid, DEP_TPB, 0, s>>>(x, y, z, w, ux, uy, uz, ig, q, invdz, zmin, Nz, invdr,
                                                            rmin, Nr, G, prefix, r0, rh, ncells);
}

#define DEP_ARGS s, ncells, x, y, z, w, ux, uy, uz, ig, q, invdz, zmin, Nz, invdr, rmin, Nr, G, prefix, r0, rh

template <int NM>
static void dispatch_dep(bool is_J, bool cubic, cudaStream_t s, int ncells, const double *x, const double *y,
                         const double *z, const double *w, const double *ux, const double *uy, const double *uz,
                         const double *ig, double q, double invdz, double zmin, int Nz, double invdr, double rmin,
                         int Nr, const B2DepGrids &G, const int32_t *prefix, const double *r0, const double *rh) {
    if (!is_J) {
        if (!cubic) launch_dep<NM, false, 2, 0, 1>(DEP_ARGS);
        else launch_dep<NM, false, 4, 0, 1>(DEP_ARGS);
    } else {
        if (!cubic) launch_dep<NM, true, 2, 0, 3>(DEP_ARGS);
        else {   // 16 stencil points: one component per launch keeps the sums in registers
            launch_dep<NM, true, 4, 0, 1>(DEP_ARGS);
            launch_dep<NM, true, 4, 1, 1>(DEP_ARGS);
            launch_dep<NM, true, 4, 2, 1>(DEP_ARGS);
            g_b2_launches.fetch_add(2);
        }
    }
}

static int deposit_any(b2_ctx *ctx, bool is_J, int64_t n, const double *x, const double *y, const double *z,
                       const double *w, double q, const double *ux, const double *uy, const double *uz,
                       const double *ig, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
                       int Nm, void *const *grids, const int32_t *prefix, const double *r0, const double *rh,
                       int cubic, void *stream) {
    if (n <= 0) return 0;
    if (Nm < 1 || Nm > 4) return b2_fail(-3, "deposit: Nm must be in 1..4", __FILE__, __LINE__);
    B2DepGrids G;
    const int ng = is_J ? 3 * Nm : Nm;
    for (int k = 0; k < ng; ++k) G.g[k] = (double2 *)grids[k];
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(is_J ? B2P_DEPOSIT_J : B2P_DEPOSIT_RHO, s);
    const int ncells = Nz * (Nr + 1);
    switch (Nm) {
        case 1: dispatch_dep<1>(is_J, cubic != 0, DEP_ARGS); break;
        case 2: dispatch_dep<2>(is_J, cubic != 0, DEP_ARGS); break;
        case 3: dispatch_dep<3>(is_J, cubic != 0, DEP_ARGS); break;
        default: dispatch_dep<4>(is_J, cubic != 0, DEP_ARGS); break;
    }
    B2_LAUNCHED();
    return 0;
}

extern "C" {

int b2_deposit_rho(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                   double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                   void *const *grids, const int32_t *prefix, const double *r0, const double *rh, int cubic,
                   void *stream) {
    return deposit_any(ctx, false, n, x, y, z, w, q, nullptr, nullptr, nullptr, nullptr, invdz, zmin, Nz, invdr,
                       rmin, Nr, Nm, grids, prefix, r0, rh, cubic, stream);
}

int b2_deposit_J(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                 double q, const double *ux, const double *uy, const double *uz, const double *ig, double invdz,
                 double zmin, int Nz, double invdr, double rmin, int Nr, int Nm, void *const *grids,
                 const int32_t *prefix, const double *r0, const double *rh, int cubic, void *stream) {
    return deposit_any(ctx, true, n, x, y, z, w, q, ux, uy, uz, ig, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                       prefix, r0, rh, cubic, stream);
}

}  // extern "C"

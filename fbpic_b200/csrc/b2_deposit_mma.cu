// b2_deposit_mma.cu -- charge / current deposition as a tensor-core segmented reduction.
//
// Replaces deposit_{rho,J}_gpu_{linear,cubic}[_one_mode] (fbpic/particles/deposition/cuda_methods.py:
// 28,202,466,751; cuda_methods_one_mode.py: one thread per cell, TPB 8, serial loop over the cell's
// particles, one pass per mode) and, in the PERMUTE flavour, write_sorting_buffer /
// rearrange_particle_arrays (cuda_sorting.py:193-213, particles.py:510-555).
//
// Every particle contributes  grid[point] += S[point] * V[value]  with NPT^2 stencil points and
// NV = ncomp*(2Nm-1) real values; all particles of a run of equal cell key hit the same points, so the
// sum over a run is a small matrix product  C[point,value] = sum_p S[p,point] * V[p,value]  -- a GEMM
// with K = particles of the run.  The kernel therefore works in two fully parallel phases per CTA
// (256 consecutive particles of the array):
//   A  thread-per-particle: coalesced load (or gather through the sort permutation + coalesced
//      write-back of the sorted SoA when PERMUTE), cylindrical coordinates, shape factors with the
//      Ruyten correction, per-mode phases -> S (two weight classes: m=0 / m>=1 use different Ruyten
//      coefficients) and V "packets" in shared memory, plus the particle's cell key;
//   B  each warp takes 32 particles, walks the runs of equal key and reduces every run on the fp64
//      tensor cores (mma.sync m8n8k4: rows = points x class, columns = values, K = 4 particles per
//      step, out-of-run particles masked), then flushes the useful entries of the 8x8 accumulator
//      tiles with one red.global.add.f64 each.
// No thread owns a cell: load balance does not depend on the density profile, the kernel is correct
// for ANY particle order (runs just get shorter when the array is less sorted), so the same kernel
// serves sorted particles, particles that moved since their last sort (the second deposition of the
// PIC cycle needs no second sort) and unsorted ones.
//
// Boundary folds follow fbpic/fields/numba_methods.py:410-461 / cuda_methods.py:167-177,670-691:
// z periodic; points below the axis fold to -(1+ir) with the flip sign (-1)^m (rho, Jz) or -(-1)^m
// (Jr, Jt) (particle_shapes.py:33-36,76-79); points beyond Nr-1 clamp to Nr-1.
#include "b2_common.cuh"

#define DM_TPB 256          // particles per CTA = threads per CTA
#define DM_PITCH (DM_TPB + 4)
#define DM_IR_BITS 12       // packed run identity: ir_u <= Nr < 2^12, iz_u < 2^19

struct B2DmGrids {
    double2 *g[3 * B2_MAX_MODES];   // rho: [m] ; J: [m][Jr,Jt,Jz]
};
struct B2DmPtrs {
    const double *src[8];           // x,y,z,w,ux,uy,uz,inv_gamma
    double *dst[8];                 // PERMUTE: sorted destination arrays
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// PUSH (rho only): the second half position push of the PIC cycle (push_x_gpu, push/cuda_methods.py:
// 17-52) and the periodic wrap are applied first and written back; the charge is deposited at the new
// position -- one pass over the particle data instead of two (main.py:519,528).
struct B2DmPush {
    double chdt;            // c*dt of the push
    int wrap;               // wrap z into [wrap_zmin, wrap_zmax)
    double wrap_zmin, wrap_zmax;
};

template <int NM, bool IS_J, int NPT, bool PERMUTE, bool PUSH>
__global__ void __launch_bounds__(DM_TPB, (NPT == 2 && NM <= 2) ? 6 : 1)
k_deposit_mma(int64_t n, B2DmPtrs P, const int32_t *__restrict__ idx32, B2DmPush push, double q,
              double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, B2DmGrids G,
              const double *__restrict__ ruyten0, const double *__restrict__ ruyten_hi) {
    constexpr int NCOMP = IS_J ? 3 : 1;
    constexpr int NP = NPT * NPT;                      // stencil points
    constexpr int NROW = 2 * NP;                       // (point, weight class)
    constexpr int MT = NROW / 8;                       // 8-row MMA tiles: 1 (linear), 4 (cubic)
    constexpr int NVT = NCOMP * (2 * NM - 1);          // real values per particle
    constexpr int NT = (NVT + 7) / 8;                  // 8-column MMA tiles
    extern __shared__ __align__(16) unsigned char dm_smem[];
    // packets, transposed and padded: element (tile, row r, particle p) at (tile*8 + r)*DM_PITCH + p.
    // Phase A writes rows with consecutive threads (conflict-free); phase B reads fragment element
    // (r = lane/4, p = k0 + lane%4): DM_PITCH % 16 == 4 makes the 16 lanes of a half-warp hit 16
    // different 8-byte banks.
    double *sW = (double *)dm_smem;                    // [MT*8][DM_PITCH]
    double *sV = sW + MT * 8 * DM_PITCH;               // [NVT][DM_PITCH]: only the rows that carry a value
    // [DM_TPB] run identity = packed upper cell indices (iz_u << DM_IR_BITS | ir_u), -1: no particle
    int *sK = (int *)(sV + NVT * DM_PITCH);
    double **sG = (double **)(sK + DM_TPB);            // [NT*8] flush table: destination of an accumulator column
    int *sF = (int *)(sG + NT * 8);                    // [NT*8] its flags

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = blockIdx.x * (int64_t)DM_TPB;
    const int64_t i = q0 + tid;

    // ------------------------------------------------------------------ phase A
    {
        // (a thread beyond the end of the array works on a copy of the last particle with zero weight and marks
        //  its slot with key -1: no zero-filled packet registers, no divergent tail)
        const bool live = i < n;
        int key;
        double S[NROW];
        double V[NT * 8];
        {
            const int64_t ii = live ? i : n - 1;
            const size_t j = PERMUTE ? (size_t)__ldg(idx32 + ii) : (size_t)ii;
            double at[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) at[k] = (PERMUTE || IS_J || PUSH || k < 4) ? __ldg(P.src[k] + j) : 0.;
            if (PUSH) {
                const double f = push.chdt * at[7];
                at[0] += f * 1. * at[4];
                at[1] += f * 1. * at[5];
                at[2] += f * 1. * at[6];
                if (push.wrap) {
                    const double l_box = push.wrap_zmax - push.wrap_zmin;
                    while (at[2] >= push.wrap_zmax) at[2] -= l_box;
                    while (at[2] < push.wrap_zmin) at[2] += l_box;
                }
                if (!PERMUTE && live) {
                    P.dst[0][i] = at[0]; P.dst[1][i] = at[1]; P.dst[2][i] = at[2];
                }
            }
            if (PERMUTE && live) {
#pragma unroll
                for (int k = 0; k < 8; ++k) P.dst[k][i] = at[k];
            }
            if (!live) at[3] = 0.;
            const B2Cyl c = b2_cyl(at[0], at[1], at[2], invdz, zmin, invdr, rmin);
            int iru = (int)ceil(c.r_cell), izu = (int)ceil(c.z_cell);
            if (iru > Nr) iru = Nr;
            if (izu < 0) izu += Nz; else if (izu > Nz - 1) izu -= Nz;
            key = live ? ((izu << DM_IR_BITS) | iru) : -1;
            const double beta0 = __ldg(ruyten0 + iru), beta_hi = __ldg(ruyten_hi + iru);
            // shape factors (particle_shapes.py:17-80); the flip sign is applied at flush time
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
                sz[NPT - 2] = (1. / 6.) * (3. * (vz * vz * vz) - 6. * (vz * vz) + 4.);
                sz[NPT - 1] = (1. / 6.) * (uz_ * uz_ * uz_);
                const double u = c.r_cell - (ceil(c.r_cell) - 2.) - 1.;
                const double v = 1. - u, t = (1. - u) * u;
                const double s0 = (1. / 6.) * (v * v * v);
                const double s1 = (1. / 6.) * (3. * (u * u * u) - 6. * (u * u) + 4.);
                const double s2 = (1. / 6.) * (3. * (v * v * v) - 6. * (v * v) + 4.);
                const double s3 = (1. / 6.) * (u * u * u);
                sr0[0] = s0; sr0[1] = s1 + beta0 * t;   sr0[NPT - 2] = s2 - beta0 * t;   sr0[NPT - 1] = s3;
                sr1[0] = s0; sr1[1] = s1 + beta_hi * t; sr1[NPT - 2] = s2 - beta_hi * t; sr1[NPT - 1] = s3;
            }
#pragma unroll
            for (int a = 0; a < NPT; ++a)
#pragma unroll
                for (int b = 0; b < NPT; ++b) {
                    S[a * NPT + b] = sz[a] * sr0[b];
                    S[NP + a * NPT + b] = sz[a] * sr1[b];
                }
            // values: [m=0 of every component | (Re, Im) of m=1..NM-1 of every component]
            const double wj = q * at[3];
            double v0[NCOMP];
            if (!IS_J) {
                v0[0] = wj;
            } else {
                const double f = wj * B2_C_LIGHT * at[7];
                v0[0] = f * (c.cs * at[4] + c.sn * at[5]);
                v0[NCOMP > 1 ? 1 : 0] = f * (c.cs * at[5] - c.sn * at[4]);
                v0[NCOMP > 2 ? 2 : 0] = f * at[6];
            }
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) {
                V[k] = v0[k];
                double re = v0[k], im = 0.;
#pragma unroll
                for (int m = 1; m < NM; ++m) {
                    const double nre = c.cs * re - c.sn * im, nim = c.cs * im + c.sn * re;
                    re = nre; im = nim;
                    V[NCOMP + (k * (NM - 1) + (m - 1)) * 2] = re;
                    V[NCOMP + (k * (NM - 1) + (m - 1)) * 2 + 1] = im;
                }
            }
        }
        sK[tid] = key;
#pragma unroll
        for (int r = 0; r < MT * 8; ++r) sW[r * DM_PITCH + tid] = S[r];
#pragma unroll
        for (int v = 0; v < NVT; ++v) sV[v * DM_PITCH + tid] = V[v];
        if (tid < DM_PITCH - DM_TPB) {
#pragma unroll
            for (int v = 0; v < NVT; ++v) sV[v * DM_PITCH + DM_TPB + tid] = 0.;
        }
        // flush tables, one entry per accumulator column: grid (+ re / im part) the column's value goes to, and
        // flags -- 1: value of a mode m >= 1 (weight class 1), 2: changes sign when folded below the axis
        if (tid < NT * 8) {
            const int col = tid;
            int comp = col, m = 0, part = 0;
            if (col >= NCOMP) {
                const int d = col - NCOMP, km = d >> 1;
                part = d & 1;
                comp = km / (NM > 1 ? NM - 1 : 1);
                m = km - comp * (NM > 1 ? NM - 1 : 1) + 1;
            }
            bool neg = (m & 1) != 0;
            if (IS_J && comp < 2) neg = !neg;
            sG[col] = (col < NVT) ? ((double *)G.g[IS_J ? (3 * m + comp) : m] + part) : nullptr;
            sF[col] = (m > 0 ? 1 : 0) | (neg ? 2 : 0);
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase B
    const int g = lane >> 2, t = lane & 3;
    // lane constants of the flush: this lane holds C[row = mt*8+g][col = nt*8 + 2t + h]; bit (mt*NT + nt)*2 + h
    // of ok_mask: the entry carries a value of the row's weight class, of neg_mask: its sign flips below the axis
    unsigned ok_mask = 0, neg_mask = 0;
    int f_a[MT], f_b[MT];                     // stencil offsets of the row's point
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int row = mt * 8 + g;
        const int cls = row / NP, pt = row - cls * NP;
        f_a[mt] = pt / NPT - NPT / 2;
        f_b[mt] = pt % NPT - NPT / 2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = nt * 8 + 2 * t + h, fl = sF[col];
                if (col < NVT && (fl & 1) == cls) ok_mask |= 1u << ((mt * NT + nt) * 2 + h);
                if (fl & 2) neg_mask |= 1u << ((mt * NT + nt) * 2 + h);
            }
    }
    double *const *const sGl = sG + 2 * t;     // entry (nt, h) at sGl[nt * 8 + h]
    // fragment rows of this lane; value rows beyond NVT do not exist: their columns are never flushed, any
    // finite operand will do
    const double *a_row[MT], *b_row[NT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) a_row[mt] = sW + (mt * 8 + g) * DM_PITCH;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) b_row[nt] = sV + (nt * 8 + g < NVT ? nt * 8 + g : NVT - 1) * DM_PITCH;
    // A run that crosses a warp boundary belongs to the warp in front of the boundary, which follows it into the
    // next warp's slots, if the two pieces together are no longer than one warp's share (32 slots); the next warp
    // then starts behind it.  With ~16 particles per cell every second warp boundary used to cut a run in two (24
    // flushes per 256 particles instead of 17: C2 J 0.469 -> 0.437 ms); long runs (64 particles per cell) are
    // still cut at the boundaries, which keeps the warps of a CTA evenly loaded (owning whole runs cost C4 8 %).
    const int lo_w = warp * 32;                              // this warp's slots (CTA-local)
    const int my_key = sK[lo_w + lane];
    int pos = lo_w;
    if (warp > 0) {
        const int hkey = __shfl_sync(0xffffffffu, my_key, 0);
        const unsigned diff = __ballot_sync(0xffffffffu, my_key != hkey);
        const unsigned dprev = __ballot_sync(0xffffffffu, sK[lo_w - 32 + lane] != hkey);
        const int head = diff ? __ffs(diff) - 1 : 32;        // slots of my first run
        const int before = dprev ? __clz(dprev) : 32;        // slots of the same run in front of the boundary
        if (before > 0 && head + before <= 32) pos = lo_w + head;
    }
    while (pos < lo_w + 32) {
        const int key = __shfl_sync(0xffffffffu, my_key, pos - lo_w);
        // end of the run of equal keys starting at pos
        const unsigned diff = __ballot_sync(0xffffffffu, lo_w + lane >= pos && my_key != key);
        int end = diff ? lo_w + __ffs(diff) - 1 : lo_w + 32;
        if (!diff && warp < DM_TPB / 32 - 1) {
            const unsigned d2 = __ballot_sync(0xffffffffu, sK[lo_w + 32 + lane] != key);
            const int head = d2 ? __ffs(d2) - 1 : 32;
            if (head + (lo_w + 32 - pos) <= 32) end += head;
        }
        if (key >= 0) {
            double acc[MT][NT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.;
            // K = 4 particles per step; a particle outside the run is masked in the weights only (its values
            // are finite numbers of this CTA's packets, times zero)
            // (steps start at the run's first slot, not at a multiple of 4: ceil(len / 4) steps instead of 4.75 on
            //  average for 16 particles; the last step may read up to 3 slots past the run -- of the next run, or of
            //  the zero-filled pad columns behind slot 255)
            for (int k0 = pos; k0 < end; k0 += 4) {
                const int p = k0 + t;
                const bool in = p < end;
                double a[MT], b[NT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { const double v = a_row[mt][p]; a[mt] = in ? v : 0.; }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) b[nt] = b_row[nt][p];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
            }
            // ---- flush the useful entries with one RED each ----
            const int iz_u = key >> DM_IR_BITS, ir_u = key & ((1 << DM_IR_BITS) - 1);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                int iz = iz_u + f_a[mt];
                if (iz < 0) iz += Nz;
                if (iz > Nz - 1) iz -= Nz;
                int ir = ir_u + f_b[mt];
                const bool below = ir < 0;
                if (below) ir = ~ir;                          // -(1 + ir)
                if (ir > Nr - 1) ir = Nr - 1;
                const unsigned o = (unsigned)(iz * Nr + ir);  // < 2^31 (checked by the launcher)
                const unsigned flip = below ? neg_mask : 0u;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int bit = (mt * NT + nt) * 2 + h;
                        if ((ok_mask >> bit) & 1u) {
                            // sign flip = the flag bit moved onto the sign bit of the high word
                            const double v = acc[mt][nt][h];
                            const int hi = __double2hiint(v) ^ (int)((flip << (31 - bit)) & 0x80000000u);
                            // (the pointer comes out of shared memory: a plain atomicAdd would take the generic
                            //  path with its address-space checks)
                            asm volatile("red.global.add.f64 [%0], %1;"
                                         ::"l"((char *)sGl[nt * 8 + h] + (size_t)o * 16),
                                           "d"(__hiloint2double(hi, __double2loint(v))) : "memory");
                        }
                    }
            }
        }
        pos = end;
    }
}

// ---------------------------------------------------------------------------------------------
struct DmArgs {
    int64_t n;
    B2DmPtrs P;
    const int32_t *idx32;
    double q, invdz, zmin, invdr, rmin;
    int Nz, Nr;
    B2DmGrids G;
    const double *r0, *rh;
    B2DmPush push;
    bool do_push;
};

template <int NM, bool IS_J, int NPT, bool PERMUTE, bool PUSH>
static int launch_dm(cudaStream_t s, const DmArgs &A) {
    constexpr int NCOMP = IS_J ? 3 : 1;
    constexpr int MT = 2 * NPT * NPT / 8, NVT = NCOMP * (2 * NM - 1), NT = (NVT + 7) / 8;
    const size_t smem = sizeof(double) * DM_PITCH * (8 * MT + NVT) + sizeof(int) * DM_TPB + (sizeof(double *) + sizeof(int)) * NT * 8;
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        B2_CUDA(cudaFuncSetAttribute(k_deposit_mma<NM, IS_J, NPT, PERMUTE, PUSH>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    unsigned grid = (unsigned)((A.n + DM_TPB - 1) / DM_TPB);
    k_deposit_mma<NM, IS_J, NPT, PERMUTE, PUSH><<<grid, DM_TPB, smem, s>>>(
        A.n, A.P, A.idx32, A.push, A.q, A.invdz, A.zmin, A.Nz, A.invdr, A.rmin, A.Nr, A.G, A.r0, A.rh);
    return 0;
}

template <int NM>
static int dispatch_dm(bool is_J, bool cubic, bool permute, cudaStream_t s, const DmArgs &A) {
#define DM_GO(J, NPT) (permute ? launch_dm<NM, J, NPT, true, false>(s, A) : launch_dm<NM, J, NPT, false, false>(s, A))
    if (!is_J && A.do_push) return cubic ? launch_dm<NM, false, 4, false, true>(s, A) : launch_dm<NM, false, 2, false, true>(s, A);
    if (!is_J) return cubic ? DM_GO(false, 4) : DM_GO(false, 2);
    return cubic ? DM_GO(true, 4) : DM_GO(true, 2);
#undef DM_GO
}

int b2_deposit_mma(b2_ctx *ctx, bool is_J, int64_t n, const double *const *src8, double *const *dst8,
                   const int32_t *idx32, const B2DmPush *push, double q, double invdz, double zmin, int Nz, double invdr, double rmin,
                   int Nr, int Nm, void *const *grids, const double *r0, const double *rh, int cubic,
                   void *stream) {
    if (n <= 0) return 0;
    if (Nm < 1 || Nm > 4) return b2_fail(-3, "deposit: Nm must be in 1..4", __FILE__, __LINE__);
    if (Nr >= (1 << DM_IR_BITS) || Nz >= (1 << (31 - DM_IR_BITS)))
        return b2_fail(-3, "deposit: grid too large for the packed run key (Nr < 4096, Nz < 524288)", __FILE__, __LINE__);
    DmArgs A;
    const bool permute = (dst8 != nullptr) && (push == nullptr);
    A.n = n;
    A.do_push = (push != nullptr);
    if (push) A.push = *push; else { A.push.chdt = 0.; A.push.wrap = 0; A.push.wrap_zmin = A.push.wrap_zmax = 0.; }
    for (int k = 0; k < 8; ++k) {
        A.P.src[k] = (k < (is_J || permute || push ? 8 : 4)) ? src8[k] : nullptr;
        A.P.dst[k] = (permute || push) ? dst8[k] : nullptr;
    }
    A.idx32 = idx32;
    A.q = q; A.invdz = invdz; A.zmin = zmin; A.invdr = invdr; A.rmin = rmin; A.Nz = Nz; A.Nr = Nr;
    const int ng = is_J ? 3 * Nm : Nm;
    for (int k = 0; k < ng; ++k) A.G.g[k] = (double2 *)grids[k];
    A.r0 = r0; A.rh = rh;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(is_J ? B2P_DEPOSIT_J : B2P_DEPOSIT_RHO, s);
    int rc;
    switch (Nm) {
        case 1: rc = dispatch_dm<1>(is_J, cubic != 0, permute, s, A); break;
        case 2: rc = dispatch_dm<2>(is_J, cubic != 0, permute, s, A); break;
        case 3: rc = dispatch_dm<3>(is_J, cubic != 0, permute, s, A); break;
        default: rc = dispatch_dm<4>(is_J, cubic != 0, permute, s, A); break;
    }
    if (rc) return rc;
    B2_LAUNCHED();
    return 0;
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int b2_deposit_rho(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                   double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                   void *const *grids, const int32_t *prefix, const double *r0, const double *rh, int cubic,
                   void *stream) {
    (void)prefix;   // kept in the signature for the reference's call shape; the runs are found on the fly
    const double *src[8] = {x, y, z, w, nullptr, nullptr, nullptr, nullptr};
    return b2_deposit_mma(ctx, false, n, src, nullptr, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                          r0, rh, cubic, stream);
}

int b2_deposit_J(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *w,
                 double q, const double *ux, const double *uy, const double *uz, const double *ig, double invdz,
                 double zmin, int Nz, double invdr, double rmin, int Nr, int Nm, void *const *grids,
                 const int32_t *prefix, const double *r0, const double *rh, int cubic, void *stream) {
    (void)prefix;
    const double *src[8] = {x, y, z, w, ux, uy, uz, ig};
    return b2_deposit_mma(ctx, true, n, src, nullptr, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                          r0, rh, cubic, stream);
}

int b2_deposit_permute(b2_ctx *ctx, int what, int64_t n, const double *const *src8, double *const *dst8,
                       double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                       void *const *grids, const int32_t *prefix, const double *r0, const double *rh, int cubic,
                       void *stream) {
    if (n <= 0) return 0;     // empty species: nothing to permute or deposit
    // the cached permutation must be the one of THIS species: same length and same prefix-sum array as the last sort
    if (!ctx->last_idx32 || ctx->last_sort_n != n || (prefix && ctx->last_sort_prefix != (const void *)prefix))
        return b2_fail(-4, "b2_deposit_permute: no matching b2_sort_cells result in this context", __FILE__, __LINE__);
    return b2_deposit_mma(ctx, what != 0, n, src8, dst8, ctx->last_idx32, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm,
                          grids, r0, rh, cubic, stream);
}

int b2_deposit_rho_displaced(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z,
                             const double *w, double q, double invdz, double zmin, int Nz, double invdr, double rmin,
                             int Nr, int Nm, void *const *grids, const int32_t *prefix, const double *r0,
                             const double *rh, void *stream) {
    (void)prefix;
    const double *src[8] = {x, y, z, w, nullptr, nullptr, nullptr, nullptr};
    return b2_deposit_mma(ctx, false, n, src, nullptr, nullptr, nullptr, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                          r0, rh, 0, stream);
}

int b2_push_deposit_rho(b2_ctx *ctx, int64_t n, double *x, double *y, double *z, const double *w,
                        const double *ux, const double *uy, const double *uz, const double *ig, double dt,
                        int wrap, double wrap_zmin, double wrap_zmax, double q, double invdz, double zmin, int Nz,
                        double invdr, double rmin, int Nr, int Nm, void *const *grids, const double *r0,
                        const double *rh, int cubic, void *stream) {
    const double *src[8] = {x, y, z, w, ux, uy, uz, ig};
    double *dst[8] = {x, y, z, nullptr, nullptr, nullptr, nullptr, nullptr};
    B2DmPush push;
    push.chdt = B2_C_LIGHT * dt; push.wrap = wrap; push.wrap_zmin = wrap_zmin; push.wrap_zmax = wrap_zmax;
    return b2_deposit_mma(ctx, false, n, src, dst, nullptr, &push, q, invdz, zmin, Nz, invdr, rmin, Nr, Nm, grids,
                          r0, rh, cubic, stream);
}

}  // extern "C"

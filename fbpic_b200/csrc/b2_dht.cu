// b2_dht.cu -- discrete Hankel transform as an fp64 tensor-core GEMM (DMMA m8n8k4).
//
// Replaces DHT.transform / inverse_transform (fbpic/fields/spectral_transform/hankel.py:182-243):
// copy complex->real [2Nz,Nr], cuBLAS dgemm (m=Nr, n=2Nz, k=Nr), copy real->complex -- three full
// passes + a separate r,t<->p,m pass per vector field.  Here ONE kernel reads the complex [Nz,Nr]
// array(s) as they lie in HBM and writes the complex result:
//   out[iz, n] = rowscale[iz] * sum_j A[iz, j] * M[j, n]
// with the real and imaginary planes fed to the tensor cores as two independent real products that
// share the B fragments (the re/im pair of one element is a single 16-byte shared-memory load).
//   NPROD = 1 : A = c1*in1 (+ c2*in2)   -- scalar transforms, and the forward vector transform with
//               the (r,t)->(p,m) combination done on the fly while staging A (prologue fusion);
//   NPROD = 2 : P = in1 @ M1, Q = in2 @ M2, out1 = P+Q, out2 = i(P-Q) -- the inverse vector
//               transform with the (p,m)->(r,t) combination in the epilogue.
// fp64 only: sm_100a has no tcgen05 f64 kind; the f64 tensor path is mma.sync (DMMA.8x8x4 in SASS,
// measured 37.1 TFLOP/s peak on B200, profiles/r01_microbench.txt).  Bound: fp64 tensor pipe
// (AI = Nr/8 flop/B).  Several (array, matrix) jobs are batched in one launch (grid.z) so that the
// 148 SMs see >1 full wave of CTAs.  8 warps per CTA (2x4), warp tile 16 iz x 32 columns: 32 fp64
// accumulators per thread keep the kernel under 128 registers, so two CTAs share an SM and one CTA's
// prologue / epilogue overlaps the other's tensor work.
#include "b2_common.cuh"

#define DHT_BM 32          // iz rows per CTA tile (=64 real rows)
#define DHT_BN 128         // output columns per CTA tile (NPROD=1) ; 64 per product (NPROD=2)
#define DHT_BK 16          // K chunk
#define DHT_AP (DHT_BK + 4)   // A row pitch in complex elements: 20 -> conflict-free LDS.128
#define DHT_BP (DHT_BN + 4)   // B row pitch in doubles: 132 -> conflict-free LDS.64
#define DHT_THREADS 256
#define DHT_MAX_JOBS 16

struct DhtJob {
    const double2 *in1, *in2;   // NPROD=1: A = c1*in1 + c2*in2 (in2 may be null); NPROD=2: p, m
    double2 *out1, *out2;       // NPROD=1: out1 ; NPROD=2: r, t
    const double *M1, *M2;      // [Nr,Nr] row-major
    const double *rowscale;     // [Nz] or null
    double2 c1, c2;
};
struct DhtJobs {
    DhtJob j[DHT_MAX_JOBS];
};

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NPROD>
__global__ void __launch_bounds__(DHT_THREADS, 2)
k_dht(DhtJobs jobs, int Nz, int Nr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: A[2 buf][NPROD][BM][AP] double2 | B[2 buf][BK][BP] double
    double2 *sA = (double2 *)smem_raw;
    double *sB = (double *)(smem_raw + sizeof(double2) * 2 * NPROD * DHT_BM * DHT_AP);
    constexpr int A_ELEMS = NPROD * DHT_BM * DHT_AP;     // per buffer (double2)
    constexpr int B_ELEMS = DHT_BK * DHT_BP;             // per buffer (double)
    constexpr int NCOL = DHT_BN / NPROD;                 // output columns per product in this CTA
    constexpr int NI = NCOL / 4 / 8;                     // 8-col MMA blocks per warp per product (4 | 2)

    const DhtJob &J = jobs.j[blockIdx.z];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int iz0 = blockIdx.y * DHT_BM;
    const int n0 = blockIdx.x * NCOL;
    const bool mix = (NPROD == 1) && (J.in2 != nullptr);

    double acc[NPROD][2][NI][2][2];   // [prod][mi][ni][re|im][2]
#pragma unroll
    for (int p = 0; p < NPROD; ++p)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int n = 0; n < NI; ++n) {
                acc[p][i][n][0][0] = acc[p][i][n][0][1] = 0.;
                acc[p][i][n][1][0] = acc[p][i][n][1][1] = 0.;
            }

    const int KT = (Nr + DHT_BK - 1) / DHT_BK;
    // register staging of the next K chunk: load_tiles only ISSUES the global loads (raw values);
    // the (r,t)->(p,m) mixing arithmetic happens in store_tiles, after the MMAs of the current
    // chunk, so that no instruction depending on the loads sits in front of the tensor work.
    double2 ra[2][2];
    double rb[8];

    auto load_tiles = [&](int kt) {
        const int k0 = kt * DHT_BK;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int e = tid + DHT_THREADS * qq;
            const int izl = e >> 4, j = e & 15;
            const int iz = iz0 + izl, jj = k0 + j;
            const bool ok = (iz < Nz) && (jj < Nr);
            const size_t o = (size_t)iz * Nr + jj;
            ra[0][qq] = ok ? __ldg(J.in1 + o) : make_double2(0., 0.);
            if (NPROD == 2 || mix) ra[1][qq] = ok ? __ldg(J.in2 + o) : make_double2(0., 0.);
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
            const int e = tid + DHT_THREADS * qq;
            const int k = e / DHT_BN, c = e % DHT_BN;          // c: column within the 128-wide B tile
            const int kk = k0 + k;
            double v = 0.;
            if (NPROD == 1) {
                const int n = n0 + c;
                if (kk < Nr && n < Nr) v = __ldg(J.M1 + (size_t)kk * Nr + n);
            } else {
                const int p = c / NCOL, n = n0 + (c % NCOL);
                if (kk < Nr && n < Nr) v = __ldg((p == 0 ? J.M1 : J.M2) + (size_t)kk * Nr + n);
            }
            rb[qq] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int e = tid + DHT_THREADS * qq;
            const int izl = e >> 4, j = e & 15;
            if (NPROD == 1) {
                double2 v = ra[0][qq];
                if (mix) {
                    const double2 a = ra[0][qq], b = ra[1][qq];
                    v.x = J.c1.x * a.x - J.c1.y * a.y + J.c2.x * b.x - J.c2.y * b.y;
                    v.y = J.c1.x * a.y + J.c1.y * a.x + J.c2.x * b.y + J.c2.y * b.x;
                }
                sA[buf * A_ELEMS + izl * DHT_AP + j] = v;
            } else {
#pragma unroll
                for (int p = 0; p < NPROD; ++p)
                    sA[buf * A_ELEMS + (p * DHT_BM + izl) * DHT_AP + j] = ra[p][qq];
            }
        }
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
            const int e = tid + DHT_THREADS * qq;
            const int k = e / DHT_BN, c = e % DHT_BN;
            sB[buf * B_ELEMS + k * DHT_BP + c] = rb[qq];
        }
    };

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_tiles(kt + 1);
        const double2 *A = sA + buf * A_ELEMS;
        const double *B = sB + buf * B_ELEMS;
#pragma unroll
        for (int k4 = 0; k4 < DHT_BK / 4; ++k4) {
            double2 a[NPROD][2];
#pragma unroll
            for (int p = 0; p < NPROD; ++p)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    a[p][i] = A[(p * DHT_BM + wm * 16 + i * 8 + g) * DHT_AP + k4 * 4 + t];
#pragma unroll
            for (int p = 0; p < NPROD; ++p)
#pragma unroll
                for (int n = 0; n < NI; ++n) {
                    const double b = B[(k4 * 4 + t) * DHT_BP + p * NCOL + wn * (NCOL / 4) + n * 8 + g];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        dmma(acc[p][i][n][0][0], acc[p][i][n][0][1], a[p][i].x, b);
                        dmma(acc[p][i][n][1][0], acc[p][i][n][1][1], a[p][i].y, b);
                    }
                }
        }
        if (kt + 1 < KT) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: thread owns rows iz = iz0 + wm*16 + i*8 + g, columns n = .. + 2t, 2t+1 ----
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int iz = iz0 + wm * 16 + i * 8 + g;
        if (iz >= Nz) continue;
        const double rs = J.rowscale ? __ldg(J.rowscale + iz) : 1.;
#pragma unroll
        for (int n = 0; n < NI; ++n) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = n0 + wn * (NCOL / 4) + n * 8 + 2 * t + h;
                if (col >= Nr) continue;
                const size_t o = (size_t)iz * Nr + col;
                if (NPROD == 1) {
                    J.out1[o] = make_double2(rs * acc[0][i][n][0][h], rs * acc[0][i][n][1][h]);
                } else {
                    const double pr = acc[0][i][n][0][h], pi = acc[0][i][n][1][h];
                    const double qr = acc[NPROD - 1][i][n][0][h], qi = acc[NPROD - 1][i][n][1][h];
                    J.out1[o] = make_double2(rs * (pr + qr), rs * (pi + qi));        // r = P + Q
                    J.out2[o] = make_double2(rs * -(pi - qi), rs * (pr - qr));       // t = i (P - Q)
                }
            }
        }
    }
}

static double g_dht_flops = 0.;

template <int NPROD>
static int launch_dht(b2_ctx *ctx, const DhtJobs &jobs, int njobs, int Nz, int Nr, cudaStream_t s) {
    const size_t smem = sizeof(double2) * 2 * NPROD * DHT_BM * DHT_AP + sizeof(double) * 2 * DHT_BK * DHT_BP;
    static bool attr_set[3] = {false, false, false};
    if (!attr_set[NPROD]) {
        B2_CUDA(cudaFuncSetAttribute(k_dht<NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[NPROD] = true;
    }
    const int ncol = DHT_BN / NPROD;
    B2Prof prof_(B2P_DHT, s);
    g_dht_flops += (double)njobs * NPROD * 4. * Nz * (double)Nr * Nr;   // real flops issued (2 planes x 2*Nz*Nr*Nr)
    dim3 grid((Nr + ncol - 1) / ncol, (Nz + DHT_BM - 1) / DHT_BM, njobs);
    k_dht<NPROD><<<grid, DHT_THREADS, smem, s>>>(jobs, Nz, Nr);
    B2_LAUNCHED();
    (void)ctx;
    return 0;
}

// b2_dht_tma.cu: TMA-fed persistent kernel, one launch for a mixed job list; returns 1 when that path is
// unavailable (B2_DHT_IMPL=legacy, a driver without cuTensorMapEncodeTiled, more than DHT_MAX_JOBS jobs) and the
// LDG-staged k_dht above runs instead
int b2_dht_tma_run(b2_ctx *ctx, const b2_dht_job *jobs, int njobs, int Nz, int Nr, cudaStream_t s);

static int run_tma(b2_ctx *ctx, const b2_dht_job *jobs, int njobs, int Nz, int Nr, cudaStream_t s) {
    B2Prof prof_(B2P_DHT, s);
    int rc = b2_dht_tma_run(ctx, jobs, njobs, Nz, Nr, s);
    if (rc == 0)
        for (int k = 0; k < njobs; ++k)
            g_dht_flops += (jobs[k].kind == B2_DHT_SCALAR ? 1. : 2.) * 4. * Nz * (double)Nr * Nr;
    return rc;
}

extern "C" {

double b2_dht_flops(void) { return g_dht_flops; }

// Batched transforms: the whole list (all modes / components, any mix of kinds) in one launch.
int b2_dht_batch(b2_ctx *ctx, int njobs, const b2_dht_job *jobs, int Nz, int Nr, void *stream) {
    {   // TMA path: the whole list in one launch
        int rc = run_tma(ctx, jobs, njobs, Nz, Nr, b2_stream_of(ctx, stream));
        if (rc != 1) return rc;          // done (0) or failed (<0 / CUDA code); 1 = path unavailable
    }
    // LDG-staged kernels: one launch per flavour, flushed whenever a list is full
    DhtJobs j1, j2;
    int n1 = 0, n2 = 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    for (int k = 0; k < njobs; ++k) {
        const b2_dht_job &u = jobs[k];
        if (u.kind == B2_DHT_SCALAR) {
            if (n1 >= DHT_MAX_JOBS) { int rc = launch_dht<1>(ctx, j1, n1, Nz, Nr, s); if (rc) return rc; n1 = 0; }
            DhtJob &j = j1.j[n1++];
            j.in1 = (const double2 *)u.in1; j.in2 = nullptr; j.out1 = (double2 *)u.out1; j.out2 = nullptr;
            j.M1 = u.M1; j.M2 = nullptr; j.rowscale = u.rowscale;
            j.c1 = make_double2(1., 0.); j.c2 = make_double2(0., 0.);
        } else if (u.kind == B2_DHT_RT_TO_PM) {
            if (n1 + 2 > DHT_MAX_JOBS) { int rc = launch_dht<1>(ctx, j1, n1, Nz, Nr, s); if (rc) return rc; n1 = 0; }
            for (int h = 0; h < 2; ++h) {
                DhtJob &j = j1.j[n1++];
                j.in1 = (const double2 *)u.in1; j.in2 = (const double2 *)u.in2;
                j.out1 = (double2 *)(h == 0 ? u.out1 : u.out2); j.out2 = nullptr;
                j.M1 = (h == 0 ? u.M1 : u.M2); j.M2 = nullptr; j.rowscale = u.rowscale;
                j.c1 = make_double2(0.5, 0.); j.c2 = make_double2(0., h == 0 ? -0.5 : 0.5);
            }
        } else if (u.kind == B2_DHT_PM_TO_RT) {
            if (n2 >= DHT_MAX_JOBS) { int rc = launch_dht<2>(ctx, j2, n2, Nz, Nr, s); if (rc) return rc; n2 = 0; }
            DhtJob &j = j2.j[n2++];
            j.in1 = (const double2 *)u.in1; j.in2 = (const double2 *)u.in2;
            j.out1 = (double2 *)u.out1; j.out2 = (double2 *)u.out2;
            j.M1 = u.M1; j.M2 = u.M2; j.rowscale = u.rowscale;
            j.c1 = make_double2(1., 0.); j.c2 = make_double2(1., 0.);
        } else {
            return b2_fail(-3, "b2_dht_batch: unknown job kind", __FILE__, __LINE__);
        }
    }
    if (n1) { int rc = launch_dht<1>(ctx, j1, n1, Nz, Nr, s); if (rc) return rc; }
    if (n2) { int rc = launch_dht<2>(ctx, j2, n2, Nz, Nr, s); if (rc) return rc; }
    return 0;
}

int b2_dht(b2_ctx *ctx, const void *in, void *out, const double *M, const double *rowscale, int Nz, int Nr,
           void *stream) {
    {
        b2_dht_job u = {in, nullptr, out, nullptr, M, nullptr, rowscale, B2_DHT_SCALAR};
        int rc = run_tma(ctx, &u, 1, Nz, Nr, b2_stream_of(ctx, stream));
        if (rc != 1) return rc;
    }
    DhtJobs jobs;
    DhtJob &j = jobs.j[0];
    j.in1 = (const double2 *)in; j.in2 = nullptr; j.out1 = (double2 *)out; j.out2 = nullptr;
    j.M1 = M; j.M2 = nullptr; j.rowscale = rowscale;
    j.c1 = make_double2(1., 0.); j.c2 = make_double2(0., 0.);
    return launch_dht<1>(ctx, jobs, 1, Nz, Nr, b2_stream_of(ctx, stream));
}

int b2_dht_rt_to_pm(b2_ctx *ctx, const void *r, const void *t, void *out_p, void *out_m, const double *Mp,
                    const double *Mm, const double *rowscale, int Nz, int Nr, void *stream) {
    {
        b2_dht_job u = {r, t, out_p, out_m, Mp, Mm, rowscale, B2_DHT_RT_TO_PM};
        int rc = run_tma(ctx, &u, 1, Nz, Nr, b2_stream_of(ctx, stream));
        if (rc != 1) return rc;
    }
    DhtJobs jobs;
    for (int k = 0; k < 2; ++k) {
        DhtJob &j = jobs.j[k];
        j.in1 = (const double2 *)r; j.in2 = (const double2 *)t;
        j.out1 = (double2 *)(k == 0 ? out_p : out_m); j.out2 = nullptr;
        j.M1 = (k == 0 ? Mp : Mm); j.M2 = nullptr; j.rowscale = rowscale;
        j.c1 = make_double2(0.5, 0.);
        j.c2 = make_double2(0., k == 0 ? -0.5 : 0.5);      // p = (r - i t)/2 ; m = (r + i t)/2
    }
    return launch_dht<1>(ctx, jobs, 2, Nz, Nr, b2_stream_of(ctx, stream));
}

int b2_dht_pm_to_rt(b2_ctx *ctx, const void *p, const void *m, void *out_r, void *out_t, const double *iMp,
                    const double *iMm, const double *rowscale, int Nz, int Nr, void *stream) {
    {
        b2_dht_job u = {p, m, out_r, out_t, iMp, iMm, rowscale, B2_DHT_PM_TO_RT};
        int rc = run_tma(ctx, &u, 1, Nz, Nr, b2_stream_of(ctx, stream));
        if (rc != 1) return rc;
    }
    DhtJobs jobs;
    DhtJob &j = jobs.j[0];
    j.in1 = (const double2 *)p; j.in2 = (const double2 *)m;
    j.out1 = (double2 *)out_r; j.out2 = (double2 *)out_t;
    j.M1 = iMp; j.M2 = iMm; j.rowscale = rowscale;
    j.c1 = make_double2(1., 0.); j.c2 = make_double2(1., 0.);
    return launch_dht<2>(ctx, jobs, 1, Nz, Nr, b2_stream_of(ctx, stream));
}

}  // extern "C"

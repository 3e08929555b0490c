// b2_fft.cu -- z-FFT of the [Nz][Nr] field arrays as a two-pass (four-step) transform, every access coalesced
// along r.
//
// Replaces FFT.transform / inverse_transform (fbpic/fields/spectral_transform/fourier.py:104-168: cuFFT Z2Z plan
// (Nz, batch Nr) + two transpose kernels + a scale pass).  cuFFT's strided plan on the native layout (b2_fields.cu)
// needs no transposes, but runs at 1.8 TB/s for Nz = 4096 and at 0.8 TB/s as soon as Nz is not a power of two
// (profiles/r02_fft_sizes.txt: 43 us for the 4224 cells of a z-slab with its guard cells, 18.6 us for 4096) --
// and every local grid with guard, damping or injection cells has such a length.  Here Nz = n1 * n2:
//   pass 1, line j2 < n2 : n1-point DFT over the rows j1*n2 + j2, times the twiddle w_N^(j2*k1) -> scratch row k1*n2 + j2
//   pass 2, line k1 < n1 : n2-point DFT over the scratch rows k1*n2 + j2 (contiguous)         -> out row k1 + n1*k2
// A CTA owns one line x 16 columns (256-byte row segments); its n = RA*RB points are transformed in two register
// stages (radix RA: butterflies for 2, 4, 8, direct sums for the odd and composite radices a grid length with
// guard cells brings: 3, 5, 6, 7, 9, 10, 11, 12, 13, 17, 23; the sums of a radix above 13 are spread over several
// threads) with one exchange through shared memory.  The scratch array
// of a group of arrays stays in the 126 MB L2 between the passes, so DRAM sees one read and one write per array.
// Sizes without an (n1, n2) plan fall back to cuFFT.
#include "b2_common.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#define FF_C 16                 // columns per CTA
#define FF_GROUP 8              // max arrays per launch pair (B2_FFT_GROUP, default 6: the scratch of a group stays L2-resident)

struct FfArgs {
    const double2 *in[FF_GROUP];
    double2 *out[FF_GROUP];
    const double2 *Wn;          // w_n^k, k < n (forward sign)
    const double2 *WN;          // w_N^k, k < N, or null (no inter-pass twiddle)
    long long line_stride_in, elem_stride_in, line_stride_out, elem_stride_out;   // in rows
    int Nr, inverse;
    double scale;
};

__device__ __forceinline__ double2 ff_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 ff_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 ff_mul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiplication by -i (forward, s = +1) or +i (inverse, s = -1)
__device__ __forceinline__ double2 ff_rot(double2 a, double s) { return make_double2(s * a.y, -s * a.x); }

// R-point DFT in registers, natural order in and out.  w[j] = w_R^j with the direction's sign already applied.
// Output q is handed to `emit(q, value)` as soon as it is formed.  The direct sums of the odd and composite radices
// use the conjugate symmetry of the twiddles, w^{(R-j)q} = conj(w^{jq}):  with a_j = x_j + x_{R-j} and
// b_j = x_j - x_{R-j},  X[q] = P + iQ and X[R-q] = P - iQ  where  P = x_0 + sum_j Re(w^{jq}) a_j (+ (-1)^q x_{R/2}),
// Q = sum_j Im(w^{jq}) b_j -- 4 real multiply-adds per (pair of outputs, pair of inputs) instead of 16: the
// 11-point sums of the 4224-cell slab transform cost 130 operations instead of 440.
template <int R> struct FfDft {
    template <class Emit>
    static __device__ __forceinline__ void run(double2 (&x)[R], const double2 (&w)[R], double, Emit emit) {
        constexpr int H = (R - 1) / 2;
        constexpr bool EVEN = (R % 2 == 0);
        double2 a[H], b[H];
#pragma unroll
        for (int j = 1; j <= H; ++j) { a[j - 1] = ff_add(x[j], x[R - j]); b[j - 1] = ff_sub(x[j], x[R - j]); }
        {
            double2 acc = x[0];
#pragma unroll
            for (int j = 0; j < H; ++j) acc = ff_add(acc, a[j]);
            if (EVEN) acc = ff_add(acc, x[R / 2]);
            emit(0, acc);
        }
        if (EVEN) {     // q = R/2: every twiddle is +-1
            double2 acc = ((R / 2) & 1) ? ff_sub(x[0], x[R / 2]) : ff_add(x[0], x[R / 2]);
#pragma unroll
            for (int j = 1; j <= H; ++j) acc = (j & 1) ? ff_sub(acc, a[j - 1]) : ff_add(acc, a[j - 1]);
            emit(R / 2, acc);
        }
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            double2 P = x[0], Q = make_double2(0., 0.);
            if (EVEN) P = (q & 1) ? ff_sub(P, x[R / 2]) : ff_add(P, x[R / 2]);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const double2 t = w[(j * q) % R];
                P.x += t.x * a[j - 1].x; P.y += t.x * a[j - 1].y;
                Q.x += t.y * b[j - 1].x; Q.y += t.y * b[j - 1].y;
            }
            emit(q, make_double2(P.x - Q.y, P.y + Q.x));
            emit(R - q, make_double2(P.x + Q.y, P.y - Q.x));
        }
    }
};
template <> struct FfDft<2> {
    template <class Emit>
    static __device__ __forceinline__ void run(double2 (&x)[2], const double2 (&)[2], double, Emit emit) {
        emit(0, ff_add(x[0], x[1])); emit(1, ff_sub(x[0], x[1]));
    }
};
template <> struct FfDft<4> {
    static __device__ __forceinline__ void inplace(double2 (&x)[4], double s) {
        const double2 t0 = ff_add(x[0], x[2]), t1 = ff_sub(x[0], x[2]);
        const double2 t2 = ff_add(x[1], x[3]), t3 = ff_rot(ff_sub(x[1], x[3]), s);
        x[0] = ff_add(t0, t2); x[2] = ff_sub(t0, t2);
        x[1] = ff_add(t1, t3); x[3] = ff_sub(t1, t3);
    }
    template <class Emit>
    static __device__ __forceinline__ void run(double2 (&x)[4], const double2 (&)[4], double s, Emit emit) {
        inplace(x, s);
#pragma unroll
        for (int q = 0; q < 4; ++q) emit(q, x[q]);
    }
};
template <> struct FfDft<8> {
    template <class Emit>
    static __device__ __forceinline__ void run(double2 (&x)[8], const double2 (&)[8], double s, Emit emit) {
        const double c = 0.70710678118654752440;
        double2 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = ff_add(x[k], x[k + 4]); b[k] = ff_sub(x[k], x[k + 4]); }
        // b[k] *= w_8^k : 1, (1 - i s)/sqrt2, -i s, (-1 - i s)/sqrt2
        b[1] = make_double2(c * (b[1].x + s * b[1].y), c * (b[1].y - s * b[1].x));
        b[2] = ff_rot(b[2], s);
        b[3] = make_double2(c * (-b[3].x + s * b[3].y), c * (-b[3].y - s * b[3].x));
        FfDft<4>::inplace(a, s);
        FfDft<4>::inplace(b, s);
#pragma unroll
        for (int m = 0; m < 4; ++m) { emit(2 * m, a[m]); emit(2 * m + 1, b[m]); }
    }
};

// P2 > 1 (pairs with a big direct radix RB, written (small, big)): P2 threads share the RB-point sums of one
// (qa, column) -- thread `part` forms the outputs qb = part, part + P2, ... -- so that the second stage has as many
// busy threads as the first (RA threads with 23^2 complex products each left the rest of the CTA idle).
// resident CTAs the register allocation has to leave room for (the symmetric sums give ptxas many independent
// outputs to keep in flight: unbounded, the 11-point pass takes 132 registers and 2 CTAs per SM instead of 3)
constexpr int ff_pass_min_ctas(int rmax, int p2) {
    return p2 > 1 ? 2 : 65536 / ((rmax == 7 ? 80 : rmax <= 8 ? 64 : rmax <= 11 ? 104 : 128) * rmax * FF_C);
}
template <int RA, int RB, int P2>
__global__ void __launch_bounds__((RA * P2 > RB ? RA * P2 : (RA > RB ? RA : RB)) * FF_C,
                                  ff_pass_min_ctas(RA > RB ? RA : RB, P2))
k_fft_pass(const FfArgs A) {
    constexpr int n = RA * RB;
    __shared__ double2 sW[n], sT[n];
    __shared__ double2 sS[n][FF_C];
    const int tid = threadIdx.x, u = tid / FF_C, col = tid % FF_C;
    // column tile fastest: the CTAs in flight together cover whole 4 KB rows (adjacent 256-byte segments of the
    // same DRAM pages) instead of the same segment of 64 rows that lie n2 rows apart
    const long long line = blockIdx.y;
    const int c = blockIdx.x * FF_C + col;
    const double2 *in = A.in[blockIdx.z];
    double2 *out = A.out[blockIdx.z];
    const double s = A.inverse ? -1. : 1.;
    const bool live = c < A.Nr;
    // the element loads go out first; the tables -- w_n^k and, in the first pass, this line's n inter-pass twiddles
    // w_N^(line*q) -- are fetched behind them and parked in shared memory (ncu: with a __ldg at the point of use,
    // the wait for the inter-pass twiddle was as long as the wait for the elements themselves)
    const bool act1 = u < RB && live;
    double2 x[RA];
    if (act1) {
#pragma unroll
        for (int ea = 0; ea < RA; ++ea)
            x[ea] = __ldg(in + (size_t)(line * A.line_stride_in + (long long)(ea * RB + u) * A.elem_stride_in) * A.Nr + c);
    }
    for (int k = tid; k < n; k += blockDim.x) {
        double2 w = __ldg(A.Wn + k);
        if (A.inverse) w.y = -w.y;
        sW[k] = w;
        if (A.WN) {
            double2 t = __ldg(A.WN + line * k);
            if (A.inverse) t.y = -t.y;
            sT[k] = t;
        }
    }
    __syncthreads();
    // ---- stage 1: thread (eb = u, column): RA-point DFT over the elements ea*RB + eb, twiddle w_n^(eb*qa)
    if (act1) {
        double2 w[RA];
#pragma unroll
        for (int j = 0; j < RA; ++j) w[j] = sW[j * RB];
        FfDft<RA>::run(x, w, s, [&](int qa, double2 v) { sS[qa * RB + u][col] = ff_mul(v, sW[u * qa]); });
    }
    __syncthreads();
    // ---- stage 2: thread (qa = u, column): RB-point DFT over eb, output element q = qa + RA*qb
    if constexpr (P2 > 1) {
        if (u < RA * P2 && live) {
            const int qa = u % RA, part = u / RA;
            // the same symmetric sums as FfDft<R>, the (RB + 1) / 2 units -- output 0, then the output pairs
            // (q, RB - q) -- dealt out to the P2 threads
            constexpr int H = (RB - 1) / 2;
            static_assert(RB % 2 == 1, "the split second stage is written for odd radices");
            // (every y_j is read once and feeds all the units of the thread: the sums stay in 4 registers per unit
            //  -- held in registers, a_j and b_j cost 88 registers for radix 23 and one CTA per SM)
            constexpr int NU = (H + 1 + P2 - 1) / P2;         // units per thread
            const double2 y0 = sS[qa * RB][col];
            double2 P[NU], Q[NU];
            int idx[NU];
#pragma unroll
            for (int k = 0; k < NU; ++k) { P[k] = y0; Q[k] = make_double2(0., 0.); idx[k] = 0; }
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const double2 yj = sS[qa * RB + j][col], yr = sS[qa * RB + RB - j][col];
                const double2 a = ff_add(yj, yr), b = ff_sub(yj, yr);
#pragma unroll
                for (int k = 0; k < NU; ++k) {
                    const int unit = part + k * P2;           // (a unit beyond H computes a copy of output 0: unused)
                    idx[k] += (unit <= H) ? unit : 0;         // (j * unit) mod RB, advanced with j
                    if (idx[k] >= RB) idx[k] -= RB;
                    const double2 t = sW[idx[k] * RA];        // w_RB^(j*unit)
                    P[k].x += t.x * a.x; P[k].y += t.x * a.y;
                    Q[k].x += t.y * b.x; Q[k].y += t.y * b.y;
                }
            }
            auto put = [&](int qb, double2 acc) {
                const int q = qa + RA * qb;
                if (A.WN) acc = ff_mul(acc, sT[q]);
                acc.x *= A.scale; acc.y *= A.scale;
                out[(size_t)(line * A.line_stride_out + (long long)q * A.elem_stride_out) * A.Nr + c] = acc;
            };
#pragma unroll
            for (int k = 0; k < NU; ++k) {
                const int unit = part + k * P2;
                if (unit > H) break;
                put(unit, make_double2(P[k].x - Q[k].y, P[k].y + Q[k].x));
                if (unit > 0) put(RB - unit, make_double2(P[k].x + Q[k].y, P[k].y - Q[k].x));
            }
        }
        return;
    }
    if (u < RA && live) {
        double2 y[RB], w[RB];
#pragma unroll
        for (int eb = 0; eb < RB; ++eb) y[eb] = sS[u * RB + eb][col];
#pragma unroll
        for (int j = 0; j < RB; ++j) w[j] = sW[j * RA];
        FfDft<RB>::run(y, w, s, [&](int qb, double2 v) {
            const int q = u + RA * qb;
            if (A.WN) v = ff_mul(v, sT[q]);
            v.x *= A.scale; v.y *= A.scale;
            out[(size_t)(line * A.line_stride_out + (long long)q * A.elem_stride_out) * A.Nr + c] = v;
        });
    }
}

// ---- host side: plans ---------------------------------------------------------------------------------
typedef void (*ff_kernel_t)(const FfArgs);
struct FfPair { int ra, rb, p2; ff_kernel_t fn; };
#define FF_PAIR(a, b) {a, b, ((b) > 13 ? (b) / (a) : 1), k_fft_pass<a, b, ((b) > 13 ? (b) / (a) : 1)>}
static const FfPair g_ff_pairs[] = {
    FF_PAIR(2, 2), FF_PAIR(2, 4), FF_PAIR(4, 4), FF_PAIR(4, 8), FF_PAIR(8, 8),        // 4, 8, 16, 32, 64
    FF_PAIR(3, 4), FF_PAIR(4, 5), FF_PAIR(4, 6), FF_PAIR(4, 7), FF_PAIR(4, 9), FF_PAIR(4, 10),   // 12 .. 40
    FF_PAIR(3, 8), FF_PAIR(5, 8), FF_PAIR(6, 8), FF_PAIR(7, 8), FF_PAIR(8, 9), FF_PAIR(8, 10),   // 24 .. 80
    FF_PAIR(8, 11), FF_PAIR(8, 12), FF_PAIR(8, 13),                                   // 88, 96, 104
    FF_PAIR(3, 11), FF_PAIR(5, 7), FF_PAIR(6, 6), FF_PAIR(5, 9), FF_PAIR(5, 10), FF_PAIR(6, 9),  // 33 .. 54
    FF_PAIR(6, 10), FF_PAIR(5, 13), FF_PAIR(6, 11), FF_PAIR(3, 23), FF_PAIR(7, 10), FF_PAIR(6, 12),   // 60 .. 72
    FF_PAIR(2, 17), FF_PAIR(4, 17), FF_PAIR(2, 23), FF_PAIR(4, 23),                   // 34, 68, 46, 92
    FF_PAIR(7, 11), FF_PAIR(6, 13), FF_PAIR(9, 9), FF_PAIR(7, 12), FF_PAIR(9, 10), FF_PAIR(7, 13),    // 77 .. 91
    FF_PAIR(9, 11), FF_PAIR(10, 10), FF_PAIR(9, 12), FF_PAIR(10, 11), FF_PAIR(10, 12), FF_PAIR(11, 11),
    FF_PAIR(10, 13), FF_PAIR(11, 12), FF_PAIR(12, 12),
};
static const int g_ff_npairs = (int)(sizeof(g_ff_pairs) / sizeof(g_ff_pairs[0]));

struct FfPlan {
    int N, p1, p2;              // indices into g_ff_pairs: n1 = size(p1) (pass 1), n2 = size(p2) (pass 2)
    double2 *Wn1, *Wn2, *WN;    // device tables
};
static std::map<int, FfPlan> g_ff_plans;        // by N; p1 < 0: no plan (cuFFT)
static std::mutex g_ff_mutex;

static double ff_cost(int r) { return (r == 2 || r == 4 || r == 8) ? 1.5 * std::log2((double)r) : (double)r; }

static int ff_upload_table(int n, double2 **out) {
    std::vector<double2> h(n);
    for (int k = 0; k < n; ++k) {
        // exact octant symmetries are not needed at the 1e-13 tolerance; long double keeps the table at 1 ulp
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        h[k] = make_double2((double)cosl(a), (double)sinl(a));
    }
    B2_CUDA(cudaMalloc(out, sizeof(double2) * n));
    B2_CUDA(cudaMemcpy(*out, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    return 0;
}

static int ff_get_plan(int N, const FfPlan **plan) {
    auto it = g_ff_plans.find(N);
    if (it == g_ff_plans.end()) {
        FfPlan P;
        P.N = N; P.p1 = P.p2 = -1; P.Wn1 = P.Wn2 = P.WN = nullptr;
        double best = 1e300;
        for (int a = 0; a < g_ff_npairs; ++a)
            for (int b = 0; b < g_ff_npairs; ++b) {
                const int n1 = g_ff_pairs[a].ra * g_ff_pairs[a].rb, n2 = g_ff_pairs[b].ra * g_ff_pairs[b].rb;
                if ((long long)n1 * n2 != N) continue;
                const double cst = ff_cost(g_ff_pairs[a].ra) + ff_cost(g_ff_pairs[a].rb) + ff_cost(g_ff_pairs[b].ra)
                                   + ff_cost(g_ff_pairs[b].rb);
                if (cst < best) { best = cst; P.p1 = a; P.p2 = b; }
            }
        if (P.p1 >= 0) {
            const int n1 = g_ff_pairs[P.p1].ra * g_ff_pairs[P.p1].rb, n2 = g_ff_pairs[P.p2].ra * g_ff_pairs[P.p2].rb;
            int rc = ff_upload_table(n1, &P.Wn1); if (rc) return rc;
            rc = ff_upload_table(n2, &P.Wn2); if (rc) return rc;
            rc = ff_upload_table(N, &P.WN); if (rc) return rc;
        }
        it = g_ff_plans.emplace(N, P).first;
    }
    *plan = &it->second;
    return 0;
}

// na transforms of length Nz down the columns of row-major [Nz][Nr] arrays.  inverse: 0 forward, 1 inverse scaled
// by 1/Nz, 2 inverse unscaled.  rc 1: no plan for this length (the caller uses cuFFT).
int b2_fft_own(b2_ctx *ctx, int na, const void *const *in, void *const *out, int Nz, int Nr, int inverse,
               cudaStream_t s) {
    static const bool off = []() { const char *e = getenv("B2_FFT_IMPL"); return e && !strcmp(e, "cufft"); }();
    if (off || Nz < 16) return 1;
    std::lock_guard<std::mutex> lk(g_ff_mutex);
    const FfPlan *P;
    int rc = ff_get_plan(Nz, &P);
    if (rc) return rc;
    if (P->p1 < 0) return 1;
    const FfPair &k1 = g_ff_pairs[P->p1], &k2 = g_ff_pairs[P->p2];
    const int n1 = k1.ra * k1.rb, n2 = k2.ra * k2.rb;
    static const int group_env = []() { const char *e = getenv("B2_FFT_GROUP"); int g = e ? atoi(e) : 6;
                                        return g < 1 ? 1 : (g > FF_GROUP ? FF_GROUP : g); }();
    const int group = group_env;
    void *scratch;
    // (a transform issued on another stream than the context's runs beside the context stream's own transforms:
    //  private scratch)
    rc = b2_scratch(ctx, s == ctx->stream ? 2 : 3, sizeof(double2) * (size_t)group * Nz * Nr, &scratch);
    if (rc) return rc;
    const unsigned ncol = (unsigned)((Nr + FF_C - 1) / FF_C);
    for (int g0 = 0; g0 < na; g0 += group) {
        const int ng = (na - g0 < group) ? na - g0 : group;
        FfArgs A1, A2;
        memset(&A1, 0, sizeof(A1)); memset(&A2, 0, sizeof(A2));
        for (int k = 0; k < ng; ++k) {
            double2 *T = (double2 *)scratch + (size_t)k * Nz * Nr;
            A1.in[k] = (const double2 *)in[g0 + k]; A1.out[k] = T;
            A2.in[k] = T; A2.out[k] = (double2 *)out[g0 + k];
        }
        // pass 1: line j2, elements j1 (rows j1*n2 + j2) -> rows k1*n2 + j2, twiddle w_N^(j2*k1)
        A1.Wn = P->Wn1; A1.WN = P->WN; A1.line_stride_in = 1; A1.elem_stride_in = n2;
        A1.line_stride_out = 1; A1.elem_stride_out = n2; A1.Nr = Nr; A1.inverse = inverse ? 1 : 0; A1.scale = 1.;
        // pass 2: line k1, elements j2 (rows k1*n2 + j2) -> rows k1 + n1*k2
        A2.Wn = P->Wn2; A2.WN = nullptr; A2.line_stride_in = n2; A2.elem_stride_in = 1;
        A2.line_stride_out = 1; A2.elem_stride_out = n1; A2.Nr = Nr; A2.inverse = inverse ? 1 : 0;
        A2.scale = (inverse == 1) ? 1. / Nz : 1.;
        auto nthreads = [](const FfPair &k) { int m = k.ra > k.rb ? k.ra : k.rb; if (k.ra * k.p2 > m) m = k.ra * k.p2; return m * FF_C; };
        const int t1 = nthreads(k1), t2 = nthreads(k2);
        k1.fn<<<dim3(ncol, (unsigned)n2, (unsigned)ng), t1, 0, s>>>(A1);
        B2_LAUNCHED();
        k2.fn<<<dim3(ncol, (unsigned)n1, (unsigned)ng), t2, 0, s>>>(A2);
        B2_LAUNCHED();
    }
    return 0;
}

// 1 if the two-pass transform has a plan for this length whose radices are all <= max_radix (grid planners pick a
// guard width that gives such a local length), 0 otherwise (cuFFT would take it)
extern "C" int b2_fft_has_plan(int Nz, int max_radix) {
    std::lock_guard<std::mutex> lk(g_ff_mutex);
    for (int a = 0; a < g_ff_npairs; ++a)
        for (int b = 0; b < g_ff_npairs; ++b) {
            const FfPair &p = g_ff_pairs[a], &q = g_ff_pairs[b];
            if ((long long)p.ra * p.rb * q.ra * q.rb != Nz) continue;
            if (p.ra <= max_radix && p.rb <= max_radix && q.ra <= max_radix && q.rb <= max_radix) return 1;
        }
    return 0;
}

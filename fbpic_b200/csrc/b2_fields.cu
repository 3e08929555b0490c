// b2_fields.cu -- element-wise grid kernels (interpolation + spectral grids), z-FFT (cuFFT) and the
// z-boundary helpers.  All kernels are HBM-bound streaming passes over complex128 [Nz][Nr] arrays:
// 2-D grids of (iz, ir) with ir fastest, one double2 (16 B) per thread access, fully coalesced.
#include "b2_common.cuh"

static __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
static __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
static __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
static __device__ __forceinline__ double2 rmul(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
static __device__ __forceinline__ double2 imul(double2 a) { return make_double2(-a.y, a.x); }      // i*a
static __device__ __forceinline__ double2 nimul(double2 a) { return make_double2(a.y, -a.x); }     // -i*a

#define B2_2D_INDEX                                                   \
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;             \
    const int iz = blockIdx.y * blockDim.y + threadIdx.y;             \
    if (ir >= Nr || iz >= Nz) return;                                 \
    const size_t o = (size_t)iz * Nr + ir;

static inline dim3 grid2d(int Nz, int Nr, dim3 b) { return dim3((Nr + b.x - 1) / b.x, (Nz + b.y - 1) / b.y); }
static const dim3 BLK(64, 4);

// ---- F[iz,ir] *= v[ir]  (cuda_divide_*_by_volume, fields/cuda_methods.py:68-117) ----
__global__ void k_scale_r(B2Ptrs A, int na, const double *__restrict__ v, int Nz, int Nr) {
    B2_2D_INDEX
    const double s = v[ir];
    for (int k = 0; k < na; ++k) {
        double2 *a = (double2 *)A.p[k];
        double2 f = a[o];
        a[o] = make_double2(f.x * s, f.y * s);
    }
}

// ---- F[iz,ir] *= fz[iz]*fr[ir]  (cuda_filter_*, fields/cuda_methods.py:467-517) ----
__global__ void k_filter(B2Ptrs A, int na, const double *__restrict__ fz, const double *__restrict__ fr,
                         int Nz, int Nr) {
    B2_2D_INDEX
    const double s = fz[iz] * fr[ir];
    for (int k = 0; k < na; ++k) {
        double2 *a = (double2 *)A.p[k];
        double2 f = a[o];
        a[o] = make_double2(s * f.x, s * f.y);
    }
}

// ---- r,t <-> p,m (spectral_transform/cuda_methods.py:120-158) ----
__global__ void k_rt_to_pm(double2 *__restrict__ r, double2 *__restrict__ t, int Nz, int Nr) {
    B2_2D_INDEX
    const double2 vr = r[o], vt = t[o];
    r[o] = rmul(0.5, csub(vr, imul(vt)));     // p = (r - i t)/2
    t[o] = rmul(0.5, cadd(vr, imul(vt)));     // m = (r + i t)/2
}
__global__ void k_pm_to_rt(double2 *__restrict__ p, double2 *__restrict__ m, int Nz, int Nr) {
    B2_2D_INDEX
    const double2 vp = p[o], vm = m[o];
    p[o] = cadd(vp, vm);                      // r = p + m
    m[o] = imul(csub(vp, vm));                // t = i (p - m)
}

// ---- scale (inverse FFT normalisation, fourier.py:157) ----
__global__ void k_scale(double2 *__restrict__ a, double s, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = make_double2(a[i].x * s, a[i].y * s);
}

// ---- spectral update ------------------------------------------------------------------------
// Coefficients are real tables for the standard PSATD and complex ones for the comoving /
// Galilean scheme (psatd_coefs.py:66-163).
template <bool COMOVING>
struct Coef {
    typedef double T;
};
template <>
struct Coef<true> {
    typedef double2 T;
};
static __device__ __forceinline__ double2 kmul(double c, double2 a) { return rmul(c, a); }
static __device__ __forceinline__ double2 kmul(double2 c, double2 a) { return cmul(c, a); }
static __device__ __forceinline__ double2 ksub(double a, double b) { return make_double2(a - b, 0.); }
static __device__ __forceinline__ double2 ksub(double2 a, double2 b) { return csub(a, b); }

// F = -inv_k2 ( [T_cc j_corr] (rho_next - [T_eb] rho_prev) + i kz Jz + kr (Jp - Jm) )
// (fields/numba_methods.py:64-86 standard, :217-241 comoving)
template <bool COMOVING>
static __device__ __forceinline__ void correct_J(double2 &Jp, double2 &Jm, double2 &Jz, double2 rho_prev,
                                                 double2 rho_next, double kz, double kr, double inv_k2,
                                                 double inv_dt, const b2_spectral_mode &M, size_t o) {
    double2 drho;
    if constexpr (!COMOVING) {
        drho = rmul(inv_dt, csub(rho_next, rho_prev));
    } else {
        const double2 Tcc = ((const double2 *)M.T_cc)[o], jc = ((const double2 *)M.j_corr_coef)[o];
        const double2 Teb = ((const double2 *)M.T_eb)[o];
        drho = cmul(cmul(Tcc, jc), csub(rho_next, cmul(rho_prev, Teb)));
    }
    const double2 div = cadd(cadd(drho, rmul(kz, imul(Jz))), rmul(kr, csub(Jp, Jm)));
    const double2 F = rmul(-inv_k2, div);
    Jp = cadd(Jp, rmul(0.5 * kr, F));
    Jm = cadd(Jm, rmul(-0.5 * kr, F));
    Jz = cadd(Jz, rmul(kz, nimul(F)));
}

// fields/numba_methods.py:119-186 (standard), :278-355 (comoving); then push_rho (:443 cuda)
template <bool COMOVING>
static __device__ __forceinline__ void push_fields(double2 &Ep, double2 &Em, double2 &Ez, double2 &Bp, double2 &Bm,
                                                   double2 &Bz, double2 Jp, double2 Jm, double2 Jz, double2 rho_prev,
                                                   double2 rho_next, double kz, double kr, double dt, double V,
                                                   bool use_true_rho, const b2_spectral_mode &M, size_t o) {
    typedef typename Coef<COMOVING>::T CT;
    const double c2 = B2_C_LIGHT * B2_C_LIGHT;
    const double mu0 = M.mu_0, eps0 = M.epsilon_0;
    const double C = M.C[o], S_w = M.S_w[o];
    const CT j_coef = ((const CT *)M.j_coef)[o];
    const CT rpc = ((const CT *)M.rho_prev_coef)[o], rnc = ((const CT *)M.rho_next_coef)[o];
    double2 Teb = make_double2(1., 0.), Tcc = make_double2(1., 0.);
    if constexpr (COMOVING) { Teb = ((const double2 *)M.T_eb)[o]; Tcc = ((const double2 *)M.T_cc)[o]; }
    const double2 Ep_old = Ep, Em_old = Em, Ez_old = Ez;
    double2 rho_diff;
    if (use_true_rho) {
        rho_diff = csub(kmul(rnc, rho_next), kmul(rpc, rho_prev));
    } else {
        const double2 divE = cadd(rmul(kr, csub(Ep, Em)), rmul(kz, imul(Ez)));
        const double2 divJ = cadd(rmul(kr, csub(Jp, Jm)), rmul(kz, imul(Jz)));
        if constexpr (!COMOVING) {
            rho_diff = csub(kmul(ksub(rnc, rpc), rmul(eps0, divE)), kmul(rnc, rmul(dt, divJ)));
        } else {
            const double2 Trho = ((const double2 *)M.T_rho)[o];
            const double2 a = csub(cmul(Teb, rnc), rpc);
            rho_diff = cadd(cmul(a, rmul(eps0, divE)), cmul(cmul(Trho, rnc), divJ));
        }
    }
    // E push
    const double2 TC = COMOVING ? rmul(C, Teb) : make_double2(C, 0.);              // T_eb*C
    const double2 TS = COMOVING ? rmul(c2 * S_w, Teb) : make_double2(c2 * S_w, 0.); // c2*T_eb*S_w
    const double2 mJp = COMOVING ? rmul(mu0, cmul(Tcc, Jp)) : rmul(mu0, Jp);
    const double2 mJm = COMOVING ? rmul(mu0, cmul(Tcc, Jm)) : rmul(mu0, Jm);
    const double2 mJz = COMOVING ? rmul(mu0, cmul(Tcc, Jz)) : rmul(mu0, Jz);
    const double2 hBz = rmul(0.5 * kr, nimul(Bz));                                  // -i 0.5 kr Bz
    double2 nEp = cadd(cadd(cmul(TC, Ep), rmul(0.5 * kr, rho_diff)), cmul(TS, csub(cadd(hBz, rmul(kz, Bp)), mJp)));
    double2 nEm = cadd(csub(cmul(TC, Em), rmul(0.5 * kr, rho_diff)), cmul(TS, csub(csub(hBz, rmul(kz, Bm)), mJm)));
    double2 nEz = cadd(cadd(cmul(TC, Ez), rmul(kz, nimul(rho_diff))),
                       cmul(TS, csub(rmul(kr, imul(cadd(Bp, Bm))), mJz)));
    if constexpr (COMOVING) {
        // + j_coef * i kz V * J
        const double2 g = cmul(j_coef, make_double2(0., kz * V));
        nEp = cadd(nEp, cmul(g, Jp));
        nEm = cadd(nEm, cmul(g, Jm));
        nEz = cadd(nEz, cmul(g, Jz));
    }
    // B push
    const double2 TSb = COMOVING ? rmul(S_w, Teb) : make_double2(S_w, 0.);
    const double2 hEz = rmul(0.5 * kr, nimul(Ez_old));
    const double2 hJz = rmul(0.5 * kr, nimul(Jz));
    const double2 nBp = cadd(csub(cmul(TC, Bp), cmul(TSb, cadd(hEz, rmul(kz, Ep_old)))),
                             kmul(j_coef, cadd(hJz, rmul(kz, Jp))));
    const double2 nBm = cadd(csub(cmul(TC, Bm), cmul(TSb, csub(hEz, rmul(kz, Em_old)))),
                             kmul(j_coef, csub(hJz, rmul(kz, Jm))));
    const double2 nBz = cadd(csub(cmul(TC, Bz), cmul(TSb, rmul(kr, imul(cadd(Ep_old, Em_old))))),
                             kmul(j_coef, rmul(kr, imul(cadd(Jp, Jm)))));
    Ep = nEp; Em = nEm; Ez = nEz; Bp = nBp; Bm = nBm; Bz = nBz;
}

template <bool COMOVING, bool DO_CORRECT, bool DO_PUSH>
__global__ void __launch_bounds__(256)
k_spectral(b2_spectral_mode M, double dt, double V, int use_true_rho, int Nz, int Nr) {
    B2_2D_INDEX
    const double kz = M.kz[iz], kr = M.kr[ir];
    double2 Jp = ((double2 *)M.Jp)[o], Jm = ((double2 *)M.Jm)[o], Jz = ((double2 *)M.Jz)[o];
    const double2 rho_prev = ((double2 *)M.rho_prev)[o], rho_next = ((double2 *)M.rho_next)[o];
    if (DO_CORRECT) {
        correct_J<COMOVING>(Jp, Jm, Jz, rho_prev, rho_next, kz, kr, M.inv_k2[o], 1. / dt, M, o);
        ((double2 *)M.Jp)[o] = Jp; ((double2 *)M.Jm)[o] = Jm; ((double2 *)M.Jz)[o] = Jz;
    }
    if (DO_PUSH) {
        double2 Ep = ((double2 *)M.Ep)[o], Em = ((double2 *)M.Em)[o], Ez = ((double2 *)M.Ez)[o];
        double2 Bp = ((double2 *)M.Bp)[o], Bm = ((double2 *)M.Bm)[o], Bz = ((double2 *)M.Bz)[o];
        push_fields<COMOVING>(Ep, Em, Ez, Bp, Bm, Bz, Jp, Jm, Jz, rho_prev, rho_next, kz, kr, dt, V,
                              use_true_rho != 0, M, o);
        ((double2 *)M.Ep)[o] = Ep; ((double2 *)M.Em)[o] = Em; ((double2 *)M.Ez)[o] = Ez;
        ((double2 *)M.Bp)[o] = Bp; ((double2 *)M.Bm)[o] = Bm; ((double2 *)M.Bz)[o] = Bz;
        ((double2 *)M.rho_prev)[o] = rho_next;                // push_rho
        ((double2 *)M.rho_next)[o] = make_double2(0., 0.);
    }
}

// ---- z boundaries ----------------------------------------------------------------------------
// damp[iz] multiplies rows [0, nd) (left) and, mirrored, rows [Nz-nd, Nz) (right)
// (boundaries/cuda_methods.py:486-638, boundary_communicator.py:828-907)
__global__ void k_damp_z(B2Ptrs A, int na, const double *__restrict__ damp, int nd, int left, int right,
                         int Nz, int Nr) {
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;
    const int id = blockIdx.y * blockDim.y + threadIdx.y;
    if (ir >= Nr || id >= nd) return;
    const double d = damp[id];
    for (int k = 0; k < na; ++k) {
        double2 *a = (double2 *)A.p[k];
        if (left) { const size_t o = (size_t)id * Nr + ir; a[o] = rmul(d, a[o]); }
        if (right) { const size_t o = (size_t)(Nz - 1 - id) * Nr + ir; a[o] = rmul(d, a[o]); }
    }
}
// moving window: F[iz,:] *= shift[iz]^n_move (shift = exp(i kz dz)), power by repeated product as
// shift_spect_array_gpu does (fbpic/boundaries/moving_window.py:205-278)
__global__ void k_shift_spect(B2Ptrs A, int na, const double2 *__restrict__ shift, int n_move, int Nz, int Nr) {
    B2_2D_INDEX
    const double2 sft = shift[iz];
    double2 pw = make_double2(1., 0.);
    const int n = n_move < 0 ? -n_move : n_move;
    for (int i = 0; i < n; ++i) pw = cmul(pw, sft);
    if (n_move < 0) pw.y = -pw.y;
    for (int k = 0; k < na; ++k) {
        double2 *a = (double2 *)A.p[k];
        a[o] = cmul(a[o], pw);
    }
}
__global__ void k_add(double2 *__restrict__ dst, const double2 *__restrict__ src, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = cadd(dst[i], src[i]);
}

// Guard-cell staging: slab `a` of packed = rows [row0, row0+nrow) of array a (contiguous, n = nrow*Nr
// elements).  MODE 0: pack (array -> packed), 1: unpack (packed -> array), 2: unpack-add.
template <int MODE>
__global__ void k_halo(B2Ptrs A, double2 *__restrict__ packed, size_t off, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 *arr = (double2 *)A.p[blockIdx.y] + off;
    double2 *pk = packed + blockIdx.y * n;
    if (MODE == 0) pk[i] = arr[i];
    else if (MODE == 1) arr[i] = pk[i];
    else arr[i] = cadd(arr[i], pk[i]);
}

// ==============================================================================================
static int get_plan(b2_ctx *ctx, int Nz, int Nr, int lane, cufftHandle *plan) {
    const uint64_t key = ((uint64_t)lane << 56) | ((uint64_t)Nz << 28) | (uint32_t)Nr;
    auto it = ctx->fft_plans.find(key);
    if (it == ctx->fft_plans.end()) {
        cufftHandle h;
        int n[1] = {Nz};
        int embed[1] = {Nz};
        // Nr independent transforms of length Nz down the columns of the row-major [Nz][Nr] array:
        // element stride Nr, batch distance 1 -- no transposes (cf. fourier.py:118-126).
        cufftResult r = cufftPlanMany(&h, 1, n, embed, Nr, 1, embed, Nr, 1, CUFFT_Z2Z, Nr);
        if (r != CUFFT_SUCCESS) return b2_fail((int)r, "cufftPlanMany failed", __FILE__, __LINE__);
        ctx->fft_plans[key] = h;
        *plan = h;
    } else {
        *plan = it->second;
    }
    return 0;
}

// lane 0 is the caller's stream; lanes 1.. are context-owned non-blocking streams
static int fft_lanes_init(b2_ctx *ctx) {
    if (ctx->fft_lanes) return 0;
    int lanes = 4;
    if (const char *e = getenv("B2_FFT_LANES")) lanes = atoi(e);
    if (lanes < 1) lanes = 1;
    if (lanes > B2_FFT_MAX_LANES) lanes = B2_FFT_MAX_LANES;
    B2_CUDA(cudaEventCreateWithFlags(&ctx->fft_fork, cudaEventDisableTiming));
    for (int k = 1; k < lanes; ++k) {
        B2_CUDA(cudaStreamCreateWithFlags(&ctx->fft_lane[k], cudaStreamNonBlocking));
        B2_CUDA(cudaEventCreateWithFlags(&ctx->fft_join[k], cudaEventDisableTiming));
    }
    ctx->fft_lanes = lanes;
    return 0;
}

static int fft_exec(b2_ctx *ctx, int lane, cudaStream_t s, const void *in, void *out, int Nz, int Nr, int inverse) {
    cufftHandle plan;
    int rc = get_plan(ctx, Nz, Nr, lane, &plan);
    if (rc) return rc;
    cufftResult r = cufftSetStream(plan, s);
    if (r != CUFFT_SUCCESS) return b2_fail((int)r, "cufftSetStream failed", __FILE__, __LINE__);
    r = cufftExecZ2Z(plan, (cufftDoubleComplex *)in, (cufftDoubleComplex *)out, inverse ? CUFFT_INVERSE : CUFFT_FORWARD);
    if (r != CUFFT_SUCCESS) return b2_fail((int)r, "cufftExecZ2Z failed", __FILE__, __LINE__);
    g_b2_launches.fetch_add(1);
    if (inverse == 1) {      // inverse == 2: raw inverse (the 1/Nz factor is folded into the Hankel matrices)
        const size_t n = (size_t)Nz * Nr;
        k_scale<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((double2 *)out, 1. / Nz, n);
        B2_LAUNCHED();
    }
    return 0;
}

extern "C" {

int b2_scale_rows_by_r(b2_ctx *ctx, int na, void *const *arrays, const double *v, int Nz, int Nr, void *stream) {
    if (na > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    B2Ptrs A;
    for (int k = 0; k < na; ++k) A.p[k] = arrays[k];
    B2Prof prof_(B2P_ELEMENTWISE, b2_stream_of(ctx, stream));
    k_scale_r<<<grid2d(Nz, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>(A, na, v, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_filter(b2_ctx *ctx, int na, void *const *arrays, const double *fz, const double *fr, int Nz, int Nr,
              void *stream) {
    if (na > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    B2Ptrs A;
    for (int k = 0; k < na; ++k) A.p[k] = arrays[k];
    B2Prof prof_(B2P_ELEMENTWISE, b2_stream_of(ctx, stream));
    k_filter<<<grid2d(Nz, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>(A, na, fz, fr, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_rt_to_pm(b2_ctx *ctx, void *r, void *t, int Nz, int Nr, void *stream) {
    k_rt_to_pm<<<grid2d(Nz, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>((double2 *)r, (double2 *)t, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}
int b2_pm_to_rt(b2_ctx *ctx, void *p, void *m, int Nz, int Nr, void *stream) {
    k_pm_to_rt<<<grid2d(Nz, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>((double2 *)p, (double2 *)m, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_fft_z(b2_ctx *ctx, const void *in, void *out, int Nz, int Nr, int inverse, void *stream) {
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_FFT, s);
    const void *ins[1] = {in};
    void *outs[1] = {out};
    int rc = b2_fft_own(ctx, 1, ins, outs, Nz, Nr, inverse, s);      // b2_fft.cu; 1: no plan for this length
    if (rc != 1) return rc;
    return fft_exec(ctx, 0, s, in, out, Nz, Nr, inverse);
}

// The transforms of one call are independent and each of them is too small to fill the GPU
// (one 4096 x 256 Z2Z = 128 CTAs): they are spread over `B2_FFT_LANES` (default 4) streams forked from
// and joined back into the caller's stream with events, so the caller still sees stream order.
int b2_fft_z_multi(b2_ctx *ctx, int na, const void *const *in, void *const *out, int Nz, int Nr, int inverse,
                   void *stream) {
    if (na <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_FFT, s);
    int rc = b2_fft_own(ctx, na, in, out, Nz, Nr, inverse, s);       // two-pass transform of b2_fft.cu
    if (rc != 1) return rc;
    // no (n1, n2) plan for this length: cuFFT, the independent transforms spread over concurrent lanes
    rc = fft_lanes_init(ctx);
    if (rc) return rc;
    const int lanes = ctx->fft_lanes < na ? ctx->fft_lanes : na;
    if (lanes > 1) {
        B2_CUDA(cudaEventRecord(ctx->fft_fork, s));
        for (int l = 1; l < lanes; ++l) B2_CUDA(cudaStreamWaitEvent(ctx->fft_lane[l], ctx->fft_fork, 0));
    }
    for (int k = 0; k < na; ++k) {
        const int l = k % lanes;
        rc = fft_exec(ctx, l, l ? ctx->fft_lane[l] : s, in[k], out[k], Nz, Nr, inverse);
        if (rc) return rc;
    }
    for (int l = 1; l < lanes; ++l) {
        B2_CUDA(cudaEventRecord(ctx->fft_join[l], ctx->fft_lane[l]));
        B2_CUDA(cudaStreamWaitEvent(s, ctx->fft_join[l], 0));
    }
    return 0;
}

static int spectral_launch(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, bool corr, bool push, double dt,
                           double V, int use_true_rho, int Nz, int Nr, void *stream) {
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_SPECTRAL, s);
    dim3 g = grid2d(Nz, Nr, BLK);
#define SPEC(CM, A, B) k_spectral<CM, A, B><<<g, BLK, 0, s>>>(*mode, dt, V, use_true_rho, Nz, Nr)
    if (comoving) {
        if (corr && push) SPEC(true, true, true); else if (corr) SPEC(true, true, false); else SPEC(true, false, true);
    } else {
        if (corr && push) SPEC(false, true, true); else if (corr) SPEC(false, true, false); else SPEC(false, false, true);
    }
#undef SPEC
    B2_LAUNCHED();
    return 0;
}

int b2_correct_currents(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double inv_dt, int Nz, int Nr,
                        void *stream) {
    return spectral_launch(ctx, mode, comoving, true, false, 1. / inv_dt, 0., 0, Nz, Nr, stream);
}
int b2_push_eb(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double dt, double V, int use_true_rho,
               int Nz, int Nr, void *stream) {
    return spectral_launch(ctx, mode, comoving, false, true, dt, V, use_true_rho, Nz, Nr, stream);
}
int b2_correct_push(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double dt, double V, int use_true_rho,
                    int Nz, int Nr, void *stream) {
    return spectral_launch(ctx, mode, comoving, true, true, dt, V, use_true_rho, Nz, Nr, stream);
}

int b2_damp_z(b2_ctx *ctx, int na, void *const *arrays, const double *damp, int nd, int left, int right, int Nz,
              int Nr, void *stream) {
    if (na > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    if (nd <= 0 || (!left && !right)) return 0;
    B2Ptrs A;
    for (int k = 0; k < na; ++k) A.p[k] = arrays[k];
    k_damp_z<<<grid2d(nd, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>(A, na, damp, nd, left, right, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_shift_spect(b2_ctx *ctx, int na, void *const *arrays, const void *shift, int n_move, int Nz, int Nr,
                   void *stream) {
    if (na > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    if (n_move == 0) return 0;
    B2Ptrs A;
    for (int k = 0; k < na; ++k) A.p[k] = arrays[k];
    B2Prof prof_(B2P_ELEMENTWISE, b2_stream_of(ctx, stream));
    k_shift_spect<<<grid2d(Nz, Nr, BLK), BLK, 0, b2_stream_of(ctx, stream)>>>(A, na, (const double2 *)shift, n_move, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_halo_stage(b2_ctx *ctx, int mode, int na, void *const *arrays, int row0, int nrow, int Nr, void *packed,
                  void *stream) {
    if (na <= 0 || nrow <= 0) return 0;
    if (na > B2_MAX_ARRAYS) return b2_fail(-3, "too many arrays", __FILE__, __LINE__);
    if (mode < 0 || mode > 2) return b2_fail(-3, "b2_halo_stage: mode must be 0 (pack), 1 (unpack), 2 (unpack-add)", __FILE__, __LINE__);
    B2Ptrs A;
    for (int k = 0; k < na; ++k) A.p[k] = arrays[k];
    const size_t n = (size_t)nrow * Nr, off = (size_t)row0 * Nr;
    dim3 g((unsigned)((n + 255) / 256), na);
    cudaStream_t s = b2_stream_of(ctx, stream);
    if (mode == 0) k_halo<0><<<g, 256, 0, s>>>(A, (double2 *)packed, off, n);
    else if (mode == 1) k_halo<1><<<g, 256, 0, s>>>(A, (double2 *)packed, off, n);
    else k_halo<2><<<g, 256, 0, s>>>(A, (double2 *)packed, off, n);
    B2_LAUNCHED();
    return 0;
}

int b2_add_rows(b2_ctx *ctx, void *dst, const void *src, int nrows, int Nr, void *stream) {
    const size_t n = (size_t)nrows * Nr;
    if (!n) return 0;
    k_add<<<(unsigned)((n + 255) / 256), 256, 0, b2_stream_of(ctx, stream)>>>((double2 *)dst, (const double2 *)src, n);
    B2_LAUNCHED();
    return 0;
}

}  // extern "C"

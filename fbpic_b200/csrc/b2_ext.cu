// b2_ext.cu -- launchers / C ABI of the solver variants around the hot loop:
// radial PML, cross-deposition current correction, laser-antenna virtual particles.
// The kernel bodies live in b2_ext_kernels.cuh (shared with the host emulation of tests/hostemu).
#include "b2_common.cuh"
#include "b2_ext_kernels.cuh"

static inline dim3 x_grid2d(int Nz, int Nr, dim3 b) { return dim3((Nr + b.x - 1) / b.x, (Nz + b.y - 1) / b.y); }
static const dim3 XBLK(64, 4);

extern "C" {

int b2_push_eb_pml(b2_ctx *ctx, void *Ep_pml, void *Em_pml, void *Bp_pml, void *Bm_pml, const void *Ez,
                   const void *Bz, const double *C, const double *S_w, const void *T_eb, const double *kr,
                   int Nz, int Nr, void *stream) {
    if (Nz <= 0 || Nr <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_SPECTRAL, s);
    const dim3 g = x_grid2d(Nz, Nr, XBLK);
    if (T_eb)
        b2ext::k_push_eb_pml<true><<<g, XBLK, 0, s>>>((double2 *)Ep_pml, (double2 *)Em_pml, (double2 *)Bp_pml,
                                                     (double2 *)Bm_pml, (const double2 *)Ez, (const double2 *)Bz, C,
                                                     S_w, (const double2 *)T_eb, kr, Nz, Nr);
    else
        b2ext::k_push_eb_pml<false><<<g, XBLK, 0, s>>>((double2 *)Ep_pml, (double2 *)Em_pml, (double2 *)Bp_pml,
                                                      (double2 *)Bm_pml, (const double2 *)Ez, (const double2 *)Bz, C,
                                                      S_w, nullptr, kr, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_damp_pml(b2_ctx *ctx, void *Et, void *Et_pml, void *Ez, void *Bt, void *Bt_pml, void *Bz,
                const double *damp, int n_pml, int Nz, int Nr, void *stream) {
    if (n_pml <= 0 || Nz <= 0) return 0;
    if (n_pml > Nr) return b2_fail(-3, "b2_damp_pml: n_pml exceeds Nr", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_ELEMENTWISE, s);
    const dim3 blk(32, 8);
    b2ext::k_damp_pml<<<x_grid2d(Nz, n_pml, blk), blk, 0, s>>>((double2 *)Et, (double2 *)Et_pml, (double2 *)Ez,
                                                              (double2 *)Bt, (double2 *)Bt_pml, (double2 *)Bz, damp,
                                                              n_pml, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_correct_currents_cross(b2_ctx *ctx, const b2_spectral_mode *M, const void *rho_next_z,
                              const void *rho_next_xy, int comoving, double inv_dt, int Nz, int Nr, void *stream) {
    if (Nz <= 0 || Nr <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_SPECTRAL, s);
    const dim3 g = x_grid2d(Nz, Nr, XBLK);
    if (comoving)
        b2ext::k_correct_cross<true><<<g, XBLK, 0, s>>>(
            (const double2 *)M->rho_prev, (const double2 *)M->rho_next, (const double2 *)rho_next_z,
            (const double2 *)rho_next_xy, (double2 *)M->Jp, (double2 *)M->Jm, (double2 *)M->Jz, M->kz, M->kr,
            (const double2 *)M->T_cc, (const double2 *)M->j_corr_coef, (const double2 *)M->T_eb, inv_dt, Nz, Nr);
    else
        b2ext::k_correct_cross<false><<<g, XBLK, 0, s>>>(
            (const double2 *)M->rho_prev, (const double2 *)M->rho_next, (const double2 *)rho_next_z,
            (const double2 *)rho_next_xy, (double2 *)M->Jp, (double2 *)M->Jm, (double2 *)M->Jz, M->kz, M->kr,
            nullptr, nullptr, nullptr, inv_dt, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_correct_divE(b2_ctx *ctx, const b2_spectral_mode *M, int Nz, int Nr, void *stream) {
    if (Nz <= 0 || Nr <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_SPECTRAL, s);
    b2ext::k_correct_divE<<<x_grid2d(Nz, Nr, XBLK), XBLK, 0, s>>>((double2 *)M->Ep, (double2 *)M->Em, (double2 *)M->Ez,
                                                                (const double2 *)M->rho_prev, M->kz, M->kr, M->inv_k2,
                                                                1. / M->epsilon_0, Nz, Nr);
    B2_LAUNCHED();
    return 0;
}

int b2_push_p_after_plane(b2_ctx *ctx, int64_t n, const double *z, double z_plane, double *ux, double *uy, double *uz,
                          double *inv_gamma, const double *Ex, const double *Ey, const double *Ez, const double *Bx,
                          const double *By, const double *Bz, double q, double m, double dt, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_PUSH, s);
    const double econst = q * dt / (m * B2_C_LIGHT), bconst = 0.5 * q * dt / m;
    b2ext::k_push_p_after_plane<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((long long)n, z, z_plane, ux, uy, uz,
                                                                           inv_gamma, Ex, Ey, Ez, Bx, By, Bz, econst, bconst);
    B2_LAUNCHED();
    return 0;
}

int b2_antenna_particles(b2_ctx *ctx, int64_t n, const double *bx, const double *by, const double *ex,
                         const double *ey, const double *vx, const double *vy, const double *vz, double sign,
                         double *x, double *y, double *ux, double *uy, double *uz, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    b2ext::k_antenna_particles<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((long long)n, bx, by, ex, ey, vx, vy, vz,
                                                                          sign, x, y, ux, uy, uz);
    B2_LAUNCHED();
    return 0;
}

int b2_axpy(b2_ctx *ctx, int64_t n, double a, const double *x, double *y, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    b2ext::k_axpy<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((long long)n, a, x, y);
    B2_LAUNCHED();
    return 0;
}

int b2_push_p_ioniz(b2_ctx *ctx, int64_t n, const uint64_t *level, double *ux, double *uy, double *uz,
                    double *inv_gamma, const double *Ex, const double *Ey, const double *Ez, const double *Bx,
                    const double *By, const double *Bz, double m, double dt, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_PUSH, s);
    const double e = 1.602176634e-19;
    b2ext::k_push_p_ioniz<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        (long long)n, (const unsigned long long *)level, ux, uy, uz, inv_gamma, Ex, Ey, Ez, Bx, By, Bz,
        e * dt / (m * B2_C_LIGHT), 0.5 * e * dt / m);
    B2_LAUNCHED();
    return 0;
}

int b2_w_times_level(b2_ctx *ctx, int64_t n, const double *w, const uint64_t *level, double *out, void *stream) {
    if (n <= 0) return 0;
    cudaStream_t s = b2_stream_of(ctx, stream);
    b2ext::k_w_times_level<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((long long)n, w,
                                                                      (const unsigned long long *)level, out);
    B2_LAUNCHED();
    return 0;
}

int b2_ionize(b2_ctx *ctx, int64_t n, uint64_t *level, int level_max, const double *adk_prefactor,
              const double *adk_power, const double *adk_exp_prefactor, const double *ux, const double *uy,
              const double *uz, const double *Ex, const double *Ey, const double *Ez, const double *Bx,
              const double *By, const double *Bz, const double *draws, uint64_t seed, int64_t cap, int64_t *d_events,
              int64_t *d_count, int64_t *h_count, void *stream) {
    if (!d_count || !h_count || (cap > 0 && !d_events))
        return b2_fail(-3, "b2_ionize: missing buffer", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), s));
    if (n > 0) {
        b2ext::k_ionize<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
            (long long)n, (unsigned long long *)level, level_max, adk_prefactor, adk_power, adk_exp_prefactor, ux, uy,
            uz, Ex, Ey, Ez, Bx, By, Bz, B2_C_LIGHT, draws, (unsigned long long)seed, (long long)cap,
            (long long *)d_events, (unsigned long long *)d_count);
        B2_LAUNCHED();
    }
    B2_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return 0;
}

static b2ext::ComptonParams compton_params(const double *p20, uint64_t seed) {
    b2ext::ComptonParams P;
    P.ct = p20[0]; P.photon_n_lab_peak = p20[1]; P.inv_laser_waist2 = p20[2]; P.inv_laser_ctau2 = p20[3];
    P.laser_initial_z0 = p20[4]; P.gamma_boost = p20[5]; P.beta_boost = p20[6];
    P.photon_p = p20[7]; P.photon_px = p20[8]; P.photon_py = p20[9]; P.photon_pz = p20[10];
    P.photon_beta_x = p20[11]; P.photon_beta_y = p20[12]; P.photon_beta_z = p20[13];
    P.dt = p20[14]; P.ratio_w_electron_photon = p20[15]; P.inv_ratio_w_elec_photon = p20[16];
    P.pi_re2 = p20[17]; P.inv_mc = p20[18]; P.c_light = p20[19];
    P.seed = seed;
    return P;
}

int b2_compton_count(b2_ctx *ctx, int64_t n, const double *x, const double *y, const double *z, const double *ux,
                     const double *uy, const double *uz, const double *inv_gamma, const double *params20,
                     uint64_t seed, int32_t *d_nscatter, int64_t *d_total, int64_t *h_total, void *stream) {
    if (!params20 || !d_total || !h_total || (n > 0 && !d_nscatter))
        return b2_fail(-3, "b2_compton_count: missing buffer", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int64_t), s));
    if (n > 0) {
        b2ext::k_compton_count<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
            (long long)n, x, y, z, ux, uy, uz, inv_gamma, compton_params(params20, seed), d_nscatter,
            (unsigned long long *)d_total);
        B2_LAUNCHED();
    }
    B2_CUDA(cudaMemcpyAsync(h_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int b2_compton_scatter(b2_ctx *ctx, int64_t n, const int32_t *d_nscatter, const double *x, const double *y,
                       const double *z, double *ux, double *uy, double *uz, const double *inv_gamma, const double *w,
                       const double *params20, uint64_t seed, double *const *photon8, int64_t *d_cursor,
                       void *stream) {
    if (n <= 0) return 0;
    if (!params20 || !photon8 || !d_cursor) return b2_fail(-3, "b2_compton_scatter: missing buffer", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2_CUDA(cudaMemsetAsync(d_cursor, 0, sizeof(int64_t), s));
    b2ext::k_compton_scatter<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        (long long)n, d_nscatter, x, y, z, ux, uy, uz, inv_gamma, w, compton_params(params20, seed), photon8[0],
        photon8[1], photon8[2], photon8[3], photon8[4], photon8[5], photon8[6], photon8[7],
        (unsigned long long *)d_cursor);
    B2_LAUNCHED();
    return 0;
}

int b2_extract_slice(b2_ctx *ctx, const void *const *fields10, int m, int Nm, int Nz, int Nr, int Nr_out, int iz,
                     double Sz, double *slice, void *stream) {
    if (m < 0 || m >= Nm || Nr_out <= 0 || Nr_out > Nr)
        return b2_fail(-3, "b2_extract_slice: bad mode or radial size", __FILE__, __LINE__);
    if (iz < 0 || iz + 1 >= Nz) return b2_fail(-3, "b2_extract_slice: slice outside of the grid", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    b2ext::SliceFields F;
    for (int k = 0; k < 10; ++k) F.f[k] = (const double2 *)fields10[k];
    const dim3 grid((unsigned)((Nr_out + 127) / 128), 10);
    b2ext::k_extract_slice<<<grid, 128, 0, s>>>(F, m, 2 * Nm - 1, Nr, Nr_out, iz, Sz, slice);
    B2_LAUNCHED();
    return 0;
}

int b2_select_crossing(b2_ctx *ctx, int64_t n, const double *z, const double *uz, const double *inv_gamma,
                       double c_light, double dt, double z_curr, double z_prev, int64_t cap, int64_t *d_idx,
                       int64_t *d_count, int64_t *h_count, void *stream) {
    if (!d_count || !h_count || (cap > 0 && !d_idx))
        return b2_fail(-3, "b2_select_crossing: missing buffer", __FILE__, __LINE__);
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), s));
    if (n > 0) {
        b2ext::k_select_crossing<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
            (long long)n, z, uz, inv_gamma, c_light, dt, z_curr, z_prev, (long long)cap, (long long *)d_idx,
            (unsigned long long *)d_count);
        B2_LAUNCHED();
    }
    B2_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"

// =================================================================================================
// External fields: user-supplied element-wise expression F' = f(F, x, y, z, t, amplitude, length_scale)
// applied to a gathered field of the particles after the gather (ExternalField.apply_expression,
// fbpic/lpa_utils/external_fields.py:183-215).  The reference JIT-compiles the user's Python function
// with Numba (`cuda.jit(func, device=True)` + `compile_cupy`, :134-147); here the expression arrives as
// CUDA C (translated from the Python function by the host layer), is compiled once by NVRTC to an
// sm_100a cubin and loaded through the runtime's library API.  For boosted-frame runs the kernel
// evaluates the lab-frame expression at zlab = g (z + b c t), tlab = g (t + b z / c) (:118-126).
// =================================================================================================
#include <nvrtc.h>
#include <dlfcn.h>
#include <string>
#include <vector>

// NVRTC is bound at first use (dlopen), not at link time: the hot path of the library must load on a
// machine without the NVRTC runtime.
struct B2Nvrtc {
    decltype(&nvrtcCreateProgram) create;
    decltype(&nvrtcCompileProgram) compile;
    decltype(&nvrtcGetProgramLogSize) log_size;
    decltype(&nvrtcGetProgramLog) log;
    decltype(&nvrtcGetCUBINSize) cubin_size;
    decltype(&nvrtcGetCUBIN) cubin;
    decltype(&nvrtcDestroyProgram) destroy;
    decltype(&nvrtcGetErrorString) error_string;
    bool ok;
};
static B2Nvrtc *b2_nvrtc() {
    static B2Nvrtc api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        api.ok = false;
        void *h = nullptr;
        const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (int k = 0; k < 3 && !h; ++k) h = dlopen(names[k], RTLD_NOW | RTLD_LOCAL);
        if (h) {
#define B2_SYM(field, name) api.field = (decltype(api.field))dlsym(h, #name)
            B2_SYM(create, nvrtcCreateProgram); B2_SYM(compile, nvrtcCompileProgram);
            B2_SYM(log_size, nvrtcGetProgramLogSize); B2_SYM(log, nvrtcGetProgramLog);
            B2_SYM(cubin_size, nvrtcGetCUBINSize); B2_SYM(cubin, nvrtcGetCUBIN);
            B2_SYM(destroy, nvrtcDestroyProgram); B2_SYM(error_string, nvrtcGetErrorString);
#undef B2_SYM
            api.ok = api.create && api.compile && api.log_size && api.log && api.cubin_size && api.cubin &&
                     api.destroy && api.error_string;
        }
    }
    return &api;
}

struct B2ExtField {
    std::vector<char> cubin;
    cudaLibrary_t lib;
    cudaKernel_t kernel;
    bool loaded;
};

static const char *kExtFieldPrologue =
    "extern \"C\" __global__ void b2_external_field(long long n_, double *F_, const double *x_, const double *y_,\n"
    "        const double *z_, double t_, double amplitude, double length_scale, double gamma_b_, double beta_b_) {\n"
    "    const long long i_ = blockIdx.x * (long long)blockDim.x + threadIdx.x;\n"
    "    if (i_ >= n_) return;\n"
    "    const double c_ = 299792458.0;\n"
    "    const double F = F_[i_], x = x_[i_], y = y_[i_];\n"
    "    const double z = gamma_b_ * (z_[i_] + beta_b_ * c_ * t_);\n"
    "    const double t = gamma_b_ * (t_ + beta_b_ * (1. / c_) * z_[i_]);\n"
    "    (void)F; (void)x; (void)y; (void)z; (void)t; (void)amplitude; (void)length_scale;\n";

extern "C" {

int b2_external_field_compile(const char *cuda_body, void **handle) {
    if (!cuda_body || !handle) return b2_fail(-3, "b2_external_field_compile: null argument", __FILE__, __LINE__);
    std::string src(kExtFieldPrologue);
    src += cuda_body;            // statements ending with `F_[i_] = <expression>;`
    src += "\n}\n";
    B2Nvrtc *rt = b2_nvrtc();
    if (!rt->ok) return b2_fail(-5, "external fields need the NVRTC runtime (libnvrtc.so.12), which could not be loaded", __FILE__, __LINE__);
    nvrtcProgram prog;
    nvrtcResult r = rt->create(&prog, src.c_str(), "b2_external_field.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return b2_fail((int)r, rt->error_string(r), __FILE__, __LINE__);
    // -fmad=false: products and sums are rounded separately, like the NumPy / Numba-CPU evaluation
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-fmad=false"};
    r = rt->compile(prog, 3, opts);
    if (r != NVRTC_SUCCESS) {
        size_t ls = 0;
        rt->log_size(prog, &ls);
        std::string log(ls ? ls : 1, '\0');
        if (ls) rt->log(prog, &log[0]);
        rt->destroy(&prog);
        snprintf(g_b2_err, sizeof(g_b2_err), "external field expression does not compile: %.400s", log.c_str());
        return (int)r;
    }
    size_t nb = 0;
    rt->cubin_size(prog, &nb);
    B2ExtField *h = new B2ExtField();
    h->cubin.resize(nb);
    rt->cubin(prog, h->cubin.data());
    rt->destroy(&prog);
    h->loaded = false;
    *handle = h;
    return 0;
}

int b2_external_field_cubin_size(void *handle, size_t *nbytes) {
    if (!handle || !nbytes) return b2_fail(-3, "b2_external_field_cubin_size: null argument", __FILE__, __LINE__);
    *nbytes = ((B2ExtField *)handle)->cubin.size();
    return 0;
}

int b2_external_field_apply(b2_ctx *ctx, void *handle, int64_t n, double *F, const double *x, const double *y,
                            const double *z, double t, double amplitude, double length_scale, double gamma_boost,
                            double beta_boost, void *stream) {
    if (!handle) return b2_fail(-3, "b2_external_field_apply: null handle", __FILE__, __LINE__);
    if (n <= 0) return 0;
    B2ExtField *h = (B2ExtField *)handle;
    if (!h->loaded) {
        B2_CUDA(cudaLibraryLoadData(&h->lib, h->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        B2_CUDA(cudaLibraryGetKernel(&h->kernel, h->lib, "b2_external_field"));
        h->loaded = true;
    }
    long long nn = (long long)n;
    void *args[] = {&nn, &F, &x, &y, &z, &t, &amplitude, &length_scale, &gamma_boost, &beta_boost};
    cudaStream_t s = b2_stream_of(ctx, stream);
    B2Prof prof_(B2P_PUSH, s);
    B2_CUDA(cudaLaunchKernel((const void *)h->kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), args, 0, s));
    B2_LAUNCHED();
    return 0;
}

int b2_external_field_free(void *handle) {
    if (!handle) return 0;
    B2ExtField *h = (B2ExtField *)handle;
    if (h->loaded) cudaLibraryUnload(h->lib);
    delete h;
    return 0;
}

}  // extern "C"

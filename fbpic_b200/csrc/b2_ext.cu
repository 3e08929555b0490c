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

}  // extern "C"

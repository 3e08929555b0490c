// tests/hostemu/emu_ext.cpp -- TEST INFRASTRUCTURE.  Runs the kernel bodies of
// fbpic_b200/csrc/b2_ext_kernels.cuh on the CPU with the launch geometry of b2_ext.cu, so that their
// index arithmetic and formulas can be checked against NumPy / reference fixtures in the GPU-less
// build container.  Never loaded by the product (fbpic_b200/_lib.py only knows libfbpic_b200.so).
#include "cuda_shim.h"
#include "../../fbpic_b200/csrc/b2_ext_kernels.cuh"

static emu_dim3 grid2d(int Nz, int Nr, emu_dim3 b) { return emu_dim3((Nr + b.x - 1) / b.x, (Nz + b.y - 1) / b.y); }

extern "C" {

int emu_push_eb_pml(void *Ep, void *Em, void *Bp, void *Bm, const void *Ez, const void *Bz, const double *C,
                    const double *S_w, const void *T_eb, const double *kr, int Nz, int Nr) {
    emu_dim3 blk(64, 4);
    if (T_eb)
        EMU_LAUNCH(grid2d(Nz, Nr, blk), blk, b2ext::k_push_eb_pml<true>, (double2 *)Ep, (double2 *)Em, (double2 *)Bp,
                   (double2 *)Bm, (const double2 *)Ez, (const double2 *)Bz, C, S_w, (const double2 *)T_eb, kr, Nz, Nr);
    else
        EMU_LAUNCH(grid2d(Nz, Nr, blk), blk, b2ext::k_push_eb_pml<false>, (double2 *)Ep, (double2 *)Em, (double2 *)Bp,
                   (double2 *)Bm, (const double2 *)Ez, (const double2 *)Bz, C, S_w, (const double2 *)nullptr, kr, Nz, Nr);
    return 0;
}

int emu_damp_pml(void *Et, void *Et_pml, void *Ez, void *Bt, void *Bt_pml, void *Bz, const double *damp, int n_pml,
                 int Nz, int Nr) {
    emu_dim3 blk(32, 8);
    EMU_LAUNCH(grid2d(Nz, n_pml, blk), blk, b2ext::k_damp_pml, (double2 *)Et, (double2 *)Et_pml, (double2 *)Ez,
               (double2 *)Bt, (double2 *)Bt_pml, (double2 *)Bz, damp, n_pml, Nz, Nr);
    return 0;
}

int emu_correct_currents_cross(const void *rho_prev, const void *rho_next, const void *rho_next_z,
                               const void *rho_next_xy, void *Jp, void *Jm, void *Jz, const double *kz,
                               const double *kr, const void *T_cc, const void *j_corr_coef, const void *T_eb,
                               int comoving, double inv_dt, int Nz, int Nr) {
    emu_dim3 blk(64, 4);
    if (comoving)
        EMU_LAUNCH(grid2d(Nz, Nr, blk), blk, b2ext::k_correct_cross<true>, (const double2 *)rho_prev,
                   (const double2 *)rho_next, (const double2 *)rho_next_z, (const double2 *)rho_next_xy, (double2 *)Jp,
                   (double2 *)Jm, (double2 *)Jz, kz, kr, (const double2 *)T_cc, (const double2 *)j_corr_coef,
                   (const double2 *)T_eb, inv_dt, Nz, Nr);
    else
        EMU_LAUNCH(grid2d(Nz, Nr, blk), blk, b2ext::k_correct_cross<false>, (const double2 *)rho_prev,
                   (const double2 *)rho_next, (const double2 *)rho_next_z, (const double2 *)rho_next_xy, (double2 *)Jp,
                   (double2 *)Jm, (double2 *)Jz, kz, kr, (const double2 *)nullptr, (const double2 *)nullptr,
                   (const double2 *)nullptr, inv_dt, Nz, Nr);
    return 0;
}

int emu_correct_divE(void *Ep, void *Em, void *Ez, const void *rho_prev, const double *kz, const double *kr,
                     const double *inv_k2, double inv_eps0, int Nz, int Nr) {
    emu_dim3 blk(64, 4);
    EMU_LAUNCH(grid2d(Nz, Nr, blk), blk, b2ext::k_correct_divE, (double2 *)Ep, (double2 *)Em, (double2 *)Ez,
               (const double2 *)rho_prev, kz, kr, inv_k2, inv_eps0, Nz, Nr);
    return 0;
}

int emu_push_p_after_plane(long long n, const double *z, double z_plane, double *ux, double *uy, double *uz,
                           double *ig, const double *Ex, const double *Ey, const double *Ez, const double *Bx,
                           const double *By, const double *Bz, double econst, double bconst) {
    EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_push_p_after_plane, n, z, z_plane, ux, uy,
               uz, ig, Ex, Ey, Ez, Bx, By, Bz, econst, bconst);
    return 0;
}

int emu_antenna_particles(long long n, const double *bx, const double *by, const double *ex, const double *ey,
                          const double *vx, const double *vy, const double *vz, double sign, double *x, double *y,
                          double *ux, double *uy, double *uz) {
    EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_antenna_particles, n, bx, by, ex, ey, vx,
               vy, vz, sign, x, y, ux, uy, uz);
    return 0;
}

int emu_axpy(long long n, double a, const double *x, double *y) {
    EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_axpy, n, a, x, y);
    return 0;
}

int emu_push_p_ioniz(long long n, const unsigned long long *level, double *ux, double *uy, double *uz, double *ig,
                     const double *Ex, const double *Ey, const double *Ez, const double *Bx, const double *By,
                     const double *Bz, double econst1, double bconst1) {
    EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_push_p_ioniz, n, level, ux, uy, uz, ig, Ex,
               Ey, Ez, Bx, By, Bz, econst1, bconst1);
    return 0;
}

int emu_w_times_level(long long n, const double *w, const unsigned long long *level, double *out) {
    EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_w_times_level, n, w, level, out);
    return 0;
}

int emu_ionize(long long n, unsigned long long *level, int level_max, const double *pre, const double *pw,
               const double *ex, const double *ux, const double *uy, const double *uz, const double *Ex,
               const double *Ey, const double *Ez, const double *Bx, const double *By, const double *Bz,
               double c_light, const double *draws, unsigned long long seed, long long cap, long long *events,
               unsigned long long *count) {
    *count = 0;
    if (n > 0)
        EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_ionize, n, level, level_max, pre, pw,
                   ex, ux, uy, uz, Ex, Ey, Ez, Bx, By, Bz, c_light, draws, seed, cap, events, count);
    return 0;
}

static b2ext::ComptonParams emu_compton_params(const double *p20, unsigned long long seed) {
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

int emu_compton_count(long long n, const double *x, const double *y, const double *z, const double *ux,
                      const double *uy, const double *uz, const double *ig, const double *p20,
                      unsigned long long seed, int *nscatter, unsigned long long *total) {
    *total = 0;
    if (n > 0)
        EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_compton_count, n, x, y, z, ux, uy, uz,
                   ig, emu_compton_params(p20, seed), nscatter, total);
    return 0;
}

int emu_compton_scatter(long long n, const int *nscatter, const double *x, const double *y, const double *z,
                        double *ux, double *uy, double *uz, const double *ig, const double *w, const double *p20,
                        unsigned long long seed, double *const *ph, unsigned long long *cursor) {
    *cursor = 0;
    if (n > 0)
        EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_compton_scatter, n, nscatter, x, y, z,
                   ux, uy, uz, ig, w, emu_compton_params(p20, seed), ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6],
                   ph[7], cursor);
    return 0;
}

int emu_select_crossing(long long n, const double *z, const double *uz, const double *inv_gamma, double c_light,
                        double dt, double z_curr, double z_prev, long long cap, long long *idx,
                        unsigned long long *count) {
    *count = 0;
    if (n > 0)
        EMU_LAUNCH(emu_dim3((unsigned)((n + 255) / 256)), emu_dim3(256), b2ext::k_select_crossing, n, z, uz, inv_gamma,
                   c_light, dt, z_curr, z_prev, cap, idx, count);
    return 0;
}

int emu_extract_slice(const void *const *fields10, int m, int Nm, int Nz, int Nr, int Nr_out, int iz, double Sz,
                      double *slice) {
    b2ext::SliceFields F;
    for (int k = 0; k < 10; ++k) F.f[k] = (const double2 *)fields10[k];
    EMU_LAUNCH(emu_dim3((unsigned)((Nr_out + 127) / 128), 10), emu_dim3(128), b2ext::k_extract_slice, F, m, 2 * Nm - 1,
               Nr, Nr_out, iz, Sz, slice);
    return 0;
}

}  // extern "C"

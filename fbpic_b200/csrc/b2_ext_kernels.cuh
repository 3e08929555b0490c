// b2_ext_kernels.cuh -- kernels of the solver variants around the hot loop: radial PML,
// cross-deposition current correction, laser-antenna virtual particles.  All of them are HBM-bound
// streaming passes: 2-D (iz, ir) grids with ir fastest (one 16-B double2 per access, coalesced) or
// 1-D particle loops.  This header holds ONLY the __global__ bodies (no launches, no runtime calls) so
// that tests/hostemu can compile the very same source for the CPU and check it without a GPU.
#pragma once

#ifndef B2_C_LIGHT
#define B2_C_LIGHT 299792458.0
#endif

namespace b2ext {

static __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
static __device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
static __device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
static __device__ __forceinline__ double2 rmul(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
static __device__ __forceinline__ double2 imul(double2 a) { return make_double2(-a.y, a.x); }      // i*a
static __device__ __forceinline__ double2 nimul(double2 a) { return make_double2(a.y, -a.x); }     // -i*a

#define B2X_2D_INDEX                                                  \
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;             \
    const int iz = blockIdx.y * blockDim.y + threadIdx.y;             \
    if (ir >= Nr || iz >= Nz) return;                                 \
    const size_t o = (size_t)iz * Nr + ir;

// ---- PML split components in spectral space ------------------------------------------------------
// cuda_push_eb_pml_standard / cuda_push_eb_pml_comoving (fbpic/fields/cuda_methods.py:305-331, 415-440;
// CPU twins fbpic/fields/numba_methods.py:189-214, 358-383):
//   E{p,m}_pml <- [T_eb] C E{p,m}_pml + c^2 [T_eb] S_w (-i kr/2 Bz)
//   B{p,m}_pml <- [T_eb] C B{p,m}_pml -     [T_eb] S_w (-i kr/2 Ez)
// Ez, Bz are the values BEFORE the regular push of the same step (spectral_grid.py:343-366 launches this
// kernel first).  T_eb == nullptr selects the standard PSATD.
template <bool COMOVING>
__global__ void k_push_eb_pml(double2 *__restrict__ Ep_pml, double2 *__restrict__ Em_pml,
                              double2 *__restrict__ Bp_pml, double2 *__restrict__ Bm_pml,
                              const double2 *__restrict__ Ez, const double2 *__restrict__ Bz,
                              const double *__restrict__ C, const double *__restrict__ S_w,
                              const double2 *__restrict__ T_eb, const double *__restrict__ kr, int Nz, int Nr) {
    B2X_2D_INDEX
    const double c2 = B2_C_LIGHT * B2_C_LIGHT;
    const double hk = 0.5 * kr[ir];
    double2 TC = make_double2(C[o], 0.), TS = make_double2(S_w[o], 0.);
    if (COMOVING) {
        const double2 t = T_eb[o];
        TC = rmul(C[o], t);
        TS = rmul(S_w[o], t);
    }
    const double2 sE = rmul(c2 * hk, nimul(Bz[o]));     // c^2 (-i kr/2 Bz)
    const double2 sB = rmul(hk, nimul(Ez[o]));          //     (-i kr/2 Ez)
    const double2 dE = cmul(TS, sE), dB = cmul(TS, sB);
    Ep_pml[o] = cadd(cmul(TC, Ep_pml[o]), dE);
    Em_pml[o] = cadd(cmul(TC, Em_pml[o]), dE);
    Bp_pml[o] = csub(cmul(TC, Bp_pml[o]), dB);
    Bm_pml[o] = csub(cmul(TC, Bm_pml[o]), dB);
}

// ---- anisotropic damping in the last n_pml radial cells -----------------------------------------
// cuda_damp_pml_EB (fbpic/boundaries/pml_damping.py:111-154): the PML part of Et, Bt is damped
// (F -= F_pml; F_pml *= d; F += F_pml), Ez and Bz are damped entirely.  Grid: (i_pml, iz).
__global__ void k_damp_pml(double2 *__restrict__ Et, double2 *__restrict__ Et_pml, double2 *__restrict__ Ez,
                           double2 *__restrict__ Bt, double2 *__restrict__ Bt_pml, double2 *__restrict__ Bz,
                           const double *__restrict__ damp, int n_pml, int Nz, int Nr) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    const int iz = blockIdx.y * blockDim.y + threadIdx.y;
    if (ip >= n_pml || iz >= Nz) return;
    const double d = damp[ip];
    const size_t o = (size_t)iz * Nr + (size_t)(Nr - n_pml + ip);
    double2 e = Et[o], ep = Et_pml[o], b = Bt[o], bp = Bt_pml[o];
    e = csub(e, ep);
    b = csub(b, bp);
    ep = rmul(d, ep);
    bp = rmul(d, bp);
    Et_pml[o] = ep;
    Bt_pml[o] = bp;
    Et[o] = cadd(e, ep);
    Bt[o] = cadd(b, bp);
    Ez[o] = rmul(d, Ez[o]);
    Bz[o] = rmul(d, Bz[o]);
}

// ---- cross-deposition current correction ---------------------------------------------------------
// cuda_correct_currents_crossdeposition_standard / _comoving (fbpic/fields/cuda_methods.py:144-171,
// 200-232; CPU twins fbpic/fields/numba_methods.py:88-116, 243-275):
//   Dz  = i kz Jz        + 1/2 a (rho_next - b rho_next_xy + rho_next_z  - b rho_prev)
//   Dxy = kr (Jp - Jm)   + 1/2 a (rho_next + b rho_next_xy - rho_next_z  - b rho_prev)
//   standard: a = 1/dt, b = 1 ; comoving: a = T_cc j_corr_coef, b = T_eb
//   Jp -= Dxy/(2 kr), Jm += Dxy/(2 kr)  (kr != 0) ;  Jz += i Dz / kz  (kz != 0)
// (the standard formula of the reference is the b = 1 case with the xy / z terms written in the other
// order: same value.)
template <bool COMOVING>
__global__ void k_correct_cross(const double2 *__restrict__ rho_prev, const double2 *__restrict__ rho_next,
                                const double2 *__restrict__ rho_next_z, const double2 *__restrict__ rho_next_xy,
                                double2 *__restrict__ Jp, double2 *__restrict__ Jm, double2 *__restrict__ Jz,
                                const double *__restrict__ kz_, const double *__restrict__ kr_,
                                const double2 *__restrict__ T_cc, const double2 *__restrict__ j_corr_coef,
                                const double2 *__restrict__ T_eb, double inv_dt, int Nz, int Nr) {
    B2X_2D_INDEX
    const double kz = kz_[iz], kr = kr_[ir];
    const double2 rp = rho_prev[o], rn = rho_next[o], rz = rho_next_z[o], rxy = rho_next_xy[o];
    double2 tz, txy;      // the rho combinations of Dz and Dxy, times a/2
    if (COMOVING) {
        const double2 a = rmul(0.5, cmul(T_cc[o], j_corr_coef[o]));
        const double2 b = T_eb[o];
        const double2 bxy = cmul(b, rxy), bp = cmul(b, rp);
        tz = cmul(a, csub(cadd(csub(rn, bxy), rz), bp));
        txy = cmul(a, csub(csub(cadd(rn, bxy), rz), bp));
    } else {
        tz = rmul(0.5 * inv_dt, csub(cadd(csub(rn, rxy), rz), rp));
        txy = rmul(0.5 * inv_dt, csub(cadd(csub(rn, rz), rxy), rp));
    }
    const double2 jp = Jp[o], jm = Jm[o], jz = Jz[o];
    const double2 Dz = cadd(rmul(kz, imul(jz)), tz);
    const double2 Dxy = cadd(rmul(kr, csub(jp, jm)), txy);
    if (kr != 0.) {
        const double inv_kr = 1. / kr;
        Jp[o] = cadd(jp, rmul(-0.5 * inv_kr, Dxy));
        Jm[o] = cadd(jm, rmul(0.5 * inv_kr, Dxy));
    }
    if (kz != 0.) {
        const double inv_kz = 1. / kz;
        Jz[o] = cadd(jz, rmul(inv_kz, imul(Dz)));
    }
}

// ---- div E correction ---------------------------------------------------------------------------------
// SpectralGrid.correct_divE (fbpic/fields/spectral_grid.py:299-314; NumPy only in the reference, also in GPU
// runs):  F = -inv_k2 (-rho_prev/eps0 + i kz Ez + kr (Ep - Em));  Ep += kr F/2; Em -= kr F/2; Ez -= i kz F
__global__ void k_correct_divE(double2 *__restrict__ Ep, double2 *__restrict__ Em, double2 *__restrict__ Ez,
                               const double2 *__restrict__ rho_prev, const double *__restrict__ kz_,
                               const double *__restrict__ kr_, const double *__restrict__ inv_k2,
                               double inv_eps0, int Nz, int Nr) {
    B2X_2D_INDEX
    const double kz = kz_[iz], kr = kr_[ir];
    const double2 ep = Ep[o], em = Em[o], ez = Ez[o];
    const double2 div = cadd(cadd(rmul(-inv_eps0, rho_prev[o]), rmul(kz, imul(ez))), rmul(kr, csub(ep, em)));
    const double2 F = rmul(-inv_k2[o], div);
    Ep[o] = cadd(ep, rmul(0.5 * kr, F));
    Em[o] = cadd(em, rmul(-0.5 * kr, F));
    Ez[o] = cadd(ez, rmul(kz, nimul(F)));
}

// ---- momentum push beyond a plane -----------------------------------------------------------------------
// Vay push of particle i (push/inline_functions.py:11-48)
__device__ __forceinline__ void vay_push(long long i, double *__restrict__ ux, double *__restrict__ uy,
                                         double *__restrict__ uz, double *__restrict__ inv_gamma,
                                         const double *__restrict__ Ex, const double *__restrict__ Ey,
                                         const double *__restrict__ Ez, const double *__restrict__ Bx,
                                         const double *__restrict__ By, const double *__restrict__ Bz, double econst,
                                         double bconst) {
    const double tx = bconst * Bx[i], ty = bconst * By[i], tz = bconst * Bz[i];
    const double t2 = tx * tx + ty * ty + tz * tz;
    const double ig0 = inv_gamma[i], u0x = ux[i], u0y = uy[i], u0z = uz[i];
    // half electric kick + full magnetic term with the old velocity
    const double px = u0x + econst * Ex[i] + ig0 * (u0y * tz - u0z * ty);
    const double py = u0y + econst * Ey[i] + ig0 * (u0z * tx - u0x * tz);
    const double pz = u0z + econst * Ez[i] + ig0 * (u0x * ty - u0y * tx);
    const double sigma = 1 + px * px + py * py + pz * pz - t2;
    const double pt = px * tx + py * ty + pz * tz;
    const double ig = sqrt(2. / (sigma + sqrt(sigma * sigma + 4 * (t2 + pt * pt))));
    const double sx = ig * tx, sy = ig * ty, sz = ig * tz, st = ig * pt;
    const double s = 1. / (1 + t2 * ig * ig);
    ux[i] = s * (px + sx * st + py * sz - pz * sy);
    uy[i] = s * (py + sy * st + pz * sx - px * sz);
    uz[i] = s * (pz + sz * st + px * sy - py * sx);
    inv_gamma[i] = ig;
}

// push_p_after_plane_gpu (fbpic/particles/push/cuda_methods.py:103-132): the Vay push
// (push/inline_functions.py:11-48) for the particles with z > z_plane; the others keep their momentum (ballistic
// motion of an injected bunch before it reaches the plasma, injection/ballistic_before_plane.py:10-61).
__global__ void k_push_p_after_plane(long long n, const double *__restrict__ z, double z_plane,
                                     double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
                                     double *__restrict__ inv_gamma, const double *__restrict__ Ex,
                                     const double *__restrict__ Ey, const double *__restrict__ Ez,
                                     const double *__restrict__ Bx, const double *__restrict__ By,
                                     const double *__restrict__ Bz, double econst, double bconst) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !(z[i] > z_plane)) return;
    vay_push(i, ux, uy, uz, inv_gamma, Ex, Ey, Ez, Bx, By, Bz, econst, bconst);
}

// ---- laser antenna: virtual particles -------------------------------------------------------------
// LaserAntenna.deposit_virtual_particles_gpu (fbpic/lpa_utils/laser/antenna_injection.py:357-391):
// positions x = baseline + sign*excursion and normalised momenta u = v/c (sign*v for x, y) of the
// positive (sign = +1) or negative (sign = -1) copy of the antenna particles; inv_gamma is 1 for them.
// The result feeds the regular deposition kernel (order-independent, linear shapes).
__global__ void k_antenna_particles(long long n, const double *__restrict__ bx, const double *__restrict__ by,
                                    const double *__restrict__ ex, const double *__restrict__ ey,
                                    const double *__restrict__ vx, const double *__restrict__ vy,
                                    const double *__restrict__ vz, double sign,
                                    double *__restrict__ x, double *__restrict__ y, double *__restrict__ ux,
                                    double *__restrict__ uy, double *__restrict__ uz) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double inv_c = 1. / B2_C_LIGHT;
    x[i] = bx[i] + sign * ex[i];
    y[i] = by[i] + sign * ey[i];
    ux[i] = sign * vx[i] * inv_c;
    uy[i] = sign * vy[i] * inv_c;
    uz[i] = vz[i] * inv_c;
}

// y[i] += a * x[i]   (LaserAntenna.push_x, antenna_injection.py:196-218; rounded product then rounded
// sum like the cupy expression `y += (dt*push) * x`)
__global__ void k_axpy(long long n, double a, const double *__restrict__ x, double *__restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));
}

// ---- back-transformed (lab-frame) field diagnostics --------------------------------------------
// extract_slice_cuda (fbpic/openpmd_diag/boosted_field_diag.py:745-820): the 10 grid fields of mode m,
// interpolated linearly between the rows iz and iz+1 (weights Sz, 1 - Sz), as REAL rows of
// slice[10][2 Nm - 1][Nr_out]: row 0 = Re(mode 0), rows 2m-1, 2m = 2 Re, 2 Im of mode m > 0.
// Grid: (ir, field).  Rounded products then a rounded sum, like the NumPy expression of the CPU path (:626-629).
struct SliceFields { const double2 *f[10]; };

__global__ void k_extract_slice(const __grid_constant__ SliceFields F, int m, int n_rows, int Nr, int Nr_out, int iz, double Sz,
                                double *__restrict__ slice) {
    const int ir = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (ir >= Nr_out) return;
    const double fac = (m > 0) ? 2. : 1.;
    const double a = fac * Sz, b = fac * (1. - Sz);
    const double2 lo = F.f[k][(size_t)iz * Nr + ir], hi = F.f[k][(size_t)(iz + 1) * Nr + ir];
    const int row = (m > 0) ? 2 * m - 1 : 0;
    double *out = slice + ((size_t)k * n_rows + row) * Nr_out + ir;
    out[0] = __dadd_rn(__dmul_rn(a, lo.x), __dmul_rn(b, hi.x));
    if (m > 0) out[Nr_out] = __dadd_rn(__dmul_rn(a, lo.y), __dmul_rn(b, hi.y));
}

// ParticleCatcher.get_particle_slice (fbpic/openpmd_diag/boosted_particle_diag.py:598-629): the particles that
// crossed the output plane of a lab-frame snapshot during the last cycle.  The plane moved from z_prev to z_curr, a
// particle from z - uz inv_gamma c dt (rounded like the NumPy expression, left to right) to z.  The indices of the
// selected particles are appended to idx (any order; the caller sorts them); count keeps counting beyond cap so that
// the caller can retry with a larger buffer.  The reference downloads a slab of particles found through the cell
// prefix sums (which needs a sorted species) and selects on the host.
__global__ void k_select_crossing(long long n, const double *__restrict__ z, const double *__restrict__ uz,
                                  const double *__restrict__ inv_gamma, double c_light, double dt, double z_curr,
                                  double z_prev, long long cap, long long *__restrict__ idx,
                                  unsigned long long *__restrict__ count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zc = z[i];
    const double zp = __dsub_rn(zc, __dmul_rn(__dmul_rn(__dmul_rn(uz[i], inv_gamma[i]), c_light), dt));
    if ((zc >= z_curr && zp <= z_prev) || (zc <= z_curr && zp >= z_prev)) {
        const unsigned long long pos = atomicAdd(count, 1ULL);
        if ((long long)pos < cap) idx[pos] = i;
    }
}

// ---- ADK ionization (fbpic/particles/elementary_process/ionization) ------------------------------------------
// push_p_ioniz_gpu (push/cuda_methods.py:134-170): the charge of an ionizable macroparticle is level * e; neutral
// ones are not pushed.  econst1, bconst1: the constants of the Vay push for a charge e.
__global__ void k_push_p_ioniz(long long n, const unsigned long long *__restrict__ level, double *__restrict__ ux,
                               double *__restrict__ uy, double *__restrict__ uz, double *__restrict__ inv_gamma,
                               const double *__restrict__ Ex, const double *__restrict__ Ey,
                               const double *__restrict__ Ez, const double *__restrict__ Bx,
                               const double *__restrict__ By, const double *__restrict__ Bz, double econst1,
                               double bconst1) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || level[i] == 0) return;
    const double q = (double)level[i];
    vay_push(i, ux, uy, uz, inv_gamma, Ex, Ey, Ez, Bx, By, Bz, econst1 * q, bconst1 * q);
}

// w_times_level = w * level: the effective weight of the ions in the deposition (ionizer.py:108-109, 
// numba_methods.py:67)
__global__ void k_w_times_level(long long n, const double *__restrict__ w, const unsigned long long *__restrict__ level,
                                double *__restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[i] * (double)level[i];
}

// uniform number in [0, 1) from a counter: splitmix64 of (seed, draw index); one independent stream per (cycle, ion)
__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long counter) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (counter + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1. / 9007199254740992.);
}

// ionize_ions_cuda (ionization/cuda_methods.py:16-71, inline_functions.py:9-45): field amplitude in the rest frame of
// the ion, ADK probability over the proper time of one cycle, one draw per ion; an ionized ion moves up one level and
// its index and former level are appended to `events` (2 entries each; any order, the caller sorts).  `draws`: host
// numbers in [0, 1) (parity tests) or null: counter-based generator.  The reference counts per batch of 10 ions and
// needs a host cumulative sum plus a second pass to place the electrons; here the short event list is all the host
// reads back.
__global__ void k_ionize(long long n, unsigned long long *__restrict__ level, int level_max,
                         const double *__restrict__ adk_prefactor, const double *__restrict__ adk_power,
                         const double *__restrict__ adk_exp_prefactor, const double *__restrict__ ux,
                         const double *__restrict__ uy, const double *__restrict__ uz, const double *__restrict__ Ex,
                         const double *__restrict__ Ey, const double *__restrict__ Ez, const double *__restrict__ Bx,
                         const double *__restrict__ By, const double *__restrict__ Bz, double c_light,
                         const double *__restrict__ draws, unsigned long long seed, long long cap,
                         long long *__restrict__ events, unsigned long long *__restrict__ count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long lv = level[i];
    if (lv >= (unsigned long long)level_max) return;
    const double vx = ux[i], vy = uy[i], vz = uz[i], ex = Ex[i], ey = Ey[i], ez = Ez[i];
    const double bx = c_light * Bx[i], by = c_light * By[i], bz = c_light * Bz[i];
    const double u_dot_E = vx * ex + vy * ey + vz * ez;
    const double gamma = sqrt(1 + vx * vx + vy * vy + vz * vz);
    const double a = gamma * ex + vy * bz - vz * by, b = gamma * ey + vz * bx - vx * bz,
                 d = gamma * ez + vx * by - vy * bx;
    const double E = sqrt(-u_dot_E * u_dot_E + a * a + b * b + d * d);
    if (E == 0) return;
    const double w_dtau = 1. / gamma * adk_prefactor[lv] * pow(E, adk_power[lv]) * exp(adk_exp_prefactor[lv] / E);
    const double p = 1. - exp(-w_dtau);
    const double r = draws ? draws[i] : uniform01(seed, (unsigned long long)i);
    if (r < p) {
        level[i] = lv + 1;
        const unsigned long long pos = atomicAdd(count, 1ULL);
        if ((long long)pos < cap) {
            events[2 * pos] = i;
            events[2 * pos + 1] = (long long)lv;
        }
    }
}

// ---- Compton scattering (fbpic/particles/elementary_process/compton) --------------------------------------------
struct ComptonParams {
    double ct;                  // c * time of the simulation frame
    double photon_n_lab_peak, inv_laser_waist2, inv_laser_ctau2, laser_initial_z0, gamma_boost, beta_boost;
    double photon_p, photon_px, photon_py, photon_pz, photon_beta_x, photon_beta_y, photon_beta_z;   // incoming flux
    double dt, ratio_w_electron_photon, inv_ratio_w_elec_photon;
    double pi_re2, inv_mc, c_light;
    unsigned long long seed;
};

// boost of the 4-momentum (p, px, py, pz) by (gamma, beta) along the unit vector n (compton/inline_functions.py:16-37)
__device__ __forceinline__ void lorentz4(double p, double px, double py, double pz, double gamma, double beta,
                                         double nx, double ny, double nz, double &po, double &pxo, double &pyo,
                                         double &pzo) {
    const double par = nx * px + ny * py + nz * pz;
    po = gamma * (p - beta * par);
    const double par_out = gamma * (par - beta * p);
    pxo = px + nx * (par_out - par);
    pyo = py + ny * (par_out - par);
    pzo = pz + nz * (par_out - par);
}

// Number of photon macroparticles each electron emits during this cycle: photon density of the Gaussian pulse at the
// electron (get_photon_density_gaussian, inline_functions.py:78-112), integrated Klein-Nishina cross-section in the
// rest frame of the electron (get_scattering_probability, :39-76), nscatter = int(p ratio + u) (numba_methods.py:73-88).
// total: sum of nscatter (the host sizes the photon arrays with it).
__global__ void k_compton_count(long long n, const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, const double *__restrict__ ux,
                                const double *__restrict__ uy, const double *__restrict__ uz,
                                const double *__restrict__ inv_gamma, const __grid_constant__ ComptonParams P,
                                int *__restrict__ nscatter, unsigned long long *__restrict__ total) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double zlab = P.gamma_boost * (z[i] + P.beta_boost * P.ct);
    const double ctlab = P.gamma_boost * (P.ct + P.beta_boost * z[i]);
    const double d = zlab - P.laser_initial_z0 + ctlab;
    const double n_lab = P.photon_n_lab_peak *
        exp(-2 * P.inv_laser_waist2 * (x[i] * x[i] + y[i] * y[i]) - 2 * P.inv_laser_ctau2 * d * d);
    const double photon_n = P.gamma_boost * n_lab * (1 + P.beta_boost);
    const double ig = inv_gamma[i];
    const double tf = 1. / ig - ux[i] * P.photon_beta_x - uy[i] * P.photon_beta_y - uz[i] * P.photon_beta_z;
    const double k = P.photon_p * tf * P.inv_mc;
    const double f1 = 2 * (2 + k * (1 + k) * (8 + k)) / (k * k * (1 + 2 * k) * (1 + 2 * k));
    const double f2 = (2 + k * (2 - k)) * log(1 + 2 * k) / (k * k * k);
    const double sigma = P.pi_re2 * (f1 - f2);
    const double p = 1 - exp(-sigma * photon_n * tf * P.c_light * P.dt * ig);
    const int ns = (int)(p * P.ratio_w_electron_photon + uniform01(P.seed, (unsigned long long)i * 65536ULL));
    nscatter[i] = ns;
    if (ns > 0) atomicAdd(total, (unsigned long long)ns);
}

// The scattered photons (scatter_photons_electrons_numba, numba_methods.py:90-264): incoming photon boosted to the
// rest frame of the electron, scattering angle from the Klein-Nishina distribution by rejection sampling, back to
// the simulation frame; the photon starts at the electron; the electron recoils with probability 1 / ratio.  Each
// electron reserves its nscatter slots at the end of the photon arrays with one atomic (ph_* point at the first free
// slot); the draws are counter-based per (electron, draw), so the result does not depend on the order of the atomics.
__global__ void k_compton_scatter(long long n, const int *__restrict__ nscatter, const double *__restrict__ x,
                                  const double *__restrict__ y, const double *__restrict__ z,
                                  double *__restrict__ ux, double *__restrict__ uy, double *__restrict__ uz,
                                  const double *__restrict__ inv_gamma, const double *__restrict__ w,
                                  const __grid_constant__ ComptonParams P, double *__restrict__ ph_x,
                                  double *__restrict__ ph_y, double *__restrict__ ph_z, double *__restrict__ ph_ux,
                                  double *__restrict__ ph_uy, double *__restrict__ ph_uz,
                                  double *__restrict__ ph_inv_gamma, double *__restrict__ ph_w,
                                  unsigned long long *__restrict__ cursor) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ns = nscatter[i];
    if (ns <= 0) return;
    unsigned long long draw = (unsigned long long)i * 65536ULL + 1ULL;
    unsigned long long slot = atomicAdd(cursor, (unsigned long long)ns);
    const double e_ux = ux[i], e_uy = uy[i], e_uz = uz[i], ig = inv_gamma[i];
    const double gamma = 1. / ig;
    const double u = sqrt(e_ux * e_ux + e_uy * e_uy + e_uz * e_uz);
    const double beta = u * ig;
    double nx = 0., ny = 0., nz = 1.;
    if (u != 0) { nx = e_ux / u; ny = e_uy / u; nz = e_uz / u; }
    double rp, rpx, rpy, rpz;
    lorentz4(P.photon_p, P.photon_px, P.photon_py, P.photon_pz, gamma, beta, nx, ny, nz, rp, rpx, rpy, rpz);
    const double cos_t = rpz / rp;
    double sin_t = 0., cos_f = 1., sin_f = 0.;
    if (cos_t * cos_t < 1) {
        sin_t = sqrt(1 - cos_t * cos_t);
        const double inv_pxy = 1. / (sin_t * rp);
        cos_f = rpx * inv_pxy;
        sin_f = rpy * inv_pxy;
    }
    double npx = 0., npy = 0., npz = 0.;
    for (int s = 0; s < ns; ++s) {
        const double k = rp * P.inv_mc;
        const double c0 = 2. * (2. * k * k + 2. * k + 1.) / ((2. * k + 1.) * (2. * k + 1.) * (2. * k + 1.));
        const double b = (2. + c0) / (2. - c0), a = 2. * b - 1.;
        double xs = 1.;
        // (acceptance is above 1/2 for every k; the bound only keeps a NaN input from spinning forever)
        for (int attempt = 0; attempt < 4096; ++attempt) {
            const double r1 = uniform01(P.seed, draw++);
            xs = b - (b + 1.) * pow(0.5 * c0, r1);
            const double h = a / (b - xs);
            const double fac = 1 + k * (1 - xs);
            const double f = ((1 + xs * xs) * fac + k * k * (1 - xs) * (1 - xs)) / (fac * fac * fac);
            if (uniform01(P.seed, draw++) < f / h) break;
        }
        const double new_p = rp / (1 + k * (1 - xs));
        const double sin_s = sqrt(1 - xs * xs);
        const double phi = 2 * 3.141592653589793 * uniform01(P.seed, draw++);
        const double pX = new_p * sin_s * cos(phi), pY = new_p * sin_s * sin(phi), pZ = new_p * xs;
        const double qx = sin_t * cos_f * pZ + cos_t * cos_f * pX - sin_f * pY;
        const double qy = sin_t * sin_f * pZ + cos_t * sin_f * pX + cos_f * pY;
        const double qz = cos_t * pZ - sin_t * pX;
        double np_;
        lorentz4(new_p, qx, qy, qz, gamma, beta, -nx, -ny, -nz, np_, npx, npy, npz);
        ph_x[slot] = x[i]; ph_y[slot] = y[i]; ph_z[slot] = z[i];
        ph_w[slot] = w[i] * P.inv_ratio_w_elec_photon;
        ph_ux[slot] = npx; ph_uy[slot] = npy; ph_uz[slot] = npz;
        ph_inv_gamma[slot] = 1. / np_;
        ++slot;
    }
    // recoil from the last photon, with probability 1 / ratio (numba_methods.py:257-263)
    if (uniform01(P.seed, draw++) < P.inv_ratio_w_elec_photon) {
        ux[i] = e_ux + P.inv_mc * (P.photon_px - npx);
        uy[i] = e_uy + P.inv_mc * (P.photon_py - npy);
        uz[i] = e_uz + P.inv_mc * (P.photon_pz - npz);
    }
}

}  // namespace b2ext

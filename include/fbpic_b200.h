/*
 * fbpic_b200.h -- C ABI of libfbpic_b200.so: the B200 (sm_100a) implementation of
 * FBPIC's per-step PIC hot loop.
 *
 * This is the drop-in boundary.  It replaces the reference's only
 * FFI-like seam, `compile_cupy.__getitem__ -> call_kernel` (fbpic/utils/cuda.py:
 * 434-541), through which every Numba kernel of the reference GPU path is
 * launched, plus the three library calls of that path (cuBLAS dgemm hankel.py:
 * 200-203, cuFFT fourier.py:78/121/153, Thrust argsort cuda_sorting.py:114).
 *
 * Conventions
 *   - every pointer named d_* (or documented "device") is a raw device pointer;
 *     complex128 grids are row-major [Nz][Nr] (z slow, r fast), interleaved re/im;
 *   - every entry point returns 0 on success, else a cudaError_t / cufftResult /
 *     ncclResult_t code (b2_error_string() formats the last failure);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context stream);
 *   - no global mutable state except inside the opaque b2_ctx (one per GPU);
 *   - no host fallback: every function needs a CUDA device.
 * Each declaration cites the reference interface it replaces (file:line relative
 * to the FBPIC source tree).
 */
#ifndef FBPIC_B200_H
#define FBPIC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_ctx b2_ctx;

#define B2_MAX_MODES 8      /* azimuthal modes handled by one launch            */
#define B2_MAX_ARRAYS 32    /* pointers carried by one multi-array launch       */

/* ---- runtime / memory: replaces cupy.asarray/.get and GpuMemoryManager,
 *      fbpic/utils/cuda.py:101-182 ------------------------------------------ */
int b2_device_count(int *count);
int b2_ctx_create(int device, b2_ctx **ctx);
int b2_ctx_destroy(b2_ctx *ctx);
void *b2_ctx_stream(b2_ctx *ctx);
const char *b2_error_string(void);
const char *b2_version(void);
/* device blocks are recycled through a size-keyed free list (B2_POOL_MB caps the parked bytes, default 65536) */
int b2_malloc(void **d_ptr, size_t nbytes);
int b2_free(void *d_ptr);
int b2_host_alloc(void **h_ptr, size_t nbytes);       /* pinned host memory */
int b2_host_free(void *h_ptr);
int b2_memcpy_h2d(void *d_dst, const void *h_src, size_t nbytes, void *stream);
int b2_memcpy_d2h(void *h_dst, const void *d_src, size_t nbytes, void *stream);
int b2_memcpy_d2d(void *d_dst, const void *d_src, size_t nbytes, void *stream);
int b2_memset(void *d_ptr, int value, size_t nbytes, void *stream);
int b2_stream_sync(void *stream);
/* a second stream for work that is off the critical path of the step (the z-FFTs that follow the guard-cell exchange
 * of E and B, boundary_communicator.py / main.py:741-766, are needed only by the next field push); a transform issued
 * on it uses its own scratch.  Order against the context stream with events. */
int b2_stream_create(void **stream);
int b2_stream_destroy(void *stream);
int b2_stream_wait_event(void *stream, void *event);
int b2_device_sync(void);
int b2_event_create(void **event);
int b2_event_destroy(void *event);
int b2_event_record(void *event, void *stream);
int b2_event_elapsed_ms(void *start, void *stop, float *ms);   /* syncs on stop */
/* number of kernels / library calls this library launched since load */
uint64_t b2_launch_count(void);
/* optional per-kernel-family device timing (CUDA events around each launch, on its stream);
 * used by bench.py to time the dominant kernel inside the timed region */
int b2_profile_enable(int on);
int b2_profile_reset(void);
int b2_profile_slots(void);
const char *b2_profile_name(int slot);
int b2_profile_read(int slot, double *total_ms, uint64_t *count);
/* CUDA-graph capture of a sequence of b2_* calls on the context stream */
int b2_graph_begin(b2_ctx *ctx);
int b2_graph_end(b2_ctx *ctx, void **graph_exec);
int b2_graph_launch(b2_ctx *ctx, void *graph_exec);
int b2_graph_destroy(void *graph_exec);

/* ---- particle sorting: get_cell_idx_per_particle (fbpic/particles/utilities/
 *      cuda_sorting.py:22-88), sort_particles_per_cell (:91-122, Thrust argsort),
 *      prefill_prefix_sum + incl_prefix_sum (:125-190), write_sorting_buffer
 *      (:193-213) / Particles.rearrange_particle_arrays (particles.py:510-555).
 *      Contract: sorted_idx == stable argsort(cell_idx); prefix_sum[c] = number of
 *      particles with cell <= c; both bit-exact. ------------------------------- */
int b2_cell_index(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
                  double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
                  int32_t *d_cell_idx, void *stream);
/* d_sorted_idx may be NULL: the permutation then stays inside the context (32-bit) for the
 * immediately following b2_permute(.., NULL, ..) / b2_deposit_permute, and d_cell_idx keeps the
 * unsorted keys */
int b2_sort_cells(b2_ctx *ctx, int64_t n, int32_t *d_cell_idx /* in: keys, out: sorted keys */,
                  int64_t *d_sorted_idx /* out */, int32_t *d_prefix_sum /* out, Nz*(Nr+1) */,
                  int Nz, int Nr, void *stream);
int b2_permute(b2_ctx *ctx, int64_t n, const int64_t *d_sorted_idx, int n_arrays,
               const double *const *d_src /* host array of device pointers */,
               double *const *d_dst, void *stream);

/* ---- gather + push: gather_field_gpu_{linear,cubic}[_one_mode]
 *      (fbpic/particles/gathering/cuda_methods.py:26,209; cuda_methods_one_mode.py:46,216),
 *      push_p_gpu / push_x_gpu (fbpic/particles/push/cuda_methods.py:55,17),
 *      shift_particles_periodic_cuda (fbpic/boundaries/particle_buffer_handling.py:637).
 *      d_grids: host array of 6*Nm device pointers ordered [m][Er,Et,Ez,Br,Bt,Bz]. */
int b2_gather(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
              double rmax_gather, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
              int Nm, const void *const *d_grids, int cubic,
              double *d_Ex, double *d_Ey, double *d_Ez, double *d_Bx, double *d_By, double *d_Bz,
              void *stream);
int b2_push_p(b2_ctx *ctx, int64_t n, double *d_ux, double *d_uy, double *d_uz, double *d_inv_gamma,
              const double *d_Ex, const double *d_Ey, const double *d_Ez,
              const double *d_Bx, const double *d_By, const double *d_Bz,
              double q, double m, double dt, void *stream);
int b2_push_x(b2_ctx *ctx, int64_t n, double *d_x, double *d_y, double *d_z,
              const double *d_ux, const double *d_uy, const double *d_uz, const double *d_inv_gamma,
              double dt, double x_push, double y_push, double z_push, void *stream);
/* fused Particles.gather + push_p + push_x(dt_x) (main.py:470-490): one pass over the SoA */
int b2_gather_push(b2_ctx *ctx, int64_t n, double *d_x, double *d_y, double *d_z,
                   double *d_ux, double *d_uy, double *d_uz, double *d_inv_gamma,
                   double rmax_gather, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr,
                   int Nm, const void *const *d_grids, int cubic,
                   double q, double m, double dt_p, double dt_x,
                   int32_t *d_cell_idx /* NULL, or out: cell key of the new position on a grid starting
                                          at key_zmin (saves the cell-index pass of the next sort) */,
                   double key_zmin, void *stream);
/* push_x(dt) + optional periodic wrap of z into [wrap_zmin, wrap_zmax) + optional cell key */
int b2_push_x_key(b2_ctx *ctx, int64_t n, double *d_x, double *d_y, double *d_z,
                  const double *d_ux, const double *d_uy, const double *d_uz, const double *d_inv_gamma,
                  double dt, int wrap, double wrap_zmin, double wrap_zmax,
                  double invdz, double key_zmin, int Nz, double invdr, double rmin, int Nr,
                  int32_t *d_cell_idx, void *stream);
int b2_shift_periodic(b2_ctx *ctx, int64_t n, double *d_z, double zmin, double zmax, void *stream);
/* particle exchange (remove_particles_cpu / add_buffers, fbpic/boundaries/particle_buffer_handling.py:
 * 58-175, 424-512): stable 3-way partition of the SoA by z -- left if z < zlo, right if z > zhi.
 * classify returns the class sizes {stay, left, right} on the host (synchronises); scatter then
 * writes every attribute to the three destinations (left/right may be NULL: particles dropped). */
int b2_exchange_classify(b2_ctx *ctx, int64_t n, const double *d_z, double zlo, double zhi,
                         int64_t *h_counts3, void *stream);
int b2_exchange_scatter(b2_ctx *ctx, int64_t n, const double *d_z, double zlo, double zhi, int n_arrays,
                        const double *const *d_src, double *const *d_stay, double *const *d_left,
                        double *const *d_right, void *stream);
/* v[i] += value : z-shift of the periodic images received across the ring closure
 * (boundary_communicator.py:815-821) */
int b2_add_scalar(b2_ctx *ctx, int64_t n, double *d_v, double value, void *stream);

/* ---- deposition: deposit_{rho,J}_gpu_{linear,cubic}[_one_mode]
 *      (fbpic/particles/deposition/cuda_methods.py:28,202,466,751; cuda_methods_one_mode.py)
 *      with the boundary folds of fbpic/fields/numba_methods.py:410-461.  Particles
 *      must be cell-sorted (b2_sort_cells + b2_permute); sums are ADDED to the grids
 *      (raw charge, not yet divided by the cell volume).
 *      d_grids: host array of device pointers, rho: [m] ; J: [m][Jr,Jt,Jz].
 *      d_ruyten0 / d_ruyten_hi: Ruyten coefficients (Nr+1) of mode 0 / modes >= 1. */
int b2_deposit_rho(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
                   const double *d_w, double q, double invdz, double zmin, int Nz,
                   double invdr, double rmin, int Nr, int Nm, void *const *d_grids,
                   const int32_t *d_prefix_sum, const double *d_ruyten0, const double *d_ruyten_hi,
                   int cubic, void *stream);
int b2_deposit_J(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
                 const double *d_w, double q, const double *d_ux, const double *d_uy, const double *d_uz,
                 const double *d_inv_gamma, double invdz, double zmin, int Nz,
                 double invdr, double rmin, int Nr, int Nm, void *const *d_grids,
                 const int32_t *d_prefix_sum, const double *d_ruyten0, const double *d_ruyten_hi,
                 int cubic, void *stream);

/* deposition fused with the SoA permutation of the sort that just ran on this context
 * (b2_sort_cells(.., d_sorted_idx=NULL, ..)): reads the 8 UNSORTED attribute arrays through the
 * permutation, writes the 8 sorted arrays and deposits (what: 0 rho, 1 J) in the same pass.
 * d_src8/d_dst8: host arrays of 8 device pointers ordered x,y,z,w,ux,uy,uz,inv_gamma. */
int b2_deposit_permute(b2_ctx *ctx, int what, int64_t n, const double *const *d_src8, double *const *d_dst8,
                       double q, double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                       void *const *d_grids, const int32_t *d_prefix_sum, const double *d_ruyten0,
                       const double *d_ruyten_hi, int cubic, void *stream);
/* rho deposition (linear shapes) for particles that were cell-sorted half a step ago and have
 * moved by about one cell at most since (d_prefix_sum is the OLD sort's): no re-sort needed */
int b2_deposit_rho_displaced(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
                             const double *d_w, double q, double invdz, double zmin, int Nz,
                             double invdr, double rmin, int Nr, int Nm, void *const *d_grids,
                             const int32_t *d_prefix_sum, const double *d_ruyten0, const double *d_ruyten_hi,
                             void *stream);

/* push_x(dt) (+ optional periodic wrap of z) fused with the rho deposition at the NEW position
 * (main.py:519 + 528 in one pass; the grid may have moved in between: zmin is the grid of the
 * deposition) */
int b2_push_deposit_rho(b2_ctx *ctx, int64_t n, double *d_x, double *d_y, double *d_z, const double *d_w,
                        const double *d_ux, const double *d_uy, const double *d_uz, const double *d_inv_gamma,
                        double dt, int wrap, double wrap_zmin, double wrap_zmax, double q,
                        double invdz, double zmin, int Nz, double invdr, double rmin, int Nr, int Nm,
                        void *const *d_grids, const double *d_ruyten0, const double *d_ruyten_hi, int cubic,
                        void *stream);

/* ---- interpolation-grid element-wise ops: cuda_erase_*, cuda_divide_*_by_volume
 *      (fbpic/fields/cuda_methods.py:18-117) ---------------------------------- */
int b2_scale_rows_by_r(b2_ctx *ctx, int n_arrays, void *const *d_arrays, const double *d_invvol,
                       int Nz, int Nr, void *stream);      /* F[iz,ir] *= invvol[ir] */

/* ---- spectral transforms: FFT.transform / inverse_transform
 *      (fbpic/fields/spectral_transform/fourier.py:104-168: cuFFT Z2Z along z of a
 *      [Nz,Nr] array, inverse scaled by 1/Nz) and DHT.transform / inverse_transform
 *      (hankel.py:182-243: out = in @ M as a real [2Nz,Nr]x[Nr,Nr] product), plus the
 *      r,t <-> p,m combinations (spectral_transform/cuda_methods.py:120-158). ------ */
/* inverse: 0 forward, 1 inverse scaled by 1/Nz (NumPy ifft convention), 2 inverse unscaled */
int b2_fft_z(b2_ctx *ctx, const void *d_in, void *d_out, int Nz, int Nr, int inverse, void *stream);
/* batched: n_arrays independent [Nz,Nr] arrays */
int b2_fft_z_multi(b2_ctx *ctx, int n_arrays, const void *const *d_in, void *const *d_out,
                   int Nz, int Nr, int inverse, void *stream);
/* 1 if the library's own two-pass z-FFT (csrc/b2_fft.cu) covers this length with radices <= max_radix, 0 if cuFFT
 * would take it (host-side query, no device work): lets a grid planner pick a guard width with a fast length */
int b2_fft_has_plan(int Nz, int max_radix);
/* out[iz,:] = rowscale[iz] * (in[iz,:] @ M)   (rowscale may be NULL); fp64 DMMA */
int b2_dht(b2_ctx *ctx, const void *d_in, void *d_out, const double *d_M, const double *d_rowscale,
           int Nz, int Nr, void *stream);
/* forward vector transform: p=(r-i t)/2, m=(r+i t)/2 ; out_p = p@Mp, out_m = m@Mm, each
 * row-scaled (spectral_transformer.py:179-223 after the FFTs) */
int b2_dht_rt_to_pm(b2_ctx *ctx, const void *d_r, const void *d_t, void *d_out_p, void *d_out_m,
                    const double *d_Mp, const double *d_Mm, const double *d_rowscale,
                    int Nz, int Nr, void *stream);
/* inverse vector transform: P = p@iMp, Q = m@iMm ; r = P+Q, t = i(P-Q)
 * (spectral_transformer.py:111-155 before the inverse FFTs) */
int b2_dht_pm_to_rt(b2_ctx *ctx, const void *d_p, const void *d_m, void *d_out_r, void *d_out_t,
                    const double *d_iMp, const double *d_iMm, const double *d_rowscale,
                    int Nz, int Nr, void *stream);
/* batched form: the whole list (all modes / components of a field) in one launch per flavour */
enum { B2_DHT_SCALAR = 0, B2_DHT_RT_TO_PM = 1, B2_DHT_PM_TO_RT = 2 };
typedef struct {
    const void *in1, *in2;      /* SCALAR: in1 ; RT_TO_PM: r, t ; PM_TO_RT: p, m          */
    void *out1, *out2;          /* SCALAR: out1 ; RT_TO_PM: p, m ; PM_TO_RT: r, t         */
    const double *M1, *M2;      /* SCALAR: M1 ; vector kinds: the +/- order matrices      */
    const double *rowscale;     /* NULL or [Nz]                                           */
    int kind;
} b2_dht_job;
/* fp64 flops issued by the Hankel GEMMs since load (4*Nz*Nr^2 per transformed array) */
double b2_dht_flops(void);
int b2_dht_batch(b2_ctx *ctx, int njobs, const b2_dht_job *jobs, int Nz, int Nr, void *stream);
int b2_rt_to_pm(b2_ctx *ctx, void *d_r_p, void *d_t_m, int Nz, int Nr, void *stream);   /* in place */
int b2_pm_to_rt(b2_ctx *ctx, void *d_p_r, void *d_m_t, int Nz, int Nr, void *stream);   /* in place */

/* ---- spectral element-wise kernels: cuda_filter_{scalar,vector} (fields/cuda_methods.py:467,492),
 *      cuda_correct_currents_curlfree_{standard,comoving} (:121,174),
 *      cuda_push_eb_{standard,comoving} (:235,334), cuda_push_rho (:443) ----------- */
int b2_filter(b2_ctx *ctx, int n_arrays, void *const *d_arrays, const double *d_filter_z,
              const double *d_filter_r, int Nz, int Nr, void *stream);
typedef struct {
    void *Ep, *Em, *Ez, *Bp, *Bm, *Bz, *Jp, *Jm, *Jz, *rho_prev, *rho_next;   /* complex [Nz,Nr] */
    const double *kz;        /* [Nz] modified kz            */
    const double *kr;        /* [Nr]                        */
    const double *inv_k2;    /* [Nz,Nr] real                */
    const double *C, *S_w;   /* [Nz,Nr] real                */
    const void *j_coef, *rho_prev_coef, *rho_next_coef;  /* real (standard) or complex (comoving) [Nz,Nr] */
    const void *T_eb, *T_cc, *T_rho, *j_corr_coef;       /* complex [Nz,Nr], comoving only */
    double mu_0, epsilon_0;  /* the host's scipy.constants values (CODATA release of the installed SciPy) */
} b2_spectral_mode;
int b2_correct_currents(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double inv_dt,
                        int Nz, int Nr, void *stream);
int b2_push_eb(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double dt, double V,
               int use_true_rho, int Nz, int Nr, void *stream);   /* also does push_rho */
/* fused correct_currents + push_eb + push_rho (fields.py:247-296), one pass */
int b2_correct_push(b2_ctx *ctx, const b2_spectral_mode *mode, int comoving, double dt, double V,
                    int use_true_rho, int Nz, int Nr, void *stream);

/* ---- z boundaries: cuda_damp_EB_left/right (fbpic/boundaries/cuda_methods.py:486,562) and the
 *      halo pack/unpack kernels (:12-483) + BoundaryCommunicator.exchange_domains
 *      (boundary_communicator.py:674-707, mpi4py Isend/Irecv) -> NCCL send/recv ---- */
int b2_damp_z(b2_ctx *ctx, int n_arrays, void *const *d_arrays, const double *d_damp, int nd,
              int left, int right, int Nz, int Nr, void *stream);
/* moving window: F[iz,:] *= shift[iz]^n_move for every listed spectral array
 * (shift_spect_array_gpu, fbpic/boundaries/moving_window.py:244-278); d_shift: complex128[Nz] */
int b2_shift_spect(b2_ctx *ctx, int n_arrays, void *const *d_arrays, const void *d_shift, int n_move,
                   int Nz, int Nr, void *stream);
int b2_add_rows(b2_ctx *ctx, void *d_dst, const void *d_src, int nrows, int Nr, void *stream);
/* Guard-cell staging for the z-slab exchange (replaces the copy_vec_to_gpu_buffer / replace_vec_from_gpu_buffer
 * / add_vec_from_gpu_buffer kernels of fbpic/boundaries/cuda_methods.py:21-253): rows [row0, row0+nrow) of the
 * n_arrays [Nz][Nr] complex arrays <-> one contiguous buffer [n_arrays][nrow][Nr].
 * mode 0: pack, 1: unpack (replace), 2: unpack and add. */
int b2_halo_stage(b2_ctx *ctx, int mode, int n_arrays, void *const *d_arrays, int row0, int nrow, int Nr,
                  void *d_packed, void *stream);
int b2_nccl_unique_id(void *id128);                       /* 128-byte ncclUniqueId */
int b2_nccl_init(b2_ctx *ctx, const void *id128, int rank, int size);
int b2_nccl_destroy(b2_ctx *ctx);
int b2_nccl_group_start(void);
int b2_nccl_group_end(void);
/* group start / end bracketed by profiler events (slot "comm") on the context stream */
int b2_comm_begin(b2_ctx *ctx);
int b2_comm_end(b2_ctx *ctx);
int b2_nccl_send(b2_ctx *ctx, const void *d_buf, size_t nbytes, int peer, void *stream);
int b2_nccl_recv(b2_ctx *ctx, void *d_buf, size_t nbytes, int peer, void *stream);
int b2_nccl_allreduce_max_f64(b2_ctx *ctx, double *d_buf, size_t count, void *stream);

/* ---- solver variants around the hot loop ----------------------------- */
/* radial PML, spectral push of the split components: cuda_push_eb_pml_standard / _comoving
 * (fbpic/fields/cuda_methods.py:305,415; dispatch spectral_grid.py:343-366).  Must run BEFORE the
 * regular E,B push of the step (it reads the old Ez, Bz).  d_T_eb: complex [Nz,Nr] (comoving / Galilean)
 * or NULL (standard PSATD); d_C, d_S_w: real [Nz,Nr]; d_kr: [Nr]. */
int b2_push_eb_pml(b2_ctx *ctx, void *d_Ep_pml, void *d_Em_pml, void *d_Bp_pml, void *d_Bm_pml,
                   const void *d_Ez, const void *d_Bz, const double *d_C, const double *d_S_w,
                   const void *d_T_eb, const double *d_kr, int Nz, int Nr, void *stream);
/* radial PML, anisotropic damping of the last n_pml radial cells in real space: cuda_damp_pml_EB
 * (fbpic/boundaries/pml_damping.py:111-154); d_damp: [n_pml] */
int b2_damp_pml(b2_ctx *ctx, void *d_Et, void *d_Et_pml, void *d_Ez, void *d_Bt, void *d_Bt_pml, void *d_Bz,
                const double *d_damp, int n_pml, int Nz, int Nr, void *stream);
/* cross-deposition current correction: cuda_correct_currents_crossdeposition_standard / _comoving
 * (fbpic/fields/cuda_methods.py:144,200; dispatch spectral_grid.py:231-258).  Uses Jp,Jm,Jz,rho_prev,
 * rho_next,kz,kr (and T_cc, j_corr_coef, T_eb when comoving) of `mode`. */
int b2_correct_currents_cross(b2_ctx *ctx, const b2_spectral_mode *mode, const void *d_rho_next_z,
                              const void *d_rho_next_xy, int comoving, double inv_dt, int Nz, int Nr,
                              void *stream);
/* div E correction in spectral space: SpectralGrid.correct_divE (fbpic/fields/spectral_grid.py:299-314, NumPy
 * only in the reference); uses Ep, Em, Ez, rho_prev, kz, kr, inv_k2, epsilon_0 of `mode` */
int b2_correct_divE(b2_ctx *ctx, const b2_spectral_mode *mode, int Nz, int Nr, void *stream);
/* momentum push only for the particles beyond z_plane: push_p_after_plane_gpu
 * (fbpic/particles/push/cuda_methods.py:103-132); the others move ballistically */
int b2_push_p_after_plane(b2_ctx *ctx, int64_t n, const double *d_z, double z_plane, double *d_ux, double *d_uy,
                          double *d_uz, double *d_inv_gamma, const double *d_Ex, const double *d_Ey,
                          const double *d_Ez, const double *d_Bx, const double *d_By, const double *d_Bz,
                          double q, double m, double dt, void *stream);
/* laser antenna (fbpic/lpa_utils/laser/antenna_injection.py:357-391): positions and normalised momenta of
 * the positive (sign=+1) / negative (sign=-1) copy of the virtual particles, to be handed to
 * b2_deposit_rho / b2_deposit_J (linear shapes, inv_gamma = 1):
 *   x = bx + sign*ex, y = by + sign*ey, ux = sign*vx/c, uy = sign*vy/c, uz = vz/c */
int b2_antenna_particles(b2_ctx *ctx, int64_t n, const double *d_bx, const double *d_by, const double *d_ex,
                         const double *d_ey, const double *d_vx, const double *d_vy, const double *d_vz,
                         double sign, double *d_x, double *d_y, double *d_ux, double *d_uy, double *d_uz,
                         void *stream);
/* y += a*x (LaserAntenna.push_x, antenna_injection.py:196-218) */
int b2_axpy(b2_ctx *ctx, int64_t n, double a, const double *d_x, double *d_y, void *stream);
/* back-transformed diagnostics: extract_slice_cuda (fbpic/openpmd_diag/boosted_field_diag.py:745-820).
 * d_fields10: host array of the 10 device grids Er, Et, Ez, Br, Bt, Bz, Jr, Jt, Jz, rho (complex [Nz, Nr]) of
 * mode m; d_slice: real [10][2 Nm - 1][Nr_out]; rows iz, iz + 1 weighted by Sz, 1 - Sz (x2 for m > 0). */
/* ADK ionization (fbpic/particles/elementary_process/ionization/).
 * b2_push_p_ioniz: push_p_ioniz_gpu (push/cuda_methods.py:134-170), charge = d_level[i] * e, neutral ones skipped.
 * b2_w_times_level: d_out = d_w * d_level, the deposition weight of the ions (ionizer.py:108-109).
 * b2_ionize: ionize_ions_cuda (ionization/cuda_methods.py:16-71): one ADK draw per ion below level_max with the
 *   per-level tables adk_* (ionizer.py:166-183); an ionized ion moves up one level; its index and former level are
 *   appended to d_events (2 x int64 each, capacity cap events, any order).  *h_count: number of events (exact also
 *   when it exceeds cap: call again is NOT possible, the levels have changed -- give cap = n).  d_draws: uniform
 *   numbers in [0, 1), one per ion, or NULL: counter-based generator keyed by (seed, ion index).
 *   The call synchronises the stream. */
int b2_push_p_ioniz(b2_ctx *ctx, int64_t n, const uint64_t *d_level, double *d_ux, double *d_uy, double *d_uz,
                    double *d_inv_gamma, const double *d_Ex, const double *d_Ey, const double *d_Ez,
                    const double *d_Bx, const double *d_By, const double *d_Bz, double m, double dt, void *stream);
int b2_w_times_level(b2_ctx *ctx, int64_t n, const double *d_w, const uint64_t *d_level, double *d_out, void *stream);
int b2_ionize(b2_ctx *ctx, int64_t n, uint64_t *d_level, int level_max, const double *d_adk_prefactor,
              const double *d_adk_power, const double *d_adk_exp_prefactor, const double *d_ux, const double *d_uy,
              const double *d_uz, const double *d_Ex, const double *d_Ey, const double *d_Ez, const double *d_Bx,
              const double *d_By, const double *d_Bz, const double *d_draws, uint64_t seed, int64_t cap,
              int64_t *d_events, int64_t *d_count, int64_t *h_count, void *stream);
/* Compton scattering of a counter-propagating Gaussian laser pulse off an electron species
 * (fbpic/particles/elementary_process/compton/).  params20 (host): ct, photon_n_lab_peak, inv_laser_waist2,
 * inv_laser_ctau2, laser_initial_z0, gamma_boost, beta_boost, photon_p, photon_px, photon_py, photon_pz,
 * photon_beta_x, photon_beta_y, photon_beta_z, dt, ratio_w_electron_photon, 1/ratio, pi r_e^2, 1/(m_e c), c.
 * b2_compton_count: determine_scatterings_* (numba_methods.py:50-88 with get_photon_density_gaussian and
 *   get_scattering_probability): d_nscatter[i] photons for electron i, their sum in *h_total (synchronises).
 * b2_compton_scatter: scatter_photons_electrons_* (numba_methods.py:90-264): writes the sum(d_nscatter) photons
 *   (x, y, z, ux, uy, uz = momentum in kg m/s, inv_gamma = 1/|p|, w) from the first free slot photon8[k] on
 *   (host array of 8 device pointers in the order x y z ux uy uz inv_gamma w) and applies the electron recoil.
 *   Same seed in both calls; draws are counter-based per (electron, draw). */
int b2_compton_count(b2_ctx *ctx, int64_t n, const double *d_x, const double *d_y, const double *d_z,
                     const double *d_ux, const double *d_uy, const double *d_uz, const double *d_inv_gamma,
                     const double *params20, uint64_t seed, int32_t *d_nscatter, int64_t *d_total, int64_t *h_total,
                     void *stream);
int b2_compton_scatter(b2_ctx *ctx, int64_t n, const int32_t *d_nscatter, const double *d_x, const double *d_y,
                       const double *d_z, double *d_ux, double *d_uy, double *d_uz, const double *d_inv_gamma,
                       const double *d_w, const double *params20, uint64_t seed, double *const *photon8,
                       int64_t *d_cursor, void *stream);
/* lab-frame particle output: ParticleCatcher.get_particle_slice (fbpic/openpmd_diag/boosted_particle_diag.py:598-629).
 * Appends to d_idx (capacity cap) the indices of the particles for which
 *   (z >= z_curr and z_old <= z_prev) or (z <= z_curr and z_old >= z_prev),  z_old = z - uz inv_gamma c dt,
 * in any order, and returns their number in *h_count (host; the call synchronises the stream).  A count above cap
 * means that only cap indices were stored: call again with a larger buffer.  d_count: 8 bytes of device scratch. */
int b2_select_crossing(b2_ctx *ctx, int64_t n, const double *d_z, const double *d_uz, const double *d_inv_gamma,
                       double c_light, double dt, double z_curr, double z_prev, int64_t cap, int64_t *d_idx,
                       int64_t *d_count, int64_t *h_count, void *stream);
int b2_extract_slice(b2_ctx *ctx, const void *const *d_fields10, int m, int Nm, int Nz, int Nr, int Nr_out, int iz,
                     double Sz, double *d_slice, void *stream);
/* external fields (ExternalField, fbpic/lpa_utils/external_fields.py:13-215): the reference turns the user's
 * Python function into a GPU kernel with Numba (:134-147); here its body arrives as CUDA C statements over
 * the scalars F, x, y, z, t, amplitude, length_scale that end with `F_[i_] = <expr>;`, is compiled once by
 * NVRTC to an sm_100a cubin (no GPU needed for this step) and applied element-wise to one gathered field of
 * a species.  gamma_boost/beta_boost: the expression is evaluated at the lab-frame (z, t) of the particle
 * (:118-126); pass 1, 0 in the lab frame. */
int b2_external_field_compile(const char *cuda_body, void **handle);
int b2_external_field_cubin_size(void *handle, size_t *nbytes);
int b2_external_field_apply(b2_ctx *ctx, void *handle, int64_t n, double *d_F, const double *d_x,
                            const double *d_y, const double *d_z, double t, double amplitude,
                            double length_scale, double gamma_boost, double beta_boost, void *stream);
int b2_external_field_free(void *handle);

#ifdef __cplusplus
}
#endif
#endif

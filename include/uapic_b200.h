/*
 * uapic_b200.h -- C ABI of libuapic_b200.so: the UA-PIC time step of JuliaVlasov/UAPIC.jl
 * (fortran/bupdate.F90, test/bupdate.jl) as hand-written sm_100a CUDA.
 *
 * The reference has no FFI boundary of its own for this path: the boundary it offers is the
 * exported Julia function set.  Every entry point below names the Julia
 * function (and the Fortran routine it shadows) that a maintainer would rebind to it with
 * `ccall`; INTEGRATION.md shows those bindings.
 *
 * Conventions
 *  - all arrays are COLUMN-MAJOR host buffers exactly as Julia/Fortran hold them:
 *      ComplexF64 (ntau,2,nbpart)  -> `double*` with interleaved (re,im), tau fastest
 *      Float64    (ntau,2,nbpart), (2,nbpart), (2,nx+1,ny+1), (nx+1,ny+1)
 *  - every function returns 0 on success, a negative UAPIC_E* code otherwise;
 *    uapic_last_error() gives the message of the last failure on the calling thread.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    UAPIC_ENODEVICE.
 *  - ntau: any even number in [2, 256], as in the reference (ua_type.F90:32-76).  Powers of two <= 32 run on the fast kernels
 *    (one tau sample per lane; 8, 16, 32 also on the one-pass kernels); every other even ntau runs on general kernels
 *    (one warp per particle, direct DFTs) behind the same entry points -- complete, not fast; sessions need UAPIC_STORE_FULL.
 *  - stage functions are synchronous (H2D, kernel, D2H inside the call); the session API keeps
 *    all state resident in HBM and is the performance path.
 */
#ifndef UAPIC_B200_H
#define UAPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAPIC_OK            0
#define UAPIC_EINVAL       -1   /* bad argument (ntau not a power of two, null pointer, ...) */
#define UAPIC_ENODEVICE    -2   /* no CUDA device / wrong architecture */
#define UAPIC_ECUDA        -3   /* a CUDA runtime call or kernel failed */
#define UAPIC_ENOMEM       -4   /* device allocation failed */
#define UAPIC_ESTATE       -5   /* session used in the wrong order */
#define UAPIC_EUNSUPPORTED -6   /* valid request this build does not implement */

/* periodic wrap convention */
#define UAPIC_WRAP_FORTRAN 0    /* px = x/dx; px = modulo(px,nx); stored x unwrapped  (compute_rho_m6.F90:86-93) */
#define UAPIC_WRAP_JULIA   1    /* x = mod(x-xmin,dimx); px = x/dx; stored x wrapped   (src/compute_rho.jl:63-70) */

/* charge accumulation */
#define UAPIC_DEPOSIT_FP64_ATOMIC 0   /* fp64 atomics: fastest, summation order not reproducible */
#define UAPIC_DEPOSIT_FIXED_POINT 1   /* int64 fixed point: bit-identical run to run and for any GPU count */

/* shape functions */
#define UAPIC_SCHEME_M6  0      /* quintic spline, the scheme the reference ships (compute_rho_m6.F90:28-45) */
#define UAPIC_SCHEME_CIC 1      /* bilinear; BUILD-DEFINED (the reference has no 2D CIC deposit and no CIC in the UA loop): weights of
                                   performance/test_cic.F90:73-76 with the wrap, ghost copy, scaling and neutralisation of the M6
                                   path; session API with UAPIC_STORE_ONEPASS_LEAN only; parity is against oracle/, not the reference */

/* what crosses the intra-step barrier */
#define UAPIC_STORE_FULL   0    /* 128 B per particle-tau kept in HBM between predictor and corrector */
#define UAPIC_STORE_HYBRID 1    /* 16 B per particle-tau (E at the tau samples); predictor recomputed */
/* one-pass modes (ntau = 8, 16, 32): the corrector deposit does not depend on the predictor field, so both deposits of a
   step come out of ONE kernel and one field barrier per step remains; compute_v needs no tau-FFT */
#define UAPIC_STORE_ONEPASS      2   /* 72 B per particle-tau: Re xt_pred, yt_pred, W_n, interv */
#define UAPIC_STORE_ONEPASS_LEAN 3   /* 48 B per particle-tau: Re xt_pred, yt_pred; W_n and interv recomputed */

typedef struct uapic_mesh {
    double  xmin, xmax, ymin, ymax;   /* src/meshfields.jl:5-12, fortran/meshfields.F90:7-14 */
    int32_t nx, ny;
} uapic_mesh_t;

/* message of the last error raised on this thread ("" if none) */
const char *uapic_last_error(void);
/* library version, and the SM architecture the kernels were compiled for (100 = sm_100a) */
int uapic_version(void);
int uapic_compiled_arch(void);
/* 2^S used by UAPIC_DEPOSIT_FIXED_POINT for a given total deposited mass (nbpart_global*w): every tap is rounded to a
   multiple of 2^-S and accumulated in int64, so sums are order independent (pure host arithmetic, no device needed) */
int uapic_fixed_point_scale(double total_mass, double *scale);
/* number of usable CUDA devices (0 and UAPIC_ENODEVICE when none) */
int uapic_device_count(int *count);
/* measurement aid (bench.py --peaks): DFMA instructions per second the whole chip sustains in a pure fused-multiply-add
   loop (8 independent chains per thread), over `launches` back-to-back launches timed with CUDA events.  The fp64-pipe
   roofline of SURVEY.md section 8d is 2 flop x this number; nothing on the product path calls it. */
int uapic_probe_fp64_peak(int device, int launches, double *dfma_per_s, double *ms_per_launch);

/* ------------------------------------------------------------------------------------------
 * Stage API: one entry point per exported Julia function on the hot path.
 * ------------------------------------------------------------------------------------------ */

/* compute_rho_m6!(fields, particles)            src/compute_rho.jl:181-316, compute_rho_m6.F90:205-335
   x (2,nbpart) is rewritten in place only for UAPIC_WRAP_JULIA.  rho (nx+1,ny+1) out.  rho_total may be NULL. */
int uapic_compute_rho_m6(const uapic_mesh_t *mesh, int64_t nbpart, double *x, double w, double *rho,
                         int wrap, int deposit_mode, double *rho_total);

/* interpol_eb_m6!(particles, fields)            src/interpolation.jl:125-247, interpolation_m6.F90:193-327
   e (2,nx+1,ny+1) in, ep (2,nbpart) out; x rewritten only for UAPIC_WRAP_JULIA. */
int uapic_interpol_eb_m6(const uapic_mesh_t *mesh, const double *e, int64_t nbpart, double *x, double *ep, int wrap);

/* The same two stages with the bilinear (CIC) shape of UAPIC_SCHEME_CIC -- build-defined: the reference has no 2D CIC deposit
   (SURVEY.md 2.4); weights of performance/test_cic.F90:73-76 = the 2D restriction of compute_rho_cic.f90:46-53, with the wrap,
   ghost copy, 1/(dx dy) and neutralisation of the M6 routines.  Same arguments as the M6 entry points. */
int uapic_compute_rho_cic(const uapic_mesh_t *mesh, int64_t nbpart, double *x, double w, double *rho,
                          int wrap, int deposit_mode, double *rho_total);
int uapic_interpol_eb_cic(const uapic_mesh_t *mesh, const double *e, int64_t nbpart, double *x, double *ep, int wrap);

/* (p::Poisson)(fields)                          src/poisson.jl:62-83, poisson_2d.f90:85-111
   rho (nx+1,ny+1) in, e (2,nx+1,ny+1) out, *energy = sum(e1^2+e2^2)*dx*dy over the ghosted array. */
int uapic_poisson(const uapic_mesh_t *mesh, const double *rho, double *e, double *energy);

/* preparation!(ua, dt, particles, xt, yt)       src/ua_steps.jl:3-78, ua_steps.F90:15-115
   in: x,v,e (2,nbpart).  out: b,t (nbpart); pl,ql ComplexF64 (ntau,nbpart); xt,yt ComplexF64 (ntau,2,nbpart). */
int uapic_preparation(int ntau, double eps, double dt, int64_t nbpart, const double *x, const double *v,
                      const double *e, double *b, double *t, double *pl, double *ql, double *xt, double *yt);

/* update_particles_e!  = interpol_eb_m6!(et, fields, xt, nbpart, ntau)
                                                 src/ua_steps.jl:82-90, src/interpolation.jl:3-123, interpolation_m6.F90:40-191 */
int uapic_interpol_eb_m6_tau(const uapic_mesh_t *mesh, const double *e, int ntau, int64_t nbpart,
                             const double *xt, double *et, int wrap);

/* compute_f!(fx, fy, ua, particles, xt, yt, et) src/ua_steps.jl:105-145, ua_steps.F90:140-198
   normalise = 0: Julia (unnormalised fft!); 1: Fortran (fft then /ntau). */
int uapic_compute_f(int ntau, double eps, int64_t nbpart, const double *b, const double *xt, const double *yt,
                    const double *et, double *fx, double *fy, int normalise);

/* mul!(x̃t, ftau, xt) / ifft!(xt,1)              test/bupdate.jl:79,82,85-86,102
   nvec length-ntau complex vectors; sign -1 forward, +1 backward; normalise divides by ntau. */
int uapic_fft_tau(int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise);

/* ua_step!(xt, x̃t, ua, particles, fx)           src/ua_steps.jl:149-170   (Fourier in, Fourier out) */
int uapic_ua_step_predict(int ntau, double eps, int64_t nbpart, const double *t, const double *pl,
                          const double *xf, const double *fx, double *xt);
/* ua_step!(xt, x̃t, ua, particles, fx, gx)       src/ua_steps.jl:172-200 */
int uapic_ua_step_correct(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql,
                          const double *xf, const double *fx, const double *gx, double *xt);

/* Fortran forms: ua_step1 / ua_step2            ua_steps.F90:200-236, 238-272
   ua_step1: xf <- FFT(xt); xt <- IFFT(elt/ntau*xf + pl*fx)      (fx normalised by 1/ntau)
   ua_step2: xt <- IFFT(elt/ntau*xf + pl*fx + ql*(gx-fx)/t) */
int uapic_ua_step1(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, double *xt, double *xf,
                   const double *fx);
int uapic_ua_step2(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql,
                   double *xt, const double *xf, const double *fx, const double *gx);

/* update_particles_x! = compute_rho_m6!(fields, particles, xt, ua)
                                                 src/ua_steps.jl:94-101, src/compute_rho.jl:29-179, compute_rho_m6.F90:47-203
   xt time-domain ComplexF64 (ntau,2,nbpart), t (nbpart) in; rho (nx+1,ny+1) and x (2,nbpart) out. */
int uapic_compute_rho_m6_tau(const uapic_mesh_t *mesh, int ntau, double eps, int64_t nbpart, const double *xt,
                             const double *t, double w, double *rho, double *x, int wrap, int deposit_mode,
                             double *rho_total);

/* compute_v!(yt, particles, ua)                 src/ua_steps.jl:204-224 (yt_is_fourier = 1)
   compute_v(ua, particles, yt, yf)              ua_steps.F90:274-307    (yt_is_fourier = 0: FFT first) */
int uapic_compute_v(int ntau, double eps, int64_t nbpart, const double *t, const double *yt, int yt_is_fourier,
                    double *v);

/* ------------------------------------------------------------------------------------------
 * Session API: the whole loop of fortran/bupdate.F90:89-128 / test/bupdate.jl:63-114 with all
 * state resident in HBM.  One session per GPU (one process per GPU); the only inter-GPU step is
 * the sum of the raw rho mesh, delegated to a caller-supplied collective.
 * ------------------------------------------------------------------------------------------ */

typedef struct uapic_session uapic_session_t;

typedef struct uapic_config {
    uapic_mesh_t mesh;
    int32_t ntau;            /* even, 2..256; powers of two <= 32 are the fast ones */
    int32_t wrap;            /* UAPIC_WRAP_*          */
    int32_t deposit_mode;    /* UAPIC_DEPOSIT_*       */
    int32_t scheme;          /* UAPIC_SCHEME_*        */
    int32_t storage_mode;    /* UAPIC_STORE_*         */
    int32_t device;          /* CUDA ordinal          */
    double  eps;             /* bupdate.F90:18        */
    double  dt;              /* bupdate.F90:66        */
    int64_t nbpart;          /* particles held by THIS session (its shard) */
    double  weight;          /* particles%w = dimx*dimy/nbpart_global (particles.F90:52) */
    double  total_mass;      /* nbpart_global*weight: bounds the fixed-point scale; 0 -> dimx*dimy */
    void   *stream;          /* cudaStream_t to run on (NULL = the legacy default stream) */
} uapic_config_t;

/* sum `count` elements at device pointer `buf` over all ranks, in place, ordered on `stream`.
   dtype: 0 = float64, 1 = int64.  Return 0 on success. */
typedef int (*uapic_allreduce_fn)(void *ctx, void *buf, int64_t count, int dtype, void *stream);

int uapic_session_create(const uapic_config_t *cfg, uapic_session_t **out);
int uapic_session_destroy(uapic_session_t *s);
int uapic_session_set_allreduce(uapic_session_t *s, uapic_allreduce_fn fn, void *ctx);

/* In-library NCCL: the library binds libnccl.so.2 at run time (dlopen; $UAPIC_NCCL_LIB overrides the
   name) and enqueues ncclAllReduce(sum) of the raw rho meshes on the session's stream itself -- no host callback, so a Julia or
   C caller gets the multi-GPU path with two calls and the step stays capturable in a CUDA graph.
   One process per GPU: rank 0 calls uapic_nccl_unique_id and hands the 128 bytes to the other ranks by any means it has (MPI,
   a file, torch.distributed); every rank then calls uapic_session_init_nccl (collective: ncclCommInitRank on the session's
   device).  The communicator is destroyed with the session.  uapic_session_set_nccl_comm adopts an existing ncclComm_t
   instead (not destroyed by the session; NULL detaches).  Takes precedence over uapic_session_set_allreduce. */
/* Peer-memory exchange over NVLink / NVSwitch: NO collective call.  Every rank folds its deposits into an exchange buffer that
   the other ranks of the node map (CUDA IPC); a step then publishes a counter with st.release.sys and the field-solve kernel of
   every rank waits for all counters (ld.acquire.sys over NVLink) and adds the ranks' buffers in rank order in its first phase --
   the all-reduce is fused into the solve, the sum has the same bits on every rank (fp64 included), and a step is 6 launches.
   Protocol: every rank calls uapic_session_peer_handle, the 64-byte handles are gathered by any means, every rank calls
   uapic_session_init_peers with all of them (rank order); uapic_session_close_peers before destroying (all ranks, after a
   barrier of the caller's).  Same node only; at most 16 ranks.  Takes precedence over NCCL and the callback. */
#define UAPIC_PEER_HANDLE_BYTES 64
int uapic_session_peer_handle(uapic_session_t *s, void *handle64);
int uapic_session_init_peers(uapic_session_t *s, const void *handles, int nranks, int rank);
int uapic_session_close_peers(uapic_session_t *s);
#define UAPIC_NCCL_ID_BYTES 128
int uapic_nccl_unique_id(void *id128);
int uapic_nccl_version(int *version);
int uapic_session_init_nccl(uapic_session_t *s, const void *id128, int nranks, int rank);
int uapic_session_set_nccl_comm(uapic_session_t *s, void *nccl_comm);

/* x, v (2,nbpart) host (pageable or pinned) -> device SoA */
int uapic_session_upload_particles(uapic_session_t *s, const double *x, const double *v);
/* particles.e (2,nbpart), the field at the particles frozen after init (bupdate.F90:93): lets a caller that keeps
   the particles on the host hand the full per-step input (x, v, e) back to a session */
int uapic_session_upload_particle_e(uapic_session_t *s, const double *ep);
/* spatial reordering of the particle arrays by coarse mesh bin (2^bin_cells_log2 cells per side) every `interval` steps;
   0 switches it off.  It only changes which particles a warp works on together (L1 hit rate of the M6 gathers): uploads
   and downloads are always in the caller's particle order.  Default: every step with 8 x 8-cell bins for the one-pass
   storage modes, off for the others.  Not in the reference (particles there are processed in array order). */
int uapic_session_set_sort(uapic_session_t *s, int interval, int bin_cells_log2);
/* Phase fusion (UAPIC_STORE_ONEPASS_LEAN): phase B of step n (gather of the predictor field, compute_v) runs inside the first
   kernel of step n+1 instead of as a kernel of its own; results are identical to the unfused sequence (same arithmetic, same
   order).  The pending phase B of the last step runs on its own as soon as something needs v (downloads, sum_v) or reorders the
   particles; with fusion on, reordering happens every 4 steps by default (uapic_session_set_sort changes it). */
int uapic_session_set_fusion(uapic_session_t *s, int enable);
/* per-kernel device timing: when enabled, CUDA events bracket the two fused phase kernels of every step;
   phase_times returns the accumulated milliseconds and the number of steps they cover, then resets them */
int uapic_session_enable_timing(uapic_session_t *s, int enable);
int uapic_session_phase_times(uapic_session_t *s, double *ms_phase_a, double *ms_phase_b, int64_t *steps);
/* milliseconds between the end of phase A and the start of phase B (fold of the deposit copies + all-reduce + field solve),
   accumulated over the steps the LAST uapic_session_phase_times call reported (one-pass modes; the exchange step of SURVEY 8e) */
int uapic_session_field_barrier_time(uapic_session_t *s, double *ms_field_barrier);
/* device-side loaders (counter-based RNG; particle index offset = first global index of this shard)
   kind 0: init_particles_2d / plasma densities (particles.F90:68-103, src/plasma.jl:17-48)
   kind 1: Landau load (the intent of src/landau.jl:19-43)                                         */
int uapic_session_generate_particles(uapic_session_t *s, int kind, uint64_t seed, int64_t first_global_index,
                                     double alpha, double kx);
/* same, with the shard holding the global indices first, first+stride, first+2*stride, ...  The Landau load assigns |v| by
   particle index (src/landau.jl:35): contiguous index ranges would give every rank a different velocity band (and the rank
   with the fast particles, whose gyro-orbits span more mesh lines, would set the pace), interleaved ones do not. */
int uapic_session_generate_particles_strided(uapic_session_t *s, int kind, uint64_t seed, int64_t first_global_index,
                                             int64_t index_stride, double alpha, double kx);
/* compute_rho_m6_real -> solve_poisson -> interpolate_eb_m6_real     bupdate.F90:89-93 */
int uapic_session_init_fields(uapic_session_t *s);
/* nsteps iterations of the loop body bupdate.F90:97-123 (dead third interpolation skipped) */
int uapic_session_step(uapic_session_t *s, int nsteps);
/* One UA step with the particles living in HOST memory: x_in, v_in (2,nbpart) and e_in (2,nbpart; NULL keeps the
   particles.e held on the device) are copied up, one step runs, x_out and v_out (2,nbpart) are copied down -- the same
   result as upload_particles + upload_particle_e + step(1) + download_particles, but pipelined: the particle range is cut
   into chunks whose copies overlap the kernels of the neighbouring chunks.  Use pinned host buffers (pageable memory works
   but serialises).  Returns when x_out and v_out are complete.  Caller side: what a Julia program that keeps `particles`
   on the host calls once per iteration of the loop test/bupdate.jl:69-114. */
int uapic_session_step_host(uapic_session_t *s, const double *x_in, const double *v_in, const double *e_in, double *x_out,
                            double *v_out);
/* block until everything queued on the session's stream has finished */
int uapic_session_synchronize(uapic_session_t *s);

int uapic_session_download_particles(uapic_session_t *s, double *x, double *v);
int uapic_session_download_particle_e(uapic_session_t *s, double *ep);
int uapic_session_download_fields(uapic_session_t *s, double *e, double *rho);
/* electric energy after every Poisson solve so far: 1 + 2*steps values (test/bupdate.jl:65,90,106) */
int uapic_session_energy_history(uapic_session_t *s, double *out, int64_t capacity, int64_t *count);
/* sum(v[1,:]), sum(v[2,:]) of this shard, the numbers bupdate prints every step (bupdate.F90:125) */
int uapic_session_sum_v(uapic_session_t *s, double *sumv2);
/* kernels launched by this session so far (for bench.py's gpu_launches) */
int uapic_session_launch_count(uapic_session_t *s, int64_t *count);
/* bytes of HBM the session holds */
int uapic_session_device_bytes(uapic_session_t *s, int64_t *bytes);

/* ------------------------------------------------------------------------------------------
 * Sibling scheme: the 3D rotation-push PIC of fortran/uapic3d.f90 --
 * CIC deposition (compute_rho_cic.f90:11-79), 3D periodic spectral Poisson solve (poisson_3d.f90:47-191),
 * CIC interpolation (interpolation_cic.f90:10-66), the velocity rotation in the field b(x) of
 * uapic3d.f90:103-121 and its multi-revolution composition (:131-204).  Arrays are column-major as the
 * Fortran holds them: x, v, e_particles (3,nbpart); rho (nx+1,ny+1,nz+1); e (3,nx+1,ny+1,nz+1).
 * The reference's quirks on this path are reproduced (uapic_mrc3d.cu lists them); no CPU fallback.
 * ------------------------------------------------------------------------------------------ */
typedef struct uapic3d_mesh {
    double  xmin[3], xmax[3];        /* meshfields.F90:80-107 */
    int32_t n[3];
} uapic3d_mesh_t;

typedef struct uapic3d_config {
    uapic3d_mesh_t mesh;
    int64_t nbpart;                  /* particles of THIS session */
    int64_t nbpart_global;           /* 0 = nbpart (bounds the fixed-point scale) */
    double  weight;                  /* p%w = dimx*dimy*dimz/nbpart_global (particles.F90:133) */
    double  ep;                      /* uapic3d.f90:19 (0.5**10) */
    double  delta;                   /* uapic3d.f90:18 (3e-3): b(x) = (d2, -d1, 1)*delta-scaled / sqrt(1 + r^2 delta^2), centre (9, 9) */
    int32_t deposit_mode;            /* UAPIC_DEPOSIT_* */
    int32_t index_quirk;             /* 1: reproduce p%x(m,1) of uapic3d.f90:179,182 (the reference's behaviour); 0: x(1,m) */
    int32_t device;
    void   *stream;
} uapic3d_config_t;

typedef struct uapic3d_session uapic3d_session_t;

int uapic3d_create(const uapic3d_config_t *cfg, uapic3d_session_t **out);
int uapic3d_destroy(uapic3d_session_t *s);
int uapic3d_upload_particles(uapic3d_session_t *s, const double *x, const double *v);
/* particles sharded over ranks (one process per GPU): the library sums the raw rho nodes with ncclAllReduce before the periodic
   copies, every rank solves the field redundantly (same scheme as uapic_session_init_nccl; id from uapic_nccl_unique_id) */
int uapic3d_init_nccl(uapic3d_session_t *s, const void *id128, int nranks, int rank);
/* init_particles_3d densities (particles.F90:152-190) from a counter-based stream keyed by first_global_index + k */
int uapic3d_generate_particles(uapic3d_session_t *s, uint64_t seed, int64_t first_global_index);
/* compute_rho_cic -> solve_poisson -> interpolate_eb_cic                         uapic3d.f90:76-83 */
int uapic3d_init_fields(uapic3d_session_t *s);
/* `count` sub-steps: push(dt/2) -> deposit -> Poisson -> interpolate -> rotate -> push(dt/2).
   kind 0: uapic3d.f90:93-127 (coef unused); kind 1: :133-165, coef = alpha; kind 2: :167-200, coef = beta */
int uapic3d_substep(uapic3d_session_t *s, int kind, double dt, double coef, int count);
/* the time loop of uapic3d.f90:44-61, :91-206 (branch on N0mrc = nint(tfinal/ep/2pi/Nmrc)); max_outer > 0 caps the number of
   outer iterations (steps of the first branch, istep of the MRC branch); *substeps = sub-steps performed */
int uapic3d_run(uapic3d_session_t *s, int nmrc, int nmrcm, double tfinal, int max_outer, int64_t *substeps);
int uapic3d_download_particles(uapic3d_session_t *s, double *x, double *v, double *e_particles);
int uapic3d_download_fields(uapic3d_session_t *s, double *e, double *rho);
int uapic3d_launch_count(uapic3d_session_t *s, int64_t *count);
/* stage functions (host buffers, synchronous): the three module routines the Fortran tests call (test_pic_3d.f90, test_poisson_3d.f90) */
int uapic3d_compute_rho_cic(const uapic3d_mesh_t *mesh, int64_t nbpart, const double *x, double w, double *rho);
int uapic3d_poisson(const uapic3d_mesh_t *mesh, const double *rho, double *e);
int uapic3d_interpolate_eb_cic(const uapic3d_mesh_t *mesh, const double *e, int64_t nbpart, const double *x, double *e_particles);

/* ------------------------------------------------------------------------------------------
 * Sibling scheme: the external-field two-scale program of fortran/efd.f90 (test/test_efd.jl is its Julia twin).  Every
 * particle is integrated on its own in the prescribed field of efd.f90:166-167 with b(x) of efd.f90:139: third-order prepared
 * datum (efd.f90:157-383), nstep IMEX2 steps in tau-Fourier space (:388-454), state read off at tau = tfinal b/eps and
 * wrapped into the box (:456-478, :526-544).  This is the text behind the program's `stop` (efd.f90:257) over ALL particles --
 * the run whose sum(v) the program prints against on efd.f90:481.  One kernel launch; x, v, x_out, v_out are (2,nbpart),
 * column-major; the outputs may alias the inputs.  ntau: any even value 2..256.
 * ------------------------------------------------------------------------------------------ */
typedef struct uapic_efd_config {
    int32_t ntau;                    /* efd.f90:18 (16) */
    int32_t nstep;                   /* 0 = nint(tfinal/dt), efd.f90:102 */
    double  eps;                     /* efd.f90:106 (1e-3) */
    double  dt;                      /* efd.f90:100 (pi/16) */
    double  tfinal;                  /* efd.f90:101 (pi/2) */
    double  xmin, xmax, ymin, ymax;  /* efd.f90:117-118 */
} uapic_efd_config_t;

/* host buffers, synchronous */
int uapic_efd_run(const uapic_efd_config_t *cfg, int64_t nbpart, const double *x, const double *v, double *x_out, double *v_out);
/* device buffers on the current device, asynchronous on `stream` (cudaStream_t, may be null) */
int uapic_efd_run_device(const uapic_efd_config_t *cfg, int64_t nbpart, const double *x, const double *v, double *x_out,
                         double *v_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UAPIC_B200_H */

// uapic_internal.h -- launcher declarations shared by uapic_kernels.cu and uapic_capi.cu (device pointers only)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "uapic_device.cuh"

// records the message uapic_last_error() returns on this thread and returns `code` (uapic_capi.cu)
int uapic_fail(int code, const char *fmt, ...);
// NCCL bound at run time (uapic_capi.cu): communicator on the current device, in-place sum on a stream
int uapic_internal_nccl_comm_init(void **comm, const void *id128, int nranks, int rank);
int uapic_internal_nccl_allreduce(void *comm, void *buf, size_t count, int is_i64, cudaStream_t st);
void uapic_internal_nccl_comm_destroy(void *comm);

namespace uapic {

struct LaunchCtx {
    cudaStream_t stream;
    int sm_count;
    int64_t *launches;   // host counter, may be null
};

inline bool ntau_supported(int ntau) { return ntau == 2 || ntau == 4 || ntau == 8 || ntau == 16 || ntau == 32; }   // the fast lane-per-sample kernels
// any other even ntau up to this goes to the one-warp-per-particle kernels of uapic_generic.cu (direct DFTs): complete, not fast
constexpr int kGenericMaxNtau = 256;
bool generic_ntau_supported(int ntau);
cudaError_t launch_preparation_generic(const LaunchCtx &c, int ntau, double eps, double dt, int64_t np, const double *x, const double *v,
                                       const double *e, double *b, double *t, double *pl, double *ql, double *xt, double *yt);
cudaError_t launch_compute_f_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *b, const double *xt, const double *yt,
                                     const double *et, double *fx, double *fy, int normalise);
cudaError_t launch_fft_tau_generic(const LaunchCtx &c, int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise);
cudaError_t launch_step_fortran_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *pl,
                                        const double *ql, double *xt, double *xf, const double *fx, const double *gx, int corrector);
cudaError_t launch_deposit_tau_generic(const LaunchCtx &c, const MeshDev &m, int ntau, double eps, int64_t np, const double *xt,
                                       const double *t, double w, const RhoAcc &acc, double *x, int wrap);
cudaError_t launch_compute_v_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *yt,
                                     int yt_is_fourier, double *v);

// ---- stage kernels (reference-shaped arrays, natural-order Fourier arrays) ----------------------------------
cudaError_t launch_preparation(const LaunchCtx &c, int ntau, double eps, double dt, int64_t np, const double *x,
                               const double *v, const double *e, double *b, double *t, double *pl, double *ql,
                               double *xt, double *yt);
cudaError_t launch_gather_tau(const LaunchCtx &c, const MeshDev &m, const double *emesh, int ntau, int64_t np,
                              const double *xt, double *et, int wrap);
cudaError_t launch_compute_f(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *b, const double *xt,
                             const double *yt, const double *et, double *fx, double *fy, int normalise);
cudaError_t launch_fft_tau(const LaunchCtx &c, int ntau, int64_t nvec, const double *in, double *out, int sign,
                           int normalise);
cudaError_t launch_step_pointwise(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t,
                                  const double *pl, const double *ql, const double *xf, const double *fx,
                                  const double *gx, double *out);
cudaError_t launch_step_fortran(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t,
                                const double *pl, const double *ql, double *xt, double *xf, const double *fx,
                                const double *gx, int corrector);
cudaError_t launch_deposit_tau(const LaunchCtx &c, const MeshDev &m, int ntau, double eps, int64_t np,
                               const double *xt, const double *t, double w, const RhoAcc &acc, double *x, int wrap);
cudaError_t launch_compute_v(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *yt,
                             int yt_is_fourier, double *v);
cudaError_t launch_deposit(const LaunchCtx &c, const MeshDev &m, int64_t np, double *x, double w, const RhoAcc &acc,
                           int wrap, int scheme = 0);
cudaError_t launch_gather(const LaunchCtx &c, const MeshDev &m, const double *emesh, int64_t np, double *x, double *ep,
                          int wrap, int scheme = 0);

// ---- mesh kernels ---------------------------------------------------------------------------------------------
// raw accumulation (fp64 or fixed point) -> neutralised rho with ghosts; rho_total (1 double) may be null
// raw[i] += sum_{c >= 1} raw[c*n + i] for i < n (fold the CTA-private copies of the one-pass deposits; exact for int64)
cudaError_t launch_fold_raw(const LaunchCtx &c, const RhoAcc &acc, int64_t n, int copies);
cudaError_t launch_rho_epilogue(const LaunchCtx &c, const MeshDev &m, const RhoAcc &acc, double *rho, double *rho_total);
struct PoissonWork { double2 *rk; double2 *ek; };   // rk: (nx/2+1)*ny ; ek: 2*(nx/2+1)*ny
cudaError_t launch_poisson(const LaunchCtx &c, const MeshDev &m, const PoissonWork &w, const double *rho, double *emesh,
                           double *energy);
bool poisson_size_supported(int n);
// the session's field solve in one cooperative launch for nb <= 2 meshes at once (k_field_solve, uapic_kernels.cu):
// fold of `fold_copies` raw copies (copy k of mesh b at acc[b] + k*fold_stride) -> rho epilogue -> Poisson -> energy -> halo copy
constexpr int kMaxPeers = 16;
struct SolveBatch {
    int nb;
    RhoAcc acc[2];         // summed raw deposits (fp64 or fixed point)
    double *rho[2];        // (nx+1)*(ny+1) out
    double2 *emesh[2];     // (nx+1)*(ny+1) out
    double2 *ehalo[2];     // halo copy out (may be null)
    double *energy[2];     // 1 double out (may be null)
    double2 *rk[2], *ek[2];   // Poisson work: (nx/2+1)*ny and 2*(nx/2+1)*ny
    double *partial;       // field_solve_scratch_bytes()
    int halo_tiled;        // 1: 2 x 4-node tiled halo (one-pass kernels); 0: linear halo (two-barrier kernels)
    int fold_copies;       // >= 1
    size_t fold_stride;    // elements between copies
    // peer mode (uapic_session_init_peers): the sum over the ranks happens in phase 0, straight out of the ranks' exchange buffers
    int npeers;                                   // 0 = off
    const unsigned long long *peer_data[kMaxPeers];   // rank r's folded deposits of this exchange: mesh b at + b * peer_mesh_stride (8-byte elements)
    const unsigned long long *peer_flag[kMaxPeers];   // rank r's "published" counter
    unsigned long long peer_seq;                  // wait until every flag >= this
    size_t peer_mesh_stride;
    int *peer_error;                              // set to 1 if a rank did not publish within ~4 s
};
cudaError_t launch_fold_publish(const LaunchCtx &c, const RhoAcc &acc, int64_t n, int copies, void *dst, unsigned long long *flag,
                                unsigned long long seq);
size_t field_solve_scratch_bytes();
cudaError_t launch_field_solve(const LaunchCtx &c, const MeshDev &m, const SolveBatch &B);
// periodic halo copy of the E mesh for the fused gathers: node (i,j), i in [-2,nx+3], j in [-2,ny+3], holds E(i mod nx, j mod ny)
inline size_t ehalo_nodes(const MeshDev &m) { return (size_t)(m.nx + 6) * (size_t)(m.ny + 6); }
cudaError_t launch_extend_emesh(const LaunchCtx &c, const MeshDev &m, const double *emesh, double2 *ehalo);
// the same nodes with every 128-byte line holding a 2 x 4 block of nodes (one-pass kernels; layout in uapic_fast.cuh)
inline size_t ehalo_tiled_nodes(const MeshDev &m) { return (size_t)((m.nx + 6 + 1) >> 1) * (size_t)((m.ny + 6 + 3) >> 2) * 8; }
cudaError_t launch_extend_emesh_tiled(const LaunchCtx &c, const MeshDev &m, const double *emesh, double2 *ehalo);

// ---- fused phase kernels (session path) -----------------------------------------------------------------------
struct PhaseParams {
    MeshDev m;
    double eps, dt, weight;
    int64_t np;
    int wrap;
    int ntau;
    double2 *x;            // (2,np): read by A, written by B
    double2 *v;            // (2,np): read by A, written by B
    const double2 *ep;     // (2,np): particles.e, frozen after init
    const double2 *emesh;  // (2,nx+1,ny+1)
    const double2 *ehalo;  // periodic halo copy of emesh: (nx+6) x (ny+6) nodes, node (i,j) at [(i+2) + (nx+6)*(j+2)]
    double2 *store;        // store-full: np * 8 * ntau complex, what crosses the intra-step barrier
    double *etstore;       // hybrid: np * 2 * ntau doubles (E at the tau samples of the predictor)
    int hybrid;
    double2 *tb;           // (t,b) per particle
    RhoAcc rho;
};
cudaError_t launch_phase_a(const LaunchCtx &c, const PhaseParams &p);
cudaError_t launch_phase_b(const LaunchCtx &c, const PhaseParams &p);

// ---- one-pass kernels (session path, one field barrier per step; uapic_onepass.cu) ------------------------------
struct OnepassParams {
    MeshDev m;
    double eps, dt, weight;
    int64_t np;
    int wrap;
    int ntau;              // 8, 16 or 32
    int full;              // 1: 72 B per particle-tau (W_n and interv stored); 0: 48 B (phase B recomputes them)
    int scheme;            // UAPIC_SCHEME_M6 (0) or UAPIC_SCHEME_CIC (1, build-defined; lean layout only)
    double2 *x;            // (2,np): read and rewritten (corrector position) by A
    double2 *v;            // (2,np): read by A, written by B
    const double2 *ep;     // (2,np): particles.e, frozen after init
    const double2 *ehalo;  // TILED periodic halo copy of the field to gather: E_n for A, E_pred for B
    const double2 *ehalo_b; // fuse_b: E_pred of the PREVIOUS step (phase B part of the fused kernel)
    int fuse_b;            // launch_onepass_a only: 1 = run phase B of the previous step inside it (lean layout)
    char *store;           // np * onepass_store_bytes_per_particle(ntau, full)
    double *rec;           // np * 8 doubles: t, b, 1/b, bracket sums (2), cos(t/eps), sin(t/eps), unused
    RhoAcc rho_p, rho_c;   // raw accumulation meshes of the predictor and the corrector deposit (A only)
    const uint32_t *out_perm;   // null, or slot -> caller's particle index: A writes the new x, B the new v, straight to
    double2 *x_out, *v_out;     //   x_out[out_perm[slot]] / v_out[out_perm[slot]] (host-resident stepping) instead of x / v in place
    int rho_copies;        // >= 1: CTA b deposits into copy b % rho_copies (copy c of both meshes starts 2*c*nrho elements on)
};
bool onepass_ntau_supported(int ntau);
size_t onepass_store_bytes_per_particle(int ntau, int full);
cudaError_t launch_onepass_a(const LaunchCtx &c, const OnepassParams &p);
cudaError_t launch_onepass_b(const LaunchCtx &c, const OnepassParams &p);

// ---- spatial reordering of the particle arrays (uapic_sort.cu) ----------------------------------------------------
int sort_bins(const MeshDev &m, int bin_cells_log2);
cudaError_t launch_sort_particles(const LaunchCtx &c, const MeshDev &m, int bin_cells_log2, int64_t np, const double2 *x,
                                  const double2 *v, const double2 *ep, const uint32_t *perm, double2 *x2, double2 *v2,
                                  double2 *ep2, uint32_t *perm2, uint16_t *binid, unsigned *hist, uint32_t index_base = 0);
cudaError_t launch_unpermute(const LaunchCtx &c, int64_t np, const uint32_t *perm, const double2 *a, double2 *out);

// ---- the external-field program of fortran/efd.f90 as one kernel (uapic_efd.cu) -------------------------------------
// box = xmin, xmax, ymin, ymax; x, v, x_out, v_out: (2, np) device arrays (x_out / v_out may alias x / v)
bool efd_ntau_supported(int ntau);
cudaError_t launch_efd(const LaunchCtx &c, int ntau, double eps, double dt, double tfinal, int nstep, const double *box, int64_t np,
                       const double *x, const double *v, double *x_out, double *v_out);

// ---- loaders / diagnostics --------------------------------------------------------------------------------------
cudaError_t launch_generate(const LaunchCtx &c, const MeshDev &m, int kind, uint64_t seed, int64_t first, int64_t stride, int64_t np,
                            int64_t np_global, double alpha, double kx, double *x, double *v);
size_t sum_v_scratch_bytes();
cudaError_t launch_sum_v(const LaunchCtx &c, int64_t np, const double *v, double *scratch, double *out2);
cudaError_t probe_fp64_peak(const LaunchCtx &c, int launches, double *dfma_per_s, double *ms_per_launch);

}  // namespace uapic

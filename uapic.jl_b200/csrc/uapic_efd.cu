// uapic_efd.cu -- the reference's external-field two-scale program (fortran/efd.f90; test/test_efd.jl is its Julia twin) as
// ONE kernel, sm_100a.
//
// What the program is: every particle is integrated on its own, in tau-Fourier space, in the prescribed field
//     E(x,t) = (cos(x1/2) sin(x2) / 2, sin(x1/2) cos(x2)) (1 + sin(t)/2),      b(x) = 1 + sin(x1) sin(x2) / 2
// (efd.f90:166-167, 139) -- a third-order prepared initial datum (efd.f90:157-383), nstep second-order IMEX steps
// (efd.f90:388-454), and the physical state read off at tau = tfinal b / eps (efd.f90:456-478).  There is no mesh, no deposit
// and no field solve inside the loop, so nothing crosses between particles: the whole program is one launch, the only memory
// traffic is 32 B in and 32 B out per particle, and the kernel is bound by the fp64 pipe (transcendentals of the positions at
// every tau sample, twice per step).  As shipped the Fortran program runs one particle and stops mid-way (efd.f90:131,257);
// this is the text behind that `stop` over all particles -- the computation that produced the constants of efd.f90:481,
// which the oracle reproduces to 13 digits from init_particles_2d's own load (tests/test_efd_oracle.py).
//
// Mapping.  Power-of-two ntau <= 32: one tau sample per lane, ntau lanes per particle, the length-ntau transforms as
// shuffle butterflies (TauLane / fft_fwd / fft_bwd of uapic_device.cuh; Fourier slots live bit-reversed, which every
// spectral operation here -- diagonal multipliers, mode 0, sums over all modes -- is indifferent to).  Any other even
// ntau <= 256: one warp per particle, samples strided over the lanes, direct DFTs out of shared memory (as
// uapic_generic.cu).  One body serves both through a small policy type.
//
// Quantities the program computes and never uses (pl, ql, gx, ave2: efd.f90:143-149,275,283) are left out.
#include "uapic_internal.h"
#include "uapic_efd_body.cuh"

namespace uapic {

namespace {

constexpr int kEfdBlock = 128;

struct EfdArgs {
    EfdScalars s;
    int ntau;
    int64_t np;
    const double2 *x, *v;
    double2 *xo, *vo;
};

// ---- one tau sample per lane ------------------------------------------------------------------------------------
template <int N> struct LaneTau {
    static constexpr int SPL = 1;
    static constexpr int kLanesPerParticle = N;
    static constexpr int kMinBlocks = 4;          // 128 registers: 16 warps per SM hide the shuffle and DFMA latencies of the butterflies
    TauLane<N> L;
    DEVINL void init(int, cd *) { L.init(threadIdx.x & 31); }
    DEVINL bool leader() const { return L.j == 0; }
    DEVINL double ct(int) const { return L.ct; }
    DEVINL double st(int) const { return L.st; }
    DEVINL bool mode_live(int) const { return true; }
    DEVINL double lmode(int) const { return L.lf; }
    DEVINL void fwd(cd (&a)[1]) const { a[0] = rmul(1.0 / (double)N, fft_fwd<N>(a[0], L)); }      // fft.f90:61-72 (carries 1/n)
    DEVINL void inv(cd (&a)[1]) const { a[0] = fft_bwd<N>(a[0], L); }                             // fft.f90:74-81
    DEVINL cd first(const cd (&a)[1]) const { return group_bcast0<N>(a[0]); }                     // tau index 0 == Fourier slot 0
    DEVINL cd sum(cd v) const { return mk(group_sum<N>(v.re), group_sum<N>(v.im)); }
};

// ---- one warp per particle, R samples per lane ------------------------------------------------------------------
template <int R> struct WarpTau {
    static constexpr int SPL = R;
    static constexpr int kLanesPerParticle = 32;
    static constexpr int kMinBlocks = 1;
    int N, lane;
    cd *buf;
    const cd *tw;
    double c_[R], s_[R];
    DEVINL void init(int ntau, cd *smem) {
        N = ntau; lane = threadIdx.x & 31;
        for (int m = threadIdx.x; m < ntau; m += blockDim.x) {
            double s, c;
            sincospi(-2.0 * (double)m / (double)ntau, &s, &c);
            smem[m] = mk(c, s);
        }
        tw = smem;
        buf = smem + ntau + (threadIdx.x >> 5) * ntau;
#pragma unroll
        for (int j = 0; j < R; ++j) sincospi(2.0 * (double)(lane + 32 * j) / (double)ntau, &s_[j], &c_[j]);
        __syncthreads();
    }
    DEVINL bool leader() const { return lane == 0; }
    DEVINL double ct(int j) const { return c_[j]; }
    DEVINL double st(int j) const { return s_[j]; }
    DEVINL bool mode_live(int j) const { return lane + 32 * j < N; }
    DEVINL double lmode(int j) const { const int k = lane + 32 * j; return (double)(k < N / 2 ? k : k - N); }
    DEVINL void dft(cd (&a)[R], bool forward) const {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < R; ++j) { const int n = lane + 32 * j; if (n < N) buf[n] = a[j]; }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int k = lane + 32 * j;
            cd acc = mk(0.0, 0.0);
            if (k < N) {
                int idx = 0;
                for (int n = 0; n < N; ++n) {
                    cd w = tw[idx];
                    if (!forward) w.im = -w.im;
                    acc = cfma(buf[n], w, acc);
                    idx += k; if (idx >= N) idx -= N;
                }
                if (forward) acc = acc / (double)N;
            }
            a[j] = acc;
        }
        __syncwarp();
    }
    DEVINL void fwd(cd (&a)[R]) const { dft(a, true); }
    DEVINL void inv(cd (&a)[R]) const { dft(a, false); }
    DEVINL cd first(const cd (&a)[R]) const { return mk(__shfl_sync(kFull, a[0].re, 0), __shfl_sync(kFull, a[0].im, 0)); }
    DEVINL cd sum(cd v) const {
#pragma unroll
        for (int h = 16; h >= 1; h >>= 1) { v.re += __shfl_xor_sync(kFull, v.re, h); v.im += __shfl_xor_sync(kFull, v.im, h); }
        return v;
    }
};


template <class P> __global__ void __launch_bounds__(kEfdBlock, P::kMinBlocks) k_efd(const EfdArgs q) {
    extern __shared__ double2 efd_smem[];
    P T;
    T.init(q.ntau, reinterpret_cast<cd *>(efd_smem));
    constexpr int per_block = kEfdBlock / P::kLanesPerParticle;
    const int64_t rounds = (q.np + per_block - 1) / per_block;           // every lane of a group takes part in the shuffles
    for (int64_t blk = blockIdx.x; blk < rounds; blk += gridDim.x) {
        const int64_t p = blk * per_block + threadIdx.x / P::kLanesPerParticle;
        const int64_t pc = p < q.np ? p : q.np - 1;
        const double2 xx = q.x[pc], vv = q.v[pc];
        double xo[2], vo[2];
        efd_particle<P>(T, q.s, xx.x, xx.y, vv.x, vv.y, xo, vo);
        if (p < q.np && T.leader()) { q.xo[p] = make_double2(xo[0], xo[1]); q.vo[p] = make_double2(vo[0], vo[1]); }
    }
}

template <class P> cudaError_t launch_one(const LaunchCtx &c, const EfdArgs &q, size_t smem) {
    constexpr int per_block = kEfdBlock / P::kLanesPerParticle;
    const int64_t rounds = (q.np + per_block - 1) / per_block;
    int64_t grid = (int64_t)c.sm_count * 8;
    if (grid > rounds) grid = rounds;
    if (grid < 1) grid = 1;
    k_efd<P><<<(unsigned)grid, kEfdBlock, smem, c.stream>>>(q);
    if (c.launches) ++*c.launches;
    return cudaGetLastError();
}

}  // namespace

bool efd_ntau_supported(int ntau) { return ntau >= 2 && ntau <= kGenericMaxNtau && (ntau & 1) == 0; }

cudaError_t launch_efd(const LaunchCtx &c, int ntau, double eps, double dt, double tfinal, int nstep, const double *box, int64_t np,
                       const double *x, const double *v, double *x_out, double *v_out) {
    if (np <= 0) return cudaSuccess;
    EfdArgs q;
    q.s.eps = eps; q.s.dt = dt; q.s.tfinal = tfinal; q.s.nstep = nstep; q.ntau = ntau; q.np = np;
    q.s.xmin = box[0]; q.s.xmax = box[1]; q.s.ymin = box[2]; q.s.ymax = box[3];
    q.x = reinterpret_cast<const double2 *>(x); q.v = reinterpret_cast<const double2 *>(v);
    q.xo = reinterpret_cast<double2 *>(x_out); q.vo = reinterpret_cast<double2 *>(v_out);
    switch (ntau) {
        case 2: return launch_one<LaneTau<2>>(c, q, 0);
        case 4: return launch_one<LaneTau<4>>(c, q, 0);
        case 8: return launch_one<LaneTau<8>>(c, q, 0);
        case 16: return launch_one<LaneTau<16>>(c, q, 0);
        case 32: return launch_one<LaneTau<32>>(c, q, 0);
        default: break;
    }
    const size_t smem = sizeof(cd) * (size_t)ntau * (1 + kEfdBlock / 32);
    if (ntau <= 32) return launch_one<WarpTau<1>>(c, q, smem);
    if (ntau <= 64) return launch_one<WarpTau<2>>(c, q, smem);
    if (ntau <= 128) return launch_one<WarpTau<4>>(c, q, smem);
    return launch_one<WarpTau<8>>(c, q, smem);
}

}  // namespace uapic

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
// Mapping.  Power-of-two ntau <= 32: one tau sample per lane, ntau lanes per particle (LaneTau) -- or, the default for
// ntau >= 4, two samples per lane on ntau/2 lanes (LaneTau2, first butterfly stage inside the thread) --, the transforms as
// shuffle butterflies (TauLane / fft_fwd / fft_bwd of uapic_device.cuh; Fourier slots live bit-reversed, which every
// spectral operation here -- diagonal multipliers, mode 0, sums over all modes -- is indifferent to).  Any other even
// ntau <= 256: one warp per particle, samples strided over the lanes, direct DFTs out of shared memory (as
// uapic_generic.cu).  One body serves both through a small policy type.
//
// Quantities the program computes and never uses (pl, ql, gx, ave2: efd.f90:143-149,275,283) are left out.
#include <cstdlib>

#include "uapic_internal.h"

namespace uapic {
namespace {
// Out-of-line transcendentals.  Every inlined fp64 sin / cos / sincos is ~60 instructions plus the slow-path call; the body has
// about a hundred of them and, inlined, was larger than the SM's instruction cache several times over.
struct sc2 { double s, c; };
__device__ __noinline__ double efd_sin(double x) { return sin(x); }
__device__ __noinline__ sc2 efd_sincos_ool(double x) { sc2 r; sincos(x, &r.s, &r.c); return r; }
DEVINL void efd_sincos(double x, double *s, double *c) { const sc2 r = efd_sincos_ool(x); *s = r.s; *c = r.c; }
}  // namespace
}  // namespace uapic

#include "uapic_efd_body.cuh"

namespace uapic {

namespace {

constexpr int kEfdBlock = 128;
constexpr bool kEfdSpl2Default = true;    // two tau samples per lane (LaneTau2) for ntau = 4..32: 7 % faster than one (profiles/r2w_efd_spl2.log); UAPIC_EFD_SPL2=0/1 overrides

struct EfdArgs {
    EfdScalars s;
    int ntau;
    int64_t np;
    const double2 *x, *v;
    double2 *xo, *vo;
};

// ---- one tau sample per lane ------------------------------------------------------------------------------------
// The butterflies are NOT inlined.  Inlined, the ~55 transforms of a particle made the kernel ~110 KB of SASS and the IMEX
// loop alone larger than the 32 KB instruction cache of an SM: ncu showed 4.4 "no instruction" stall cycles per issued
// instruction and the fp64 pipe at 37 % (profiles/r2g_efd_ncu_full.csv).  As two out-of-line functions that take both
// components at once (fft_2d / ifft_2d of fft.f90:37-59; the two butterflies interleave), the loop body fits.
struct cd2 { cd a, b; };

#ifndef UAPIC_EFD_TW_REGS
#define UAPIC_EFD_TW_REGS 0     // 0: stage twiddles read from a per-CTA shared-memory table; 1: passed in registers (measured 4.8 % slower: 128 registers + spills, profiles/r2k2_efd_twiddle_ab.log)
#endif
#if UAPIC_EFD_TW_REGS
struct Tw4 { double r[4], i[4]; };     // stage twiddles of this lane (stage s <-> half-size N >> (s+1)); unused stages 1, 0

template <int N> struct LaneFft {
    static constexpr int LOG = Log2<N>::v;
    // twiddles travel in registers (by value): as LDS.128 from a per-CTA table they were a quarter of the kernel's L1 data-pipe
    // wavefronts, the unit this kernel is bound by once the instruction fetch is out of the way (profiles/r2h_efd_ncu_full.csv)
    static __device__ __noinline__ cd2 fwd(cd a, cd b, double w0r, double w0i, double w1r, double w1i, double w2r, double w2i,
                                           double w3r, double w3i) {
        const double wr[4] = {w0r, w1r, w2r, w3r}, wi[4] = {w0i, w1i, w2i, w3i};
        const int j = threadIdx.x & (N - 1);
#pragma unroll
        for (int s = 0; s < LOG; ++s) {
            const int h = N >> (s + 1);
            const cd oa = shfl_xor(a, h), ob = shfl_xor(b, h);
            const double sg = (j & h) ? -1.0 : 1.0;
            cd da = mk(fma(sg, a.re, oa.re), fma(sg, a.im, oa.im));      // upper: v+o ; lower: o-v
            cd db = mk(fma(sg, b.re, ob.re), fma(sg, b.im, ob.im));
            if (h > 1) { const cd w = mk(wr[s], wi[s]); da = cmul(da, w); db = cmul(db, w); }
            a = da; b = db;
        }
        cd2 r;
        r.a = rmul(1.0 / (double)N, a); r.b = rmul(1.0 / (double)N, b);       // fft.f90:44-59 carries 1/n
        return r;
    }
    static __device__ __noinline__ cd2 inv(cd a, cd b, double w0r, double w0i, double w1r, double w1i, double w2r, double w2i,
                                           double w3r, double w3i) {
        const double wr[4] = {w0r, w1r, w2r, w3r}, wi[4] = {w0i, w1i, w2i, w3i};
        const int j = threadIdx.x & (N - 1);
#pragma unroll
        for (int s = LOG - 1; s >= 0; --s) {
            const int h = N >> (s + 1);
            if (h > 1) { const cd w = mk(wr[s], wi[s]); a = cmulc(a, w); b = cmulc(b, w); }
            const cd oa = shfl_xor(a, h), ob = shfl_xor(b, h);
            const double sg = (j & h) ? -1.0 : 1.0;
            a = mk(fma(sg, a.re, oa.re), fma(sg, a.im, oa.im));
            b = mk(fma(sg, b.re, ob.re), fma(sg, b.im, ob.im));
        }
        cd2 r;
        r.a = a; r.b = b;
        return r;
    }
};

template <int N> struct LaneTau {
    static constexpr int SPL = 1;
    static constexpr int kLanesPerParticle = N;
    static constexpr int kMinBlocks = 4;          // 128 registers: 16 warps per SM hide the shuffle and DFMA latencies of the butterflies
    static constexpr int LOG = Log2<N>::v;
    int j;
    double c_, s_, lf_;
    Tw4 w;
    static size_t smem_bytes(int) { return 0; }
    DEVINL void init(int, cd *) {
        TauLane<N> L;
        L.init(threadIdx.x & 31);
        j = L.j; c_ = L.ct; s_ = L.st; lf_ = L.lf;
#pragma unroll
        for (int s = 0; s < 4; ++s) { w.r[s] = (s < LOG - 1) ? L.twr[s] : 1.0; w.i[s] = (s < LOG - 1) ? L.twi[s] : 0.0; }
    }
    DEVINL bool leader() const { return j == 0; }
    DEVINL double ct(int) const { return c_; }
    DEVINL double st(int) const { return s_; }
    DEVINL bool mode_live(int) const { return true; }
    DEVINL double lmode(int) const { return lf_; }                        // Fourier slot of lane j is bitrev(j)
    DEVINL void fwd2(cd (&a)[1], cd (&b)[1]) const {
        const cd2 r = LaneFft<N>::fwd(a[0], b[0], w.r[0], w.i[0], w.r[1], w.i[1], w.r[2], w.i[2], w.r[3], w.i[3]);
        a[0] = r.a; b[0] = r.b;
    }
    DEVINL void inv2(cd (&a)[1], cd (&b)[1]) const {
        const cd2 r = LaneFft<N>::inv(a[0], b[0], w.r[0], w.i[0], w.r[1], w.i[1], w.r[2], w.i[2], w.r[3], w.i[3]);
        a[0] = r.a; b[0] = r.b;
    }
    DEVINL void fwd(cd (&a)[1]) const {
        a[0] = LaneFft<N>::fwd(a[0], mk(0.0, 0.0), w.r[0], w.i[0], w.r[1], w.i[1], w.r[2], w.i[2], w.r[3], w.i[3]).a;
    }
    DEVINL cd first(const cd (&a)[1]) const { return group_bcast0<N>(a[0]); }                     // tau index 0 == Fourier slot 0
    DEVINL cd sum(cd v) const { return mk(group_sum<N>(v.re), group_sum<N>(v.im)); }
};

#else
template <int N> struct LaneFft {
    static constexpr int LOG = Log2<N>::v;
    // stage twiddles of lane j, written once per CTA by LaneTau::init: tw[s * N + j], s < LOG - 1
    static __device__ __noinline__ cd2 fwd(cd a, cd b) {
        extern __shared__ double2 efd_smem[];
        const cd *tw = reinterpret_cast<const cd *>(efd_smem);
        const int j = threadIdx.x & (N - 1);
#pragma unroll
        for (int s = 0; s < LOG; ++s) {
            const int h = N >> (s + 1);
            const cd oa = shfl_xor(a, h), ob = shfl_xor(b, h);
            const double sg = (j & h) ? -1.0 : 1.0;
            cd da = mk(fma(sg, a.re, oa.re), fma(sg, a.im, oa.im));      // upper: v+o ; lower: o-v
            cd db = mk(fma(sg, b.re, ob.re), fma(sg, b.im, ob.im));
            if (h > 1) { const cd w = tw[s * N + j]; da = cmul(da, w); db = cmul(db, w); }
            a = da; b = db;
        }
        cd2 r;
        r.a = rmul(1.0 / (double)N, a); r.b = rmul(1.0 / (double)N, b);       // fft.f90:44-59 carries 1/n
        return r;
    }
    static __device__ __noinline__ cd2 inv(cd a, cd b) {
        extern __shared__ double2 efd_smem[];
        const cd *tw = reinterpret_cast<const cd *>(efd_smem);
        const int j = threadIdx.x & (N - 1);
#pragma unroll
        for (int s = LOG - 1; s >= 0; --s) {
            const int h = N >> (s + 1);
            if (h > 1) { const cd w = tw[s * N + j]; a = cmulc(a, w); b = cmulc(b, w); }
            const cd oa = shfl_xor(a, h), ob = shfl_xor(b, h);
            const double sg = (j & h) ? -1.0 : 1.0;
            a = mk(fma(sg, a.re, oa.re), fma(sg, a.im, oa.im));
            b = mk(fma(sg, b.re, ob.re), fma(sg, b.im, ob.im));
        }
        cd2 r;
        r.a = a; r.b = b;
        return r;
    }
};

template <int N> struct LaneTau {
    static constexpr int SPL = 1;
    static constexpr int kLanesPerParticle = N;
    static constexpr int kMinBlocks = 4;          // 128 registers: 16 warps per SM hide the shuffle and DFMA latencies of the butterflies
    static constexpr int LOG = Log2<N>::v;
    int j;
    double c_, s_, lf_;
    static size_t smem_bytes(int) { return sizeof(cd) * (size_t)N * (LOG > 1 ? LOG - 1 : 1); }
    DEVINL void init(int, cd *smem) {
        TauLane<N> L;
        L.init(threadIdx.x & 31);
        j = L.j; c_ = L.ct; s_ = L.st; lf_ = L.lf;
        if (threadIdx.x < N) {
#pragma unroll
            for (int s = 0; s < LOG - 1; ++s) smem[s * N + j] = mk(L.twr[s], L.twi[s]);
        }
        __syncthreads();
    }
    DEVINL bool leader() const { return j == 0; }
    DEVINL double ct(int) const { return c_; }
    DEVINL double st(int) const { return s_; }
    DEVINL bool mode_live(int) const { return true; }
    DEVINL double lmode(int) const { return lf_; }                        // Fourier slot of lane j is bitrev(j)
    DEVINL void fwd2(cd (&a)[1], cd (&b)[1]) const { const cd2 r = LaneFft<N>::fwd(a[0], b[0]); a[0] = r.a; b[0] = r.b; }
    DEVINL void inv2(cd (&a)[1], cd (&b)[1]) const { const cd2 r = LaneFft<N>::inv(a[0], b[0]); a[0] = r.a; b[0] = r.b; }
    DEVINL void fwd(cd (&a)[1]) const { a[0] = LaneFft<N>::fwd(a[0], mk(0.0, 0.0)).a; }
    DEVINL cd first(const cd (&a)[1]) const { return group_bcast0<N>(a[0]); }                     // tau index 0 == Fourier slot 0
    DEVINL cd sum(cd v) const { return mk(group_sum<N>(v.re), group_sum<N>(v.im)); }
};
#endif

#if !UAPIC_EFD_TW_REGS
// ---- two tau samples per lane (N >= 4): lane j of the N/2 lanes of a particle holds samples j and j + N/2 -------------------
// The first butterfly stage of the length-N transform stays inside the thread (its twiddle exp(-2 pi i j / N) is the lane's own
// (cos tau_j, -sin tau_j)); the two half-length transforms that follow run across the N/2 lanes for both slots at once.  Slot 0
// then holds the even mode 2 bitrev(j), slot 1 the odd mode 2 bitrev(j) + 1.  Per particle and transform pair this is 12 + 4 L1
// data-pipe wavefronts (shuffles + twiddle loads) instead of 16 + 6, the unit the kernel is bound by; selected at run time with
// UAPIC_EFD_SPL2 (uapic_efd.cu, launch_efd).
template <int N> struct LaneTau2 {
    static constexpr int SPL = 2;
    static constexpr int M = N / 2;
    static constexpr int kLanesPerParticle = M;
    static constexpr int kMinBlocks = 3;
    static constexpr int LOG = Log2<M>::v;
    int j;
    double c_, s_, l0_, l1_;
    static size_t smem_bytes(int) { return sizeof(cd) * (size_t)M * (LOG > 1 ? LOG - 1 : 1); }
    DEVINL void init(int, cd *smem) {
        TauLane<M> L;                              // lane constants of the half-length transform
        L.init(threadIdx.x & 31);
        j = L.j;
        sincospi(2.0 * (double)j / (double)N, &s_, &c_);
        const int k0 = 2 * L.k, k1 = 2 * L.k + 1;
        l0_ = (double)(k0 < N / 2 ? k0 : k0 - N);
        l1_ = (double)(k1 < N / 2 ? k1 : k1 - N);
        if (threadIdx.x < M) {
#pragma unroll
            for (int s = 0; s < LOG - 1; ++s) smem[s * M + j] = mk(L.twr[s], L.twi[s]);
        }
        __syncthreads();
    }
    DEVINL bool leader() const { return j == 0; }
    DEVINL double ct(int q) const { return q ? -c_ : c_; }                // tau_{j + N/2} = tau_j + pi
    DEVINL double st(int q) const { return q ? -s_ : s_; }
    DEVINL bool mode_live(int) const { return true; }
    DEVINL double lmode(int q) const { return q ? l1_ : l0_; }
    // forward, carrying 1/N: in-thread stage (with 1/2), then the half-length transform (which carries 1/M)
    DEVINL void split(cd (&a)[2]) const {
        const cd u = mk(0.5 * (a[0].re + a[1].re), 0.5 * (a[0].im + a[1].im));
        const cd d = mk(0.5 * (a[0].re - a[1].re), 0.5 * (a[0].im - a[1].im));
        a[0] = u; a[1] = cmul(d, mk(c_, -s_));
    }
    DEVINL void join(cd (&a)[2]) const {
        const cd t = cmul(a[1], mk(c_, s_));
        const cd u = a[0];
        a[0] = mk(u.re + t.re, u.im + t.im); a[1] = mk(u.re - t.re, u.im - t.im);
    }
    DEVINL void fwd(cd (&a)[2]) const { split(a); const cd2 r = LaneFft<M>::fwd(a[0], a[1]); a[0] = r.a; a[1] = r.b; }
    DEVINL void fwd2(cd (&a)[2], cd (&b)[2]) const { fwd(a); fwd(b); }
    DEVINL void inv1(cd (&a)[2]) const { const cd2 r = LaneFft<M>::inv(a[0], a[1]); a[0] = r.a; a[1] = r.b; join(a); }
    DEVINL void inv2(cd (&a)[2], cd (&b)[2]) const { inv1(a); inv1(b); }
    DEVINL cd first(const cd (&a)[2]) const { return group_bcast0<M>(a[0]); }    // tau index 0 and mode 0 are both slot 0 of lane 0
    DEVINL cd sum(cd v) const { return mk(group_sum<M>(v.re), group_sum<M>(v.im)); }
};
#endif

// ---- one warp per particle, R samples per lane ------------------------------------------------------------------
// the DFT is out of line for the same reason as LaneFft (inlined at ~110 call sites with R unrolled, WarpTau<8> was 2 MB of SASS)
template <int R> __device__ __noinline__ void warp_dft(cd (&a)[R], cd *buf, const cd *tw, int N, int lane, bool forward) {
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
    DEVINL void dft(cd (&a)[R], bool forward) const { warp_dft<R>(a, buf, tw, N, lane, forward); }
    static size_t smem_bytes(int ntau) { return sizeof(cd) * (size_t)ntau * (1 + kEfdBlock / 32); }
    DEVINL void fwd(cd (&a)[R]) const { dft(a, true); }
    DEVINL void inv(cd (&a)[R]) const { dft(a, false); }
    DEVINL void fwd2(cd (&a)[R], cd (&b)[R]) const { dft(a, true); dft(b, true); }
    DEVINL void inv2(cd (&a)[R], cd (&b)[R]) const { dft(a, false); dft(b, false); }
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

template <class P> cudaError_t launch_one(const LaunchCtx &c, const EfdArgs &q) {
    const size_t smem = P::smem_bytes(q.ntau);
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
#if !UAPIC_EFD_TW_REGS
    static const bool spl2 = [] { const char *e = getenv("UAPIC_EFD_SPL2"); return e ? atoi(e) != 0 : kEfdSpl2Default; }();
    if (spl2) {
        switch (ntau) {
            case 4: return launch_one<LaneTau2<4>>(c, q);
            case 8: return launch_one<LaneTau2<8>>(c, q);
            case 16: return launch_one<LaneTau2<16>>(c, q);
            case 32: return launch_one<LaneTau2<32>>(c, q);
            default: break;
        }
    }
#endif
    switch (ntau) {
        case 2: return launch_one<LaneTau<2>>(c, q);
        case 4: return launch_one<LaneTau<4>>(c, q);
        case 8: return launch_one<LaneTau<8>>(c, q);
        case 16: return launch_one<LaneTau<16>>(c, q);
        case 32: return launch_one<LaneTau<32>>(c, q);
        default: break;
    }
    if (ntau <= 32) return launch_one<WarpTau<1>>(c, q);
    if (ntau <= 64) return launch_one<WarpTau<2>>(c, q);
    if (ntau <= 128) return launch_one<WarpTau<4>>(c, q);
    return launch_one<WarpTau<8>>(c, q);
}

}  // namespace uapic

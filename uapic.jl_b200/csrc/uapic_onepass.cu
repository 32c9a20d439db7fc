// uapic_onepass.cu -- "one-pass" kernels of the session path, sm_100a: ONE field barrier per UA step.
//
// Two exact identities of the reference's formulas (tests/onepass_algebra_np.py states and checks them in numpy):
//   1. gx of the second compute_f is R(-tau) yt_pred / b (ua_steps.F90:174-175): independent of the predictor field.
//      The corrected x coefficients elt*xf + pl*fx + ql*(gx-fx)/t (ua_steps.F90:260), the corrector deposit position
//      (compute_rho_m6.F90:74-87) and hence rho_{n+1} are known before the predictor Poisson solve.
//   2. only the tau* evaluation sum_k Yc_k exp(+i l_k t/eps) of the corrected y coefficients is used
//      (ua_steps.F90:293-303) and it is linear in gy:
//        sum_k (Yp_k + qt_k (Gy_k - Fy_k)) conj(elt_k) = [sum_k (Yp_k - qt_k Fy_k) conj(elt_k)] + sum_n gy(tau_n) W_n,
//        qt = ql/t,  W_n = (1/N) sum_k qt_k conj(elt_k) exp(-i k tau_n).
//
//   k_onepass_a = preparation + gather(E_n) + compute_f + ua_step1 + BOTH deposit positions + both deposits
//                 (ua_steps.F90:15-236, :252-265 for x, interpolation_m6.F90:40-191, compute_rho_m6.F90:47-189)
//   k_onepass_b = gather(E_pred) + time-domain gy + compute_v           (ua_steps.F90:160-185, :274-307)
//
// What crosses the barrier, per particle-tau: Re xt_pred (2 doubles), yt_pred (2 complex) and, in the FULL layout, W_n
// (1 complex) and interv (1 double): 72 B (FULL) or 48 B (LEAN: phase B recomputes W and interv) instead of 128 B.
// Per particle: an 8-double record (t, b, 1/b, the two bracket sums, cos(t/eps), sin(t/eps)).
//
// Work mapping of phase A: a particle is spread over G = N/8 adjacent lanes, EIGHT tau samples per lane.
//   time domain   : lane g, register s  <->  sample n = g + G*s
//   Fourier domain: lane g, register k1 <->  mode   k = 8*kappa(g) + k1,   kappa = bit reversal of g in log2(G) bits
// A length-N transform is an in-register 8-point FFT, a twiddle, and a G-point FFT across lanes whose twiddles are
// +-1, +-i (no multiplications): 14.5 fp64 instructions and 2 shuffles per sample instead of 26 and 20 for the
// one-sample-per-lane butterfly network of uapic_fused.cu; everything per particle (b, 1/b, sincos(t/eps)) is amortised
// over 8 samples; exp(-i l t/eps) comes from a recurrence over the lane's 8 consecutive modes.
// The M6 gathers run in the "lane = tau sample" layout (positions and fields are exchanged through shared memory), where
// one load instruction touches the orbit of a single particle; E is read from a halo copy whose 128-byte lines hold 2 x 4
// nodes (uapic_fast.cuh).  Phase B lives entirely in that layout.
//
// Execution shape (every choice below was measured, profiles/README.md): phase A runs ONE CTA of 12 warps per SM at 168
// registers, split into four lock-step groups of 3 warps (named barriers between the stages of a tile): a group shares
// the instruction fetches of this 6100-instruction straight-line kernel, the groups sit in different stages so that the
// gather (LSU) and the FFT (fp64) phases overlap on the SM; the barriers also fence the scheduler, which keeps the live
// ranges short (without them ptxas spills 1.3 KB per thread).  Phase B runs 2 CTAs x 8 free-running warps, a contiguous
// range of tiles per CTA.  Both kernels are bound by the L1TEX LSU data pipe, not by HBM.
// The UAPIC_OP_* macros keep the measured alternatives compilable; the defaults are the winners.
#include "uapic_internal.h"
#include "uapic_fast.cuh"

namespace uapic {

namespace {

#ifndef UAPIC_OP_MINB_A
#define UAPIC_OP_MINB_A 1     // one CTA of 12 warps at 168 registers per SM, in four lock-step groups of 3 warps: measured best
#endif
#ifndef UAPIC_OP_MINB_B
#define UAPIC_OP_MINB_B 2
#endif
#ifndef UAPIC_OP_LOCKSTEP
#define UAPIC_OP_LOCKSTEP 1
#endif
#ifndef UAPIC_OP_LOCKGROUP
#define UAPIC_OP_LOCKGROUP 3      // warps per lock-step group inside a CTA of phase A; 0 = the whole CTA
#endif
#if UAPIC_OP_LOCKSTEP && UAPIC_OP_LOCKGROUP
#define OP_STEP() asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 5) / UAPIC_OP_LOCKGROUP), "r"(32 * UAPIC_OP_LOCKGROUP) : "memory")
#elif UAPIC_OP_LOCKSTEP
#define OP_STEP() __syncthreads()
#else
#define OP_STEP() __syncwarp()
#endif
#ifndef UAPIC_OP_PREFETCH
#define UAPIC_OP_PREFETCH 1
#endif
// sync points of phase A.  Besides keeping a lock-step group together they fence ptxas' scheduler: without the three that
// no shared-memory hand-over needs, live ranges grow and the kernel spills 1.3 KB per thread (measured, +17 % time).
#ifndef UAPIC_OP_GATHER_UNROLL
#define UAPIC_OP_GATHER_UNROLL 1
#endif
#ifndef UAPIC_OP_BLOCK_A
#define UAPIC_OP_BLOCK_A 384
#endif
#ifndef UAPIC_OP_BLOCK_B
#define UAPIC_OP_BLOCK_B 256
#endif
#ifndef UAPIC_OP_PAIR_LOADS
#define UAPIC_OP_PAIR_LOADS 0     // 1: M6 gathers fetch two taps per 256-bit load (gather_tiled_pairs). MEASURED SLOWER (profiles/README.md r2a):
                                  // an LDG.256 costs 11.7 data-pipe wavefronts against 5.3 for an LDG.128 -- the pipe is bound by the 128 B/clk
                                  // register write-back, not by the number of lines touched
#endif
#if UAPIC_OP_NO_GATHER
DEVINL void gather_none(const MeshDev &, const double2 *, const Cell &c, double &e1, double &e2) { e1 = c.dpx * 1e-3; e2 = c.dpy * 1e-3; }
#define OP_GATHER_M6 gather_none
#elif UAPIC_OP_PAIR_LOADS
#define OP_GATHER_M6 gather_tiled_pairs
#else
#define OP_GATHER_M6 gather_tiled
#endif
#ifndef UAPIC_OP_STREAM_HINTS
#define UAPIC_OP_STREAM_HINTS 2   // 1: the once-only streams (store, records, particle arrays) use ld/st.global.cs so that they do not
#endif                            //    displace the E-halo lines the gathers live on in L1; 2: ld.global.L1::no_allocate for them
                                  //    (measured, profiles/README.md r2g/r2h: 1 = +-0, 2 = phase B -2.8 %)
#ifndef UAPIC_OP_NO_GATHER        // measurement only: skip the 36 tap loads (E := position-dependent dummy) to count their wavefronts
#define UAPIC_OP_NO_GATHER 0
#endif
#if UAPIC_OP_STREAM_HINTS == 2       // loads of the once-only streams do not allocate in L1 at all
DEVINL double2 ld_noalloc(const double2 *p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
#define OP_LDS(p) ld_noalloc(p)
#define OP_STS(p, v) __stcs(p, v)
#elif UAPIC_OP_STREAM_HINTS
#define OP_LDS(p) __ldcs(p)
#define OP_STS(p, v) __stcs(p, v)
#else
#define OP_LDS(p) (*(p))
#define OP_STS(p, v) (*(p) = (v))
#endif
constexpr int kOpBlockA = UAPIC_OP_BLOCK_A, kOpBlockB = UAPIC_OP_BLOCK_B;      // threads per CTA of the two kernels
constexpr int kGatherUnroll = UAPIC_OP_GATHER_UNROLL;
#ifndef UAPIC_OP_TWIDDLE_TABLE
#define UAPIC_OP_TWIDDLE_TABLE 1     // 0: powers of u1 by recurrence (measured slower: +4 live registers tip phase A into 400 B of spills)
#endif
#ifndef UAPIC_OP_BLOCKED_A
#define UAPIC_OP_BLOCKED_A 0
#endif
#ifndef UAPIC_OP_BLOCKED_B
#define UAPIC_OP_BLOCKED_B 1
#endif
#ifndef UAPIC_OP_PREFETCH_B
#define UAPIC_OP_PREFETCH_B 2
#endif
// OP_SYNC(i): sync point i of a tile (0..6).  Bit i of UAPIC_OP_BARRIER_MASK set = lock-step barrier of the group; clear = just
// __syncwarp(), which still orders the warp's shared-memory accesses and still fences ptxas' scheduler
#ifndef UAPIC_OP_BARRIER_MASK
#define UAPIC_OP_BARRIER_MASK 0x47
#endif
#define OP_SYNC(i) do { if ((UAPIC_OP_BARRIER_MASK >> (i)) & 1) { OP_STEP(); } else { __syncwarp(); } } while (0)
// The sync points of phase A (OP_STEP) do two jobs: they keep a lock-step group together, and they fence ptxas' scheduler --
// without the three that no shared-memory hand-over needs, live ranges grow and the kernel spills 1.3 KB per thread
// (measured: +17 % time).
#ifndef UAPIC_OP_GATHER_UNROLL_B
#define UAPIC_OP_GATHER_UNROLL_B 2
#endif
constexpr int kGatherUnrollB = UAPIC_OP_GATHER_UNROLL_B;
constexpr int kRow = 36;                       // padded row (double2 units) of the per-warp exchange area: conflict-free
constexpr int kTab = 108;                      // cos/sin table (32) + two [4][9]-padded tables: rows in distinct bank groups
constexpr int kWarpSmA = 8 * kRow + 4 * kRow + 16 * 32;  // exchange rows + sin-product rows + yhat stash (double2 units)
constexpr int kWarpSmB = 8 * kRow + 256;       // W_n exchange rows + partial sums of the tau* evaluation

constexpr double kRsqrt2 = 0.70710678118654752440;
constexpr int kSchemeM6 = 0, kSchemeCic = 1;      // UAPIC_SCHEME_*

// ---- 8-point FFT in registers, natural order in and out; SGN = -1 forward (exp(-i..)), +1 backward ---------------
template <int SGN> DEVINL cd mul_i(cd a) { return SGN > 0 ? mk(-a.im, a.re) : mk(a.im, -a.re); }            // a * (SGN i)
template <int SGN> DEVINL cd mul_w8_1(cd a) {                                                              // a * exp(SGN i pi/4)
    return SGN > 0 ? mk((a.re - a.im) * kRsqrt2, (a.im + a.re) * kRsqrt2) : mk((a.re + a.im) * kRsqrt2, (a.im - a.re) * kRsqrt2);
}
template <int SGN> DEVINL cd mul_w8_3(cd a) {                                                              // a * exp(SGN 3 i pi/4)
    return SGN > 0 ? mk((-a.re - a.im) * kRsqrt2, (a.re - a.im) * kRsqrt2) : mk((a.im - a.re) * kRsqrt2, (-a.im - a.re) * kRsqrt2);
}
template <int SGN> DEVINL void fft8(cd (&a)[8]) {
    const cd u0 = cadd(a[0], a[4]), u1 = cadd(a[1], a[5]), u2 = cadd(a[2], a[6]), u3 = cadd(a[3], a[7]);
    const cd v0 = csub(a[0], a[4]), v1 = mul_w8_1<SGN>(csub(a[1], a[5])), v2 = mul_i<SGN>(csub(a[2], a[6])),
             v3 = mul_w8_3<SGN>(csub(a[3], a[7]));
    const cd p0 = cadd(u0, u2), p1 = cadd(u1, u3), q0 = csub(u0, u2), q1 = mul_i<SGN>(csub(u1, u3));
    a[0] = cadd(p0, p1); a[4] = csub(p0, p1); a[2] = cadd(q0, q1); a[6] = csub(q0, q1);
    const cd r0 = cadd(v0, v2), r1 = cadd(v1, v3), s0 = csub(v0, v2), s1 = mul_i<SGN>(csub(v1, v3));
    a[1] = cadd(r0, r1); a[5] = csub(r0, r1); a[3] = cadd(s0, s1); a[7] = csub(s0, s1);
}

// ---- TMEM as lane-private table memory (phase A, UAPIC_OP_TMEM_TABLES) ---------------------------------------------
// Phase A re-reads three small per-lane tables in every tile (twiddles, 1/l and 1/l^2, cos/sin tau: 23 LDS.128 = 92 wavefronts
// of the LSU data pipe per lane and tile, broadcast reads cost as much as any other).  Tensor memory has its own path into the
// register file (LDTM, 12-cycle latency) and every thread can address its own lane of it: the tables live there instead --
// tcgen05.st once per kernel, tcgen05.ld in the tile loop.  Lanes 32*(warp % 4).., 96 columns per warp at (warp / 4) * 96.
// MEASURED (profiles/README.md, r2r/r2s): LSU wavefronts of phase A 464 -> 435 per particle, the pipe 80 % -> 74 % busy, time
// +1 % with one load + wait per entry (1), +-0 with 16-column batches (2): phase A is latency bound, not pipe bound.  Off.
#ifndef UAPIC_OP_TMEM_TABLES
#define UAPIC_OP_TMEM_TABLES 0
#endif
constexpr int kTmemColsPerWarp = 96;
DEVINL unsigned smem_addr_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
DEVINL void tmem_st2(unsigned taddr, double2 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__double2loint(v.x)), "r"(__double2hiint(v.x)),
                 "r"(__double2loint(v.y)), "r"(__double2hiint(v.y)));
}
DEVINL double2 tmem_ld2(unsigned taddr) {
    unsigned r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3));
    return make_double2(__hiloint2double(r1, r0), __hiloint2double(r3, r2));
}

// 8 consecutive double2 (32 columns) with ONE wait: two 16-column loads in flight together
DEVINL void tmem_ld2x8(unsigned taddr, double2 (&out)[8]) {
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]),
                   "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
                   "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),
                   "+r"(r[31]));
#pragma unroll
    for (int q = 0; q < 8; ++q) out[q] = make_double2(__hiloint2double(r[4 * q + 1], r[4 * q]), __hiloint2double(r[4 * q + 3], r[4 * q + 2]));
}

// ---- per-lane constants -----------------------------------------------------------------------------------------
template <int G, bool TM = false> struct OpLane {
    static constexpr int N = 8 * G;
    int lane, g, kap;
    double sg1, sg2;            // -1 on the upper lane of an xor-1 / xor-2 pair
    bool rot;                   // G == 4: lane 3 carries the +-i twiddle of the 4-point cross-lane FFT
    int src_m1, src_p1;         // lane (inside the group) holding Fourier block kappa-1 / kappa+1
    int src_neg, src_neg0;      // ... block G-1-kappa / (G-kappa) mod G  (conjugate partners)
    double l0;                  // l of the lane's first mode
    cd u1;                      // exp(-2 pi i g / N): the lane's twiddles are its powers
    const double2 *cs;          // [N]    (cos tau_n, sin tau_n)                         ua_type.F90:60-62
    const double2 *tw;          // [G][9] exp(-2 pi i g k1 / N), row g (stride 9: the G lanes of a group hit distinct banks)
    const double2 *il;          // [G][9] (1/l_k, 1/l_k^2) of mode k = 8*kappa + k1, row kappa; zero for k = 0   ua_type.F90:51-56

    __host__ __device__ static constexpr int kappa_of(int g) { return G == 4 ? (((g & 1) << 1) | (g >> 1)) : g; }
    DEVINL static double lmode(int k) { return (double)(k < N / 2 ? k : k - N); }

    // tab: kTab double2 of shared memory; every thread of the CTA must call; ends with __syncthreads
    DEVINL void init(int lane_, double2 *tab) {
        lane = lane_;
        g = lane & (G - 1);
        kap = kappa_of(g);
        sg1 = (g & 1) ? -1.0 : 1.0;
        sg2 = (g & 2) ? -1.0 : 1.0;
        rot = (G == 4) && g == 3;
        src_m1 = kappa_of((kap + G - 1) & (G - 1));
        src_p1 = kappa_of((kap + 1) & (G - 1));
        src_neg = kappa_of(G - 1 - kap);
        src_neg0 = kappa_of((G - kap) & (G - 1));
        l0 = lmode(8 * kap);
        { double sn, cs; sincospi(-2.0 * (double)g / (double)N, &sn, &cs); u1 = mk(cs, sn); }
        cs = tab; tw = tab + 32 + 9 * g; il = tab + 68 + 9 * kap;
        const int i = threadIdx.x;
        if (i < N) {
            double s, c;
            sincospi(2.0 * (double)i / (double)N, &s, &c);
            tab[i] = make_double2(c, s);
            const int gg = i >> 3, k1 = i & 7;
            sincospi(-2.0 * (double)(gg * k1) / (double)N, &s, &c);
            tab[32 + 9 * gg + k1] = make_double2(c, s);
            const double l = lmode(i);
            tab[68 + 9 * gg + k1] = (i == 0) ? make_double2(0.0, 0.0) : make_double2(1.0 / l, 1.0 / (l * l));
        }
        __syncthreads();
    }
    DEVINL double lf(int k1) const { return G == 1 ? lmode(k1) : l0 + (double)k1; }

    unsigned taddr = 0;         // TM: this warp's 96 TMEM columns: [0,32) tw, [32,64) il, [64,96) cs of the lane's samples
    DEVINL double2 twv(int k1) const { return TM ? tmem_ld2(taddr + 4 * k1) : tw[k1]; }
    DEVINL void tw8(double2 (&w)[8]) const {
        if (TM && UAPIC_OP_TMEM_TABLES == 2) { tmem_ld2x8(taddr, w); return; }
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) w[k1] = twv(k1);
    }
    DEVINL double2 ilv(int k1) const { return TM ? tmem_ld2(taddr + 32 + 4 * k1) : il[k1]; }
    DEVINL double2 cs_s(int s) const { return TM ? tmem_ld2(taddr + 64 + 4 * s) : cs[g + G * s]; }     // (cos, sin) tau of sample g + G s
    // TM only; every thread of the CTA must call, after init(): allocate (warp 0), fill this lane's tables
    DEVINL void tmem_setup(unsigned *tbase_smem) {
        if ((threadIdx.x >> 5) == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr_u32(tbase_smem)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int w = threadIdx.x >> 5;
        taddr = *tbase_smem + ((unsigned)(32 * (w & 3)) << 16) + (unsigned)((w >> 2) * kTmemColsPerWarp);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            tmem_st2(taddr + 4 * q, tw[q]);
            tmem_st2(taddr + 32 + 4 * q, il[q]);
            tmem_st2(taddr + 64 + 4 * q, cs[g + G * q]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    DEVINL void tmem_release(const unsigned *tbase_smem) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tbase_smem));
    }
};

template <int G> DEVINL double shfl_grp(double v, int src) { return G == 1 ? v : __shfl_sync(kFull, v, src, G); }
template <int G> DEVINL cd shfl_grp(cd v, int src) { return mk(shfl_grp<G>(v.re, src), shfl_grp<G>(v.im, src)); }
template <int G> DEVINL double grp_sum(double v) {
#pragma unroll
    for (int h = G / 2; h >= 1; h >>= 1) v += __shfl_xor_sync(kFull, v, h);
    return v;
}

// butterfly across lanes: a <- partner + sg * own  (lower lane: own + partner ; upper lane: partner - own)
DEVINL void xbfly(cd (&a)[8], int mask, double sg) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double pr = __shfl_xor_sync(kFull, a[i].re, mask), pi = __shfl_xor_sync(kFull, a[i].im, mask);
        a[i] = mk(fma(sg, a[i].re, pr), fma(sg, a[i].im, pi));
    }
}

// forward length-N transform, unnormalised: time layout in, Fourier layout out
template <int G, bool TM> DEVINL void fwdN(cd (&a)[8], const OpLane<G, TM> &L) {
    fft8<-1>(a);
    if (G > 1) {
#if UAPIC_OP_TWIDDLE_TABLE
        if (TM && UAPIC_OP_TMEM_TABLES == 2) {
            double2 w[8]; L.tw8(w);
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) a[k1] = cmul(a[k1], mk(w[k1].x, w[k1].y));
        } else {
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) { const double2 w = L.twv(k1); a[k1] = cmul(a[k1], mk(w.x, w.y)); }
        }
#else
        // powers of u1 by recurrence: 24 fp64 instructions instead of 7 LDS.128 (the LSU data pipe is the scarce unit)
        cd u = L.u1;
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) { a[k1] = cmul(a[k1], u); if (k1 < 7) u = cmul(u, L.u1); }
#endif
    }
    if (G == 4) {
        xbfly(a, 2, L.sg2);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = L.rot ? mk(a[i].im, -a[i].re) : a[i];     // lane 3: times -i
    }
    if (G >= 2) xbfly(a, 1, L.sg1);
}

// backward length-N transform, unnormalised: Fourier layout in, time layout out
template <int G, bool TM> DEVINL void bwdN(cd (&a)[8], const OpLane<G, TM> &L) {
    if (G >= 2) xbfly(a, 1, L.sg1);
    if (G == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = L.rot ? mk(-a[i].im, a[i].re) : a[i];     // lane 3: times +i
        xbfly(a, 2, L.sg2);
    }
    if (G > 1) {
#if UAPIC_OP_TWIDDLE_TABLE
        if (TM && UAPIC_OP_TMEM_TABLES == 2) {
            double2 w[8]; L.tw8(w);
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) a[k1] = cmulc(a[k1], mk(w[k1].x, w[k1].y));
        } else {
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) { const double2 w = L.twv(k1); a[k1] = cmulc(a[k1], mk(w.x, w.y)); }
        }
#else
        cd u = L.u1;
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) { a[k1] = cmulc(a[k1], u); if (k1 < 7) u = cmul(u, L.u1); }
#endif
    }
    fft8<+1>(a);
}

// exp(-i l_k t/eps) for the lane's 8 modes: one sincos for the first mode, a recurrence with e1 = exp(-i t/eps) for the
// others (the reference evaluates every mode directly, ua_steps.F90:64,224,258; the recurrence error, a few ulp, is
// below the rounding of the phase l*t/eps itself)
template <int G, bool TM> DEVINL void elt_modes(const OpLane<G, TM> &L, double t, double eps, cd e1, cd (&elt)[8]) {
    if (G == 1) {
        elt[0] = mk(1.0, 0.0);
        elt[1] = e1;
        elt[2] = cmul(elt[1], e1);
        elt[3] = cmul(elt[2], e1);
        const cd e4 = cmul(elt[3], e1);
        elt[4] = mk(e4.re, -e4.im);                    // l = -4
        elt[5] = mk(elt[3].re, -elt[3].im);
        elt[6] = mk(elt[2].re, -elt[2].im);
        elt[7] = mk(elt[1].re, -elt[1].im);
    } else {
        double s, c;
        sincos(-(L.l0 * t) / eps, &s, &c);
        elt[0] = mk(c, s);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) elt[k1] = cmul(elt[k1 - 1], e1);
    }
}

// pl, ql/t of ua_steps.F90:60-66 for mode (lane, k1); il = (1/l, 1/l^2)
template <int G, bool TM>
DEVINL void pl_qt(const OpLane<G, TM> &L, int k1, double t, double rt, double eps, cd elt, cd &pl, cd &qt) {
    const double2 il = L.ilv(k1);
    const double l = L.lf(k1);
    pl = mk(-eps * elt.im * il.x, eps * (elt.re - 1.0) * il.x);
    qt = mk(eps * eps * (1.0 - elt.re) * il.y * rt, -eps * fma(eps, elt.im, l * t) * il.y * rt);
    if (k1 == 0 && L.kap == 0) { pl = mk(t, 0.0); qt = mk(0.5 * t, 0.0); }
}

// F[R(-tau) y / b]: F[cos(tau) y]_k = (Y_{k-1}+Y_{k+1})/2, F[sin(tau) y]_k = -i (Y_{k-1}-Y_{k+1})/2   (ua_steps.F90:174-175)
DEVINL void fx_modes(double hb, cd p1, cd m1, cd p2, cd m2, cd &fx1, cd &fx2) {
    const cd a1 = cadd(p1, m1), d1 = csub(p1, m1), a2 = cadd(p2, m2), d2 = csub(p2, m2);
    fx1 = mk(hb * (a1.re + d2.im), hb * (a1.im - d2.re));
    fx2 = mk(hb * (a2.re - d1.im), hb * (a2.im + d1.re));
}

// time-domain fy of ua_steps.F90:177-183
DEVINL void fy_time(double ct, double st, double rb, double interv, cd yt1, cd yt2, double et1, double et2, cd &fy1, cd &fy2) {
    const cd t1 = mk(fma(ct * yt2.re - st * yt1.re, interv, et1), (ct * yt2.im - st * yt1.im) * interv);
    const cd t2 = mk(fma(-(ct * yt1.re + st * yt2.re), interv, et2), -(ct * yt1.im + st * yt2.im) * interv);
    fy1 = mk((ct * t1.re - st * t2.re) * rb, (ct * t1.im - st * t2.im) * rb);
    fy2 = mk((st * t1.re + ct * t2.re) * rb, (st * t1.im + ct * t2.im) * rb);
}

DEVINL double re_mulc(cd a, cd b) { return fma(a.re, b.re, a.im * b.im); }    // Re(a * conj(b))
DEVINL double re_mul(cd a, cd b) { return fma(a.re, b.re, -a.im * b.im); }    // Re(a * b)

// three M6 weights of one half of the stencil: u = 1-dp for offsets -2,-1,0 ; u = dp for offsets 3,2,1 (in this order)
DEVINL void m6_half_weights(double u, double (&w)[3]) {
    const double a = pow5(u), b = pow5(1.0 + u), c = pow5(2.0 + u);
    const double k = 1.0 / 120.0;
    w[0] = a * k;
    w[1] = fma(-6.0, a, b) * k;
    w[2] = fma(15.0, a, fma(-6.0, b, c)) * k;
}

// deposit of one particle by its G lanes (compute_rho_m6.F90:89-187): the 36 live taps are split 3x3 (G = 4), 3x6 (G = 2)
template <int G>
DEVINL void deposit_split(const MeshDev &m, const RhoAcc &r, const Cell &c, double weight, int g, bool valid) {
    constexpr int NXP = (G >= 2) ? 3 : 6, NYP = (G == 4) ? 3 : 6;
    double wx[6], wy[6];
    int ox[6], oy[6];
    if (G >= 2) {
        const bool hi = g & 1;
        double w3[3];
        m6_half_weights(hi ? c.dpx : 1.0 - c.dpx, w3);
#pragma unroll
        for (int i = 0; i < 3; ++i) { wx[i] = w3[i]; ox[i] = hi ? 3 - i : i - 2; }
    } else {
        m6_weights_fast(c.dpx, wx);
#pragma unroll
        for (int i = 0; i < 6; ++i) ox[i] = i - 2;
    }
    if (G == 4) {
        const bool hi = g & 2;
        double w3[3];
        m6_half_weights(hi ? c.dpy : 1.0 - c.dpy, w3);
#pragma unroll
        for (int i = 0; i < 3; ++i) { wy[i] = w3[i]; oy[i] = hi ? 3 - i : i - 2; }
    } else {
        m6_weights_fast(c.dpy, wy);
#pragma unroll
        for (int i = 0; i < 6; ++i) oy[i] = i - 2;
    }
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < NYP; ++j) {
        const int jy = wrap_fast(c.j, oy[j], m.ny) * m.ld;
        const double wyw = wy[j] * weight;
#pragma unroll
        for (int i = 0; i < NXP; ++i) rho_add(r, wrap_fast(c.i, ox[i], m.nx) + jy, wx[i] * wyw);
    }
}

struct OpDev {
    OnepassParams p;
    MeshFast f;
    double inv_eps;
};

// byte offsets inside a particle's block of the store (N = ntau): xtr | yt1 | yt2 | W | interv
template <int N, bool FULL> struct StoreMap {
    static constexpr size_t stride = (FULL ? 72 : 48) * (size_t)N;
    DEVINL static double2 *xtr(char *base) { return reinterpret_cast<double2 *>(base); }
    DEVINL static double2 *yt1(char *base) { return reinterpret_cast<double2 *>(base) + N; }
    DEVINL static double2 *yt2(char *base) { return reinterpret_cast<double2 *>(base) + 2 * N; }
    DEVINL static double2 *wn(char *base) { return reinterpret_cast<double2 *>(base) + 3 * N; }
    DEVINL static double *iv(char *base) { return reinterpret_cast<double *>(base + 64 * (size_t)N); }
};

// =================================================================================================
// FUSEB (lean layout only): the tile first runs phase B of the PREVIOUS step on its particles (gather of E_pred at the stored
// predicted samples, gy, the tau* sums, compute_v) and hands the new v straight to this step's preparation -- k_onepass_b and
// k_onepass_a of consecutive steps in one kernel.  Phase A alone leaves the LSU data pipe at 60 % and the fp64 pipe at 46 %
// outside its own gather (measured with the tap loads compiled out), phase B alone is bound by its 36 tap loads: in one
// kernel B's loads overlap A's transforms the way A's own gather already does.
template <int G, bool FULL, int SCHEME, bool FUSEB>
__global__ void __launch_bounds__(kOpBlockA, UAPIC_OP_MINB_A) k_onepass_a(OpDev D) {
    constexpr int N = 8 * G, PW = 32 / G, PPI = 32 / N;     // particles per warp tile / per gather iteration
    constexpr int kOpWarps = kOpBlockA / 32;
    using SM = StoreMap<N, FULL>;
    const OnepassParams &P = D.p;
    extern __shared__ double2 smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *gx = smem + kTab + wib * kWarpSmA;    // [8][kRow]  positions -> E at the samples
    double *sps = reinterpret_cast<double *>(gx + 8 * kRow);   // [8][kRow] doubles: sin(xt1) sin(xt2) at the samples (row stride 36: conflict-free both ways)
    double2 *yhs = gx + 12 * kRow;                 // [16][32]   yhat1[k1] at slot k1, yhat2[k1] at slot 8+k1
    constexpr bool TM = UAPIC_OP_TMEM_TABLES != 0;
    OpLane<G, TM> L; L.init(lane, smem);
    __shared__ unsigned tmem_base;
    if (TM) L.tmem_setup(&tmem_base);
    const int g = L.g, pin = lane / G, gbase = lane - g;
    const double eps = P.eps, inv_eps = D.inv_eps, invN = 1.0 / (double)N;
    const int64_t ntiles = (P.np + PW - 1) / PW;
#if UAPIC_OP_BLOCKED_A
    // every CTA owns a CONTIGUOUS range of tiles (experiment: did NOT help phase A -- the strided sweep, where all CTAs write
    // one moving window of the store, is 6 % faster)
    const int64_t tiles_per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t_first = (int64_t)blockIdx.x * tiles_per_cta, t_end = min(t_first + tiles_per_cta, ntiles);
    const int64_t nwarps = kOpWarps;
#else
    const int64_t t_first = (int64_t)blockIdx.x * kOpWarps, t_end = ntiles;
    const int64_t nwarps = (int64_t)gridDim.x * kOpWarps;
#endif

    // the trip count is uniform over the CTA (a warp past the end works on a clamped copy of the last particle with its
    // stores and deposits switched off) so that the warps of a CTA can be kept in step: they then share instruction
    // fetches of this long straight-line kernel (UAPIC_OP_LOCKSTEP)
    // the particle data of a tile are loaded while the previous tile is finishing (its deposits hide the latency)
    double2 xx_n, vv_n, ee_n;
    {
        const int64_t k0 = (t_first + wib) * PW + pin;
        const int64_t i0 = (t_first + wib < t_end && k0 < P.np) ? k0 : P.np - 1;
        xx_n = P.x[i0]; vv_n = P.v[i0]; ee_n = P.ep[i0];
    }
    for (int64_t tbase = t_first; tbase < t_end; tbase += nwarps) {
        const int64_t tile = tbase + wib;
        const int64_t kraw = tile * PW + pin;
        const bool valid = tile < t_end && kraw < P.np;
        const int64_t ip = valid ? kraw : P.np - 1;
        const double2 xx = xx_n, ee = ee_n;
        double2 vv = vv_n;
        if (FUSEB && !FULL) {
            // ================= phase B of the previous step for this tile (k_onepass_b, same arithmetic) =================
            double2 *wx = gx, *red = yhs;              // the exchange rows / the yhat stash are free until the preparation
            {
                const double2 *rec = reinterpret_cast<const double2 *>(P.rec + 8 * ip);
                const double tb = OP_LDS(rec + 1).x, rtb = 1.0 / tb;
                const double2 rc2 = OP_LDS(rec + 2), rc3 = OP_LDS(rec + 3);
                const cd e1b = mk(rc2.y, -rc3.x);
                cd eltb[8], wv[8];
                elt_modes(L, tb, eps, e1b, eltb);
#pragma unroll
                for (int k1 = 0; k1 < 8; ++k1) {
                    cd pl, qt;
                    pl_qt(L, k1, tb, rtb, eps, eltb[k1], pl, qt);
                    const cd w = cmulc(qt, eltb[k1]);
                    wv[k1] = mk(invN * w.re, -invN * w.im);
                }
                bwdN(wv, L);
                __syncwarp();
#pragma unroll
                for (int s = 0; s < 8; ++s) wx[s * kRow + lane] = make_double2(wv[s].re, -wv[s].im);
                __syncwarp();
            }
            {
                const int n = lane & (N - 1);
                const double2 csn = L.cs[n];
#pragma unroll (kGatherUnrollB)
                for (int j = 0; j < 8; ++j) {
                    const int pp = j * PPI + lane / N;
                    const int64_t kb = tile * PW + pp;
                    const int64_t ib = (tile < t_end && kb < P.np) ? kb : P.np - 1;
                    const double2 r0 = OP_LDS(reinterpret_cast<const double2 *>(P.rec + 8 * ib));
                    const double bb = r0.x, rbb = r0.y;
                    char *sb = P.store + (size_t)ib * SM::stride;
                    const double2 xs = OP_LDS(SM::xtr(sb) + n), ya = OP_LDS(SM::yt1(sb) + n), yb = OP_LDS(SM::yt2(sb) + n);
                    const double2 wn = wx[(n / G) * kRow + pp * G + (n & (G - 1))];
                    const double iv = (1.0 + 0.5 * sin(xs.x) * sin(xs.y) - bb) * inv_eps;      // ua_steps.F90:177
                    double xwb, ywb, eb1, eb2;
                    const Cell cell = cell_fast(P.m, D.f, xs.x, xs.y, P.wrap, xwb, ywb);
                    if (SCHEME == kSchemeCic) gather_cic_tiled(P.m, P.ehalo_b, cell, eb1, eb2); else OP_GATHER_M6(P.m, P.ehalo_b, cell, eb1, eb2);
                    cd gy1, gy2;
                    fy_time(csn.x, csn.y, rbb, iv, mk(ya.x, ya.y), mk(yb.x, yb.y), eb1, eb2, gy1, gy2);     // :177-183
                    red[pp * N + (n & ~7) + (((n & 7) + (n >> 3) + 4 * (pp & 1)) & 7)] =
                        make_double2(fma(gy1.re, wn.x, -gy1.im * wn.y), fma(gy2.re, wn.x, -gy2.im * wn.y));
                }
            }
            __syncwarp();
            {
                double sx = 0.0, sy = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double2 q = red[pin * N + 8 * g + ((i + g + 4 * (pin & 1)) & 7)];
                    sx += q.x; sy += q.y;
                }
                sx = grp_sum<G>(sx); sy = grp_sum<G>(sy);
                const double2 *rec = reinterpret_cast<const double2 *>(P.rec + 8 * ip);
                const double2 r1 = OP_LDS(rec + 1), r2 = OP_LDS(rec + 2), r3 = OP_LDS(rec + 3);
                const double px = r1.y + sx, py = r2.x + sy, cs = r2.y, sn = r3.x;
                vv = make_double2(cs * px + sn * py, cs * py - sn * px);                           // :302-303, every lane of the particle
                if (valid && g == 0) { if (P.out_perm) P.v_out[P.out_perm[ip]] = vv; else P.v[ip] = vv; }
            }
            OP_SYNC(0);        // the exchange rows are reused by the preparation below
        }
        const double x1 = xx.x, x2 = xx.y, vx = vv.x, vy = vv.y;
#if UAPIC_OP_PREFETCH
        {   // the CTA's next tiles: their particle data would otherwise be a cold HBM access all warps of the CTA wait on together
            const int64_t nk = kraw + nwarps * PW;
            if (g == 0 && nk < P.np && tile + nwarps < t_end) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.x + nk));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.v + nk));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.ep + nk));
            }
        }
#endif

        // ---- preparation (ua_steps.F90:49-113) ----
        const double b = bfield(x1, x2);                                     // :54
        const double t = P.dt * b;                                           // :55
        const double rb = 1.0 / b, rt = 1.0 / t;
        const double vxb = vx * rb, vyb = vy * rb;                           // :73-74
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const double2 c = L.cs_s(s);
            const double xt1 = x1 + eps * (c.y * vxb - c.x * vyb) + eps * vyb;   // :78-81
            const double xt2 = x2 + eps * (c.y * vyb + c.x * vxb) - eps * vxb;   // :79-82
            gx[s * kRow + lane] = make_double2(xt1, xt2);
        }
        OP_SYNC(0);
        // ---- gather E_n at the samples, lane = sample (interpolation_m6.F90:83-189); the sines of :87 ride along so that
        //      the 16 sin() of a lane sit in this rolled loop instead of the unrolled code below ----
#pragma unroll (kGatherUnroll)
        for (int j = 0; j < 8; ++j) {
            const int n = lane & (N - 1), pp = j * PPI + lane / N;
            const int col = pp * G + (n & (G - 1)), row = n / G;
            const double2 pos = gx[row * kRow + col];
            double xw, yw, e1, e2;
            const Cell cell = cell_fast(P.m, D.f, pos.x, pos.y, P.wrap, xw, yw);
            if (SCHEME == kSchemeCic) gather_cic_tiled(P.m, P.ehalo, cell, e1, e2); else OP_GATHER_M6(P.m, P.ehalo, cell, e1, e2);
            gx[row * kRow + col] = make_double2(e1, e2);
            sps[row * kRow + col] = sin(pos.x) * sin(pos.y);
        }
        OP_SYNC(1);

        cd z[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const double2 c = L.cs_s(s);
            const double iv = (1.0 + 0.5 * sps[s * kRow + lane] - b) * inv_eps;       // :87
            const double exb = ((c.x * vy - c.y * vx) * iv + ee.x) * rb;                    // :89
            const double eyb = ((-c.x * vx - c.y * vy) * iv + ee.y) * rb;                   // :90
            z[s] = mk(c.x * exb - c.y * eyb, c.y * exb + c.x * eyb);                        // r1 + i r2   :92-93
        }
        fwdN(z, L);                                                       // :97-98, both real signals at once
        // split the two spectra, filter (:100-103; the k = 0 coefficient cancels in :109-110 and is dropped)
        {
            cd zp[8];
            zp[0] = shfl_grp<G>(z[0], L.src_neg0);
#pragma unroll
            for (int k1 = 1; k1 < 8; ++k1) zp[k1] = shfl_grp<G>(z[8 - k1], L.src_neg);
            cd c1[8], c2[8], cs1 = mk(0.0, 0.0), cs2 = mk(0.0, 0.0);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const double s = 0.5 * invN * L.ilv(k1).x;
                c1[k1] = mk(s * (z[k1].im - zp[k1].im), -s * (z[k1].re + zp[k1].re));
                c2[k1] = mk(-s * (z[k1].re - zp[k1].re), -s * (z[k1].im + zp[k1].im));
                cs1 = cadd(cs1, c1[k1]); cs2 = cadd(cs2, c2[k1]);
            }
            cs1 = mk(grp_sum<G>(cs1.re), grp_sum<G>(cs1.im));                // value of the filtered signal at tau = 0
            cs2 = mk(grp_sum<G>(cs2.re), grp_sum<G>(cs2.im));
            // FFT(yt)/N: eps*c_k, and the mean value for k = 0   (:109-110)
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                cd y1 = rmul(eps, c1[k1]), y2 = rmul(eps, c2[k1]);
                if (k1 == 0 && L.kap == 0) { y1 = mk(vx - eps * cs1.re, -eps * cs1.im); y2 = mk(vy - eps * cs2.re, -eps * cs2.im); }
                yhs[k1 * 32 + lane] = make_double2(y1.re, y1.im);
                yhs[(8 + k1) * 32 + lane] = make_double2(y2.re, y2.im);
            }
        }
        OP_SYNC(2);

        // ---- exp(-i l t/eps), pl, ql/t, w = ql/t * conj(elt) for the lane's modes ----
        cd e1;
        { double s, c; sincos(-t / eps, &s, &c); e1 = mk(c, s); }
        cd elt[8];
        elt_modes(L, t, eps, e1, elt);

        // ---- x: fhat_x from yhat (a +-1 shift in Fourier space), ua_step1 (:226-227), position sums ----
        double posp1 = 0.0, posp2 = 0.0, swf1 = 0.0, swf2 = 0.0;
        cd X1[8], X2[8];
        {
            const double hb = 0.5 * rb, he = 0.5 * eps;
            const double2 *y1s = yhs + lane, *y2s = yhs + 8 * 32 + lane;
            // sliding window over the modes k-1, k, k+1 (neighbouring blocks live in other lanes' slots)
            double2 a1 = yhs[7 * 32 + gbase + L.src_m1], a2 = yhs[15 * 32 + gbase + L.src_m1];
            double2 b1 = y1s[0], b2 = y2s[0];
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const double2 n1 = (k1 < 7) ? y1s[(k1 + 1) * 32] : yhs[gbase + L.src_p1];
                const double2 n2 = (k1 < 7) ? y2s[(k1 + 1) * 32] : yhs[8 * 32 + gbase + L.src_p1];
                cd fx1, fx2;
                fx_modes(hb, mk(a1.x, a1.y), mk(n1.x, n1.y), mk(a2.x, a2.y), mk(n2.x, n2.y), fx1, fx2);
                cd pl, qt;
                pl_qt(L, k1, t, rt, eps, elt[k1], pl, qt);
                const cd w = cmulc(qt, elt[k1]);
                // FFT(xt)/N of the first-order profile: modes 0, +-1 only
                cd xh1 = mk(0.0, 0.0), xh2 = mk(0.0, 0.0);
                if (k1 == 0 && L.kap == 0) { xh1 = mk(x1 + eps * vyb, 0.0); xh2 = mk(x2 - eps * vxb, 0.0); }
                if (k1 == 1 && L.kap == 0) { xh1 = mk(-he * vyb, -he * vxb); xh2 = mk(he * vxb, -he * vyb); }
                if (k1 == 7 && L.kap == G - 1) { xh1 = mk(-he * vyb, he * vxb); xh2 = mk(he * vxb, he * vyb); }
                if (k1 == 0 || k1 == 1 || k1 == 7) {
                    X1[k1] = cfma(pl, fx1, cmul(elt[k1], xh1));                  // :226
                    X2[k1] = cfma(pl, fx2, cmul(elt[k1], xh2));                  // :227
                } else {
                    X1[k1] = cmul(pl, fx1);
                    X2[k1] = cmul(pl, fx2);
                }
                posp1 += re_mulc(X1[k1], elt[k1]);                               // compute_rho_m6.F90:74-84
                posp2 += re_mulc(X2[k1], elt[k1]);
                swf1 += re_mul(w, fx1);
                swf2 += re_mul(w, fx2);
                a1 = b1; a2 = b2; b1 = n1; b2 = n2;
            }
        }
        bwdN(X1, L);                                                      // :231
        bwdN(X2, L);
        char *sbase = P.store + (size_t)ip * SM::stride;
        if (valid) {
#pragma unroll
            for (int s = 0; s < 8; ++s) OP_STS(SM::xtr(sbase) + g + G * s, make_double2(X1[s].re, X2[s].re));
            if (FULL) {
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    SM::iv(sbase)[g + G * s] = (1.0 + 0.5 * sin(X1[s].re) * sin(X2[s].re) - b) * inv_eps;   // :177 of the 2nd compute_f
            }
        }

        // ---- y: yt = B(yhat) (:105-110), fy in the time domain (:177-183), FFT, ua_step1 (:226), bracket sums ----
        OP_SYNC(3);
        // yhat1, yhat2 are spectra of real signals (:92-106) except for two coefficients: the Nyquist mode, which the filter
        // -i/l turns purely imaginary, and the mean, which carries minus its value (:109-110).  So ONE backward transform of
        // H = H1 + i H2 (H: Hermitian parts) gives Re yt1 + i Re yt2, and Im yt_n = Im yhat_0 + (-1)^n Im yhat_{N/2}.
        cd y1[8], y2[8];
        {
            constexpr int kNyqLane = OpLane<G, TM>::kappa_of(G / 2), kNyqReg = (G == 1) ? 4 : 0;
            const double2 m1 = yhs[gbase], m2 = yhs[8 * 32 + gbase];                                   // yhat_0 (lane kappa = 0)
            const double2 q1 = yhs[kNyqReg * 32 + gbase + kNyqLane], q2 = yhs[(8 + kNyqReg) * 32 + gbase + kNyqLane];   // yhat_{N/2}
            cd h[8];
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const double2 a = yhs[k1 * 32 + lane], c = yhs[(8 + k1) * 32 + lane];
                h[k1] = mk(a.x - c.y, a.y + c.x);
                if (k1 == 0 && L.kap == 0) h[k1] = mk(a.x, c.x);
                if (k1 == kNyqReg && L.kap == G / 2) h[k1] = mk(0.0, 0.0);
            }
            bwdN(h, L);
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const bool odd = (G == 1) ? (s & 1) : (g & 1);                 // parity of n = g + G s
                y1[s] = mk(h[s].re, odd ? m1.y - q1.y : m1.y + q1.y);
                y2[s] = mk(h[s].im, odd ? m2.y - q2.y : m2.y + q2.y);
            }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const double2 c = L.cs_s(s), et = gx[s * kRow + lane];
            const double iv = (1.0 + 0.5 * sps[s * kRow + lane] - b) * inv_eps;       // :177, same as :87
            fy_time(c.x, c.y, rb, iv, y1[s], y2[s], et.x, et.y, y1[s], y2[s]);
        }
        OP_SYNC(4);
        fwdN(y1, L);                                                      // :189-190
        fwdN(y2, L);
        double qa1 = 0.0, qa2 = 0.0;
        cd wv[8];
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const cd fy1 = rmul(invN, y1[k1]), fy2 = rmul(invN, y2[k1]);     // :194-195
            const double2 h1 = yhs[k1 * 32 + lane], h2 = yhs[(8 + k1) * 32 + lane];
            cd pl, qt;
            pl_qt(L, k1, t, rt, eps, elt[k1], pl, qt);
            wv[k1] = cmulc(qt, elt[k1]);
            const cd yp1 = cfma(pl, fy1, cmul(elt[k1], mk(h1.x, h1.y)));     // :226 for y
            const cd yp2 = cfma(pl, fy2, cmul(elt[k1], mk(h2.x, h2.y)));
            qa1 += re_mulc(yp1, elt[k1]) - re_mul(wv[k1], fy1);              // Re[(Yp - qt Fy) conj(elt)]
            qa2 += re_mulc(yp2, elt[k1]) - re_mul(wv[k1], fy2);
            y1[k1] = yp1; y2[k1] = yp2;
        }
        // corrector position (:260 + compute_rho_m6.F90:74-84): needs ghat_x = shift of the predicted yhat only
        double swg1 = 0.0, swg2 = 0.0;
        {
            const double hb = 0.5 * rb;
            const cd m1 = shfl_grp<G>(y1[7], L.src_m1), m2 = shfl_grp<G>(y2[7], L.src_m1);
            const cd p1 = shfl_grp<G>(y1[0], L.src_p1), p2 = shfl_grp<G>(y2[0], L.src_p1);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                cd gx1, gx2;
                fx_modes(hb, k1 == 0 ? m1 : y1[k1 - 1], k1 == 7 ? p1 : y1[k1 + 1], k1 == 0 ? m2 : y2[k1 - 1], k1 == 7 ? p2 : y2[k1 + 1],
                         gx1, gx2);
                swg1 += re_mul(wv[k1], gx1);
                swg2 += re_mul(wv[k1], gx2);
            }
        }
        OP_SYNC(5);
        bwdN(y1, L);                                                      // :232
        bwdN(y2, L);
        if (valid) {
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                OP_STS(SM::yt1(sbase) + g + G * s, make_double2(y1[s].re, y1[s].im));
                OP_STS(SM::yt2(sbase) + g + G * s, make_double2(y2[s].re, y2[s].im));
            }
        }
        if (FULL) {
            // W_n = (1/N) sum_k w_k exp(-i k tau_n) = conj( B(conj(w)) ) / N
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) wv[k1] = mk(invN * wv[k1].re, -invN * wv[k1].im);
            bwdN(wv, L);
            if (valid) {
#pragma unroll
                for (int s = 0; s < 8; ++s) SM::wn(sbase)[g + G * s] = make_double2(wv[s].re, -wv[s].im);
            }
        }

        {   // next tile of this warp (clamped like the current one)
            const int64_t kn = kraw + nwarps * PW;
            const int64_t in = (tile + nwarps < t_end && kn < P.np) ? kn : P.np - 1;
            xx_n = OP_LDS(P.x + in); vv_n = OP_LDS(P.v + in); ee_n = OP_LDS(P.ep + in);
        }
        // ---- both deposit positions, deposits, per-particle record ----
        posp1 = grp_sum<G>(posp1); posp2 = grp_sum<G>(posp2);
        const double posc1 = posp1 + grp_sum<G>(swg1 - swf1), posc2 = posp2 + grp_sum<G>(swg2 - swf2);
        qa1 = grp_sum<G>(qa1); qa2 = grp_sum<G>(qa2);
        double xw, yw;
        // the CTAs of a sweep all work in the same sorted bin, i.e. on the same few hundred mesh nodes: spread their atomics
        // over P.rho_copies private copies of the two raw meshes (folded before the field solve)
        RhoAcc rp = P.rho_p, rc = P.rho_c;
        {
            const size_t off = (size_t)(blockIdx.x % P.rho_copies) * 2 * (size_t)P.m.ld * (P.m.ny + 1);
            if (rp.f64) { rp.f64 += off; rc.f64 += off; } else { rp.i64 += off; rc.i64 += off; }
        }
#define rho_p_sel rp
#define rho_c_sel rc
#ifndef UAPIC_OP_NO_DEPOSIT     // measurement only (profiles/README.md): phase A without its 72 deposit atomics per particle
#define UAPIC_OP_NO_DEPOSIT 0
#endif
#if !UAPIC_OP_NO_DEPOSIT
        const Cell cp = cell_fast(P.m, D.f, posp1, posp2, P.wrap, xw, yw);
        if (SCHEME == kSchemeCic) {
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4 / G; ++q) deposit_cic_tap(P.m, rho_p_sel, cp, P.weight, g * (4 / G) + q);
            }
        } else {
            deposit_split<G>(P.m, rho_p_sel, cp, P.weight, g, valid);          // compute_rho_m6.F90:89-187 (predictor)
        }
        const Cell cc = cell_fast(P.m, D.f, posc1, posc2, P.wrap, xw, yw);
        if (SCHEME == kSchemeCic) {
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4 / G; ++q) deposit_cic_tap(P.m, rho_c_sel, cc, P.weight, g * (4 / G) + q);
            }
        } else {
            deposit_split<G>(P.m, rho_c_sel, cc, P.weight, g, valid);          // (corrector)
        }
#else
        const Cell cp = cell_fast(P.m, D.f, posp1, posp2, P.wrap, xw, yw);
        const Cell cc = cell_fast(P.m, D.f, posc1, posc2, P.wrap, xw, yw);
        if (cp.i + cc.i == -12345) rho_add(rp, 0, xw);      // keeps the position sums alive
#endif
        if (valid && g == 0) {
            if (P.out_perm) P.x_out[P.out_perm[ip]] = make_double2(xw, yw);
            else P.x[ip] = make_double2(xw, yw);                             // compute_rho_m6.F90:86-87
            double2 *rec = reinterpret_cast<double2 *>(P.rec + 8 * ip);
            rec[0] = make_double2(b, rb);                                    // what every sample lane of phase B needs
            rec[1] = make_double2(t, qa1);
            rec[2] = make_double2(qa2, e1.re);                               // cos(t/eps)
            rec[3] = make_double2(-e1.im, 0.0);                              // sin(t/eps)
        }
        OP_SYNC(6);
    }
    if (TM) L.tmem_release(&tmem_base);
}

// =================================================================================================
template <int G, bool FULL, int SCHEME>
__global__ void __launch_bounds__(kOpBlockB, UAPIC_OP_MINB_B) k_onepass_b(OpDev D) {
    constexpr int N = 8 * G, PW = 32 / G, PPI = 32 / N;
    constexpr int kOpWarps = kOpBlockB / 32;
    using SM = StoreMap<N, FULL>;
    const OnepassParams &P = D.p;
    extern __shared__ double2 smem[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *wx = smem + kTab + wib * kWarpSmB;    // [8][kRow]  W_n of the tile (LEAN only)
    double2 *red = wx + 8 * kRow;                  // [PW][N]    per-sample terms of the two tau* sums
    OpLane<G> L; L.init(lane, smem);
    const int g = L.g, pin = lane / G;
    const double eps = P.eps, inv_eps = D.inv_eps, invN = 1.0 / (double)N;
    const int n = lane & (N - 1);
    const double2 csn = L.cs[n];
    const int64_t ntiles = (P.np + PW - 1) / PW;
#if UAPIC_OP_BLOCKED_B   // a contiguous range of tiles per CTA: measured -2 % for phase B (+6 % for phase A, which keeps the strided sweep)
    const int64_t tiles_per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t_first = (int64_t)blockIdx.x * tiles_per_cta, t_end = min(t_first + tiles_per_cta, ntiles);
    const int64_t nwarps = kOpWarps;
#else
    const int64_t t_first = (int64_t)blockIdx.x * kOpWarps, t_end = ntiles;
    const int64_t nwarps = (int64_t)gridDim.x * kOpWarps;
#endif

    for (int64_t tile = t_first + wib; tile < t_end; tile += nwarps) {
        if (!FULL) {
            // W_n of the tile's particles in the phase-A layout, handed to the sample lanes through shared memory
            const int64_t kraw = tile * PW + pin;
            const int64_t ip = kraw < P.np ? kraw : P.np - 1;
            const double2 *rec = reinterpret_cast<const double2 *>(P.rec + 8 * ip);
            const double t = rec[1].x, rt = 1.0 / t;
            const cd e1 = mk(rec[2].y, -rec[3].x);
            cd elt[8], wv[8];
            elt_modes(L, t, eps, e1, elt);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                cd pl, qt;
                pl_qt(L, k1, t, rt, eps, elt[k1], pl, qt);
                const cd w = cmulc(qt, elt[k1]);
                wv[k1] = mk(invN * w.re, -invN * w.im);
            }
            bwdN(wv, L);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 8; ++s) wx[s * kRow + lane] = make_double2(wv[s].re, -wv[s].im);
            __syncwarp();
        }
#pragma unroll (kGatherUnrollB)
        for (int j = 0; j < 8; ++j) {
            const int pp = j * PPI + lane / N;
            const int64_t kraw = tile * PW + pp;
            const bool valid = kraw < P.np;
            const int64_t ip = valid ? kraw : P.np - 1;
            const double2 r0 = OP_LDS(reinterpret_cast<const double2 *>(P.rec + 8 * ip));
            const double b = r0.x, rb = r0.y;
            char *sbase = P.store + (size_t)ip * SM::stride;
#if UAPIC_OP_PREFETCH_B
            {   // the store is streamed exactly once: ask L2 for the lines of the particle two iterations ahead
                const int j2 = j + UAPIC_OP_PREFETCH_B;
                const int64_t k2 = (j2 < 8 ? tile : tile + nwarps) * PW + (j2 & 7) * PPI + lane / N;
                if ((n & 7) == 0 && k2 < P.np && (j2 < 8 || tile + nwarps < t_end)) {
                    const char *b2 = P.store + (size_t)k2 * SM::stride + 16 * n;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(b2));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + 16 * N));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + 32 * N));
                }
            }
#endif
            const double2 xs = OP_LDS(SM::xtr(sbase) + n), ya = OP_LDS(SM::yt1(sbase) + n), yb = OP_LDS(SM::yt2(sbase) + n);
            double2 wn;
            double iv;
            if (FULL) {
                wn = SM::wn(sbase)[n];
                iv = SM::iv(sbase)[n];
            } else {
                wn = wx[(n / G) * kRow + pp * G + (n & (G - 1))];
                iv = (1.0 + 0.5 * sin(xs.x) * sin(xs.y) - b) * inv_eps;      // ua_steps.F90:177
            }
            // ---- gather E_pred at the predicted samples (interpolation_m6.F90:83-189) ----
            double xw, yw, e1, e2;
            const Cell cell = cell_fast(P.m, D.f, xs.x, xs.y, P.wrap, xw, yw);
            if (SCHEME == kSchemeCic) gather_cic_tiled(P.m, P.ehalo, cell, e1, e2); else OP_GATHER_M6(P.m, P.ehalo, cell, e1, e2);
            cd gy1, gy2;
            fy_time(csn.x, csn.y, rb, iv, mk(ya.x, ya.y), mk(yb.x, yb.y), e1, e2, gy1, gy2);     // :177-183
            // terms of Re sum_n gy(tau_n) W_n: summed over n after the loop (one pass through shared memory instead of a
            // shuffle tree per particle)
            // (column swizzle: the reader below takes its 8 terms in natural order, which keeps a particle's result independent
            //  of the slot it sits in, and still hits 8 distinct bank groups per quarter warp)
            red[pp * N + (n & ~7) + (((n & 7) + (n >> 3) + 4 * (pp & 1)) & 7)] =
                make_double2(fma(gy1.re, wn.x, -gy1.im * wn.y), fma(gy2.re, wn.x, -gy2.im * wn.y));
        }
        __syncwarp();
        {
            // lane (p, g) of the phase-A layout adds the 8 consecutive terms n = 8g .. 8g+7 of particle p in this fixed order,
            // then a log2(G)-stage butterfly; lane g = 0 finishes compute_v (ua_steps.F90:293-303)
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double2 q = red[pin * N + 8 * g + ((i + g + 4 * (pin & 1)) & 7)];
                sx += q.x; sy += q.y;
            }
            sx = grp_sum<G>(sx); sy = grp_sum<G>(sy);
            const int64_t kraw = tile * PW + pin;
            if (g == 0 && kraw < P.np) {
                const double2 *rec = reinterpret_cast<const double2 *>(P.rec + 8 * kraw);
                const double2 r1 = rec[1], r2 = rec[2], r3 = rec[3];
                const double px = r1.y + sx, py = r2.x + sy, cs = r2.y, sn = r3.x;
                const double2 vn = make_double2(cs * px + sn * py, cs * py - sn * px);             // :302-303
                if (P.out_perm) P.v_out[P.out_perm[kraw]] = vn; else P.v[kraw] = vn;
            }
        }
        __syncwarp();
    }
}

inline int op_grid(const LaunchCtx &c, int64_t np, int ntau, int minb, int block) {
    const int pw = 32 / (ntau / 8), warps = block / 32;
    int64_t need = ((np + pw - 1) / pw + warps - 1) / warps;
    const int64_t cap = (int64_t)c.sm_count * minb;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

OpDev make_opdev(const OnepassParams &p) {
    OpDev D;
    D.p = p;
    D.f.inv_dx = 1.0 / p.m.dx; D.f.inv_dy = 1.0 / p.m.dy;
    D.f.inv_nx = 1.0 / (double)p.m.nx; D.f.inv_ny = 1.0 / (double)p.m.ny;
    D.f.inv_dimx = 1.0 / p.m.dimx; D.f.inv_dimy = 1.0 / p.m.dimy;
    D.inv_eps = 1.0 / p.eps;
    return D;
}

template <typename K> cudaError_t op_launch(K kernel, const LaunchCtx &c, const OpDev &D, int grid, int block, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, block, smem, c.stream>>>(D);
    if (c.launches) *c.launches += 1;
    return cudaGetLastError();
}

}  // namespace

bool onepass_ntau_supported(int ntau) { return ntau == 8 || ntau == 16 || ntau == 32; }
size_t onepass_store_bytes_per_particle(int ntau, int full) { return (size_t)(full ? 72 : 48) * (size_t)ntau; }

#define COMMA_TRUE , true
#define COMMA_FALSE , false
#define UAPIC_OP_LAUNCH(KERNEL, GG, MINB, BLOCK, SMEM, ...)                                                                       \
    (p.scheme == kSchemeCic ? op_launch(KERNEL<GG, false, kSchemeCic __VA_ARGS__>, c, D, op_grid(c, p.np, 8 * GG, MINB, BLOCK), BLOCK, SMEM)     \
     : p.full             ? op_launch(KERNEL<GG, true, kSchemeM6 __VA_ARGS__>, c, D, op_grid(c, p.np, 8 * GG, MINB, BLOCK), BLOCK, SMEM)       \
                          : op_launch(KERNEL<GG, false, kSchemeM6 __VA_ARGS__>, c, D, op_grid(c, p.np, 8 * GG, MINB, BLOCK), BLOCK, SMEM))
#define UAPIC_OP_DISPATCH(KERNEL, MINB, BLOCK, SMEM, ...)                  \
    if (p.scheme == kSchemeCic && p.full) return cudaErrorInvalidValue; \
    switch (p.ntau) {                                                 \
        case 8:  return UAPIC_OP_LAUNCH(KERNEL, 1, MINB, BLOCK, SMEM, __VA_ARGS__); \
        case 16: return UAPIC_OP_LAUNCH(KERNEL, 2, MINB, BLOCK, SMEM, __VA_ARGS__); \
        case 32: return UAPIC_OP_LAUNCH(KERNEL, 4, MINB, BLOCK, SMEM, __VA_ARGS__); \
        default: return cudaErrorInvalidValue;                        \
    }

cudaError_t launch_onepass_a(const LaunchCtx &c, const OnepassParams &p) {
    if (p.np <= 0) return cudaSuccess;
    if (p.m.nx < 4 || p.m.ny < 4) return cudaErrorInvalidValue;
    const OpDev D = make_opdev(p);
    const size_t smem = sizeof(double2) * (size_t)(kTab + (kOpBlockA / 32) * kWarpSmA);
    if (p.fuse_b) {
        if (p.full || !p.ehalo_b) return cudaErrorInvalidValue;      // the fused kernel exists for the lean layout only
        UAPIC_OP_DISPATCH(k_onepass_a, UAPIC_OP_MINB_A, kOpBlockA, smem, COMMA_TRUE)
    }
    UAPIC_OP_DISPATCH(k_onepass_a, UAPIC_OP_MINB_A, kOpBlockA, smem, COMMA_FALSE)
}

cudaError_t launch_onepass_b(const LaunchCtx &c, const OnepassParams &p) {
    if (p.np <= 0) return cudaSuccess;
    if (p.m.nx < 4 || p.m.ny < 4) return cudaErrorInvalidValue;
    const OpDev D = make_opdev(p);
    const size_t smem = sizeof(double2) * (size_t)(kTab + (kOpBlockB / 32) * kWarpSmB);
    UAPIC_OP_DISPATCH(k_onepass_b, UAPIC_OP_MINB_B, kOpBlockB, smem, )
}

}  // namespace uapic

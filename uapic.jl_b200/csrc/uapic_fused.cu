// uapic_fused.cu -- the two fused phase kernels of the session (performance) path, sm_100a.
//
//   phase A = preparation + gather(E_old) + compute_f + ua_step1 (x and y) + predictor deposit
//             ua_steps.F90:15-236, interpolation_m6.F90:40-191, compute_rho_m6.F90:47-189
//   phase B = gather(E_new) + compute_f + ua_step2 (x and y) + corrector deposit + compute_v
//             ua_steps.F90:117-307
//
// Between them sits the only global dependency of the scheme (deposit -> Poisson -> gather).  What crosses it
// lives in HBM as 8 complex arrays per particle, tau fastest (128 B per particle-tau):
//   [k][0..3][n] = predicted xt1, xt2, yt1, yt2 (time domain, natural order)
//   [k][4..7][s] = fhat_x1, fhat_x2, fhat_y1, fhat_y2 (tau-Fourier, slot order = lane order, normalised by 1/N)
// A warp writes/reads 8 x 16 x N contiguous bytes per particle with 16-byte accesses.
//
// Arithmetic is algebraically the reference's, reorganised to cut fp64 work (the kernels are fp64-pipe / issue
// bound, not HBM bound -- see DESIGN.md section 5):
//   * FFT(xt) of the first-order profile is written down analytically (only modes 0, +-1 are non-zero);
//   * FFT(yt) is available from the filtered coefficients of `preparation` without a transform;
//   * the two real signals r1, r2 of `preparation` go through ONE complex FFT and are separated afterwards;
//   * fx = R(-tau) yt / b is a +-1 shift in tau-Fourier space, so fhat_x needs lane permutations, not FFTs;
//   * FFT(IFFT(xhat))/N == xhat: the tau* evaluation of the deposits reads the Fourier coefficients directly;
//   * elt/N*xf + pl*fhat == FFT(predicted xt)/N: the corrector needs no stored xf/yf;
//   * every division by a loop-invariant becomes a multiplication by its reciprocal (the phase l*t/eps keeps
//     its true division: at small eps it is the one place where 1 ulp is amplified by t/eps).
// That is 15 length-N FFTs per particle per step instead of the reference's 30.
#include "uapic_internal.h"
#include "uapic_fast.cuh"

namespace uapic {

namespace {

#ifndef UAPIC_PHASE_BLOCK
#define UAPIC_PHASE_BLOCK 256
#endif
#ifndef UAPIC_PHASE_MINB
#define UAPIC_PHASE_MINB 2
#endif
constexpr int kPhaseBlock = UAPIC_PHASE_BLOCK;
constexpr int kPhaseMinBlocks = UAPIC_PHASE_MINB;


template <int N> struct FusedLane {
    static constexpr int LOG = Log2<N>::v;
    TauLane<N> T;
    unsigned sgn[LOG];            // 0x80000000 on the lower lane of each butterfly, per stage
    int src_neg, src_m1, src_p1;  // lanes (inside the group) holding Fourier slots -k, k-1, k+1
    double s1, sN;                // +-1 and +-1/N: minus on odd lanes (sign convention of fft_*_b below)
    // rarely used per-lane constants live in shared memory (they only depend on the lane): frees ~22 registers
    const double2 *tws;           // [LOG-1][32] signed twiddles of this lane
    const double *invs;           // [3][32]: 1/l, 1/l^2, 1/(l*N) of this lane's slot (0 for k = 0)
    int lane;
    DEVINL cd tw(int s) const { const double2 w = tws[s * 32 + lane]; return mk(w.x, w.y); }
    DEVINL double inv_l() const { return invs[lane]; }
    DEVINL double inv_l2() const { return invs[32 + lane]; }
    DEVINL double inv_lN() const { return invs[64 + lane]; }

    DEVINL static int lane_of(int freq) { return (int)(__brev((unsigned)(freq & (N - 1))) >> (32 - LOG)); }

    // sm_tw: 4*32 double2, sm_inv: 3*32 double, both __shared__; call from every thread, ends with __syncthreads
    DEVINL void init(int lane_, double2 *sm_tw, double *sm_inv) {
        lane = lane_;
        T.init(lane);
#pragma unroll
        for (int s = 0; s < LOG; ++s) sgn[s] = (T.j & (N >> (s + 1))) ? 0x80000000u : 0u;
        s1 = (T.j & 1) ? -1.0 : 1.0;
        sN = s1 / (double)N;
        src_neg = lane_of(N - T.k);
        src_m1 = lane_of(T.k - 1 + N);
        src_p1 = lane_of(T.k + 1);
        if (threadIdx.x < 32) {
            // lower lanes carry the NEGATED twiddle (see fft_fwd_b); upper lanes keep 1
#pragma unroll
            for (int s = 0; s < LOG - 1; ++s) {
                const double sg = (T.j & (N >> (s + 1))) ? -1.0 : 1.0;
                sm_tw[s * 32 + lane] = make_double2(sg * T.twr[s], sg * T.twi[s]);
            }
            const bool nz = T.k != 0;
            sm_inv[lane] = nz ? 1.0 / T.lf : 0.0;
            sm_inv[32 + lane] = nz ? 1.0 / (T.lf * T.lf) : 0.0;
            sm_inv[64 + lane] = nz ? 1.0 / (T.lf * (double)N) : 0.0;
        }
        tws = sm_tw; invs = sm_inv;
        __syncthreads();
    }
};

template <int N> DEVINL cd shfl_idx(cd v, int src) {
    return mk(__shfl_sync(kFull, v.re, src, N), __shfl_sync(kFull, v.im, src, N));
}

// shuffle a double with lane^h and flip its sign where `m` has the sign bit: the XOR lands on the freshly received
// high word, so no register-pair shuffling is needed
DEVINL double shfl_xor_flip(double v, int h, unsigned m) {
    const int lo = __shfl_xor_sync(kFull, __double2loint(v), h);
    const int hi = __shfl_xor_sync(kFull, __double2hiint(v), h) ^ (int)m;
    return __hiloint2double(hi, lo);
}

// B independent length-N transforms at once, interleaved for ILP.  Sign bookkeeping keeps every butterfly a plain
// add of the local value and the (possibly sign-flipped) received one:
//   forward (DIF): d = v + flipL(o), then times the lane's twiddle register, which holds -w on lower lanes
//                  (lower: (v - o)(-w) = (o - v) w).  The last stage has no twiddle to absorb the sign, so
//                  OUTPUTS ARE NEGATED ON ODD LANES: multiply them by L.s1 (or L.sN to normalise as well).
//   backward (DIT): v *= conj(twiddle register) (= -conj(w) on lower lanes), d = v - flipL(o).
//                  INPUTS MUST BE PRE-NEGATED ON ODD LANES (fold L.s1 into whatever produces them).
template <int N, int B> DEVINL void fft_fwd_b(cd (&v)[B], const FusedLane<N> &L) {
    constexpr int LOG = FusedLane<N>::LOG;
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int h = N >> (s + 1);
        const cd w = (h > 1) ? L.tw(s) : mk(1.0, 0.0);
#pragma unroll
        for (int q = 0; q < B; ++q) {
            cd d = mk(v[q].re + shfl_xor_flip(v[q].re, h, L.sgn[s]), v[q].im + shfl_xor_flip(v[q].im, h, L.sgn[s]));
            if (h > 1) d = cmul(d, w);
            v[q] = d;
        }
    }
}

template <int N, int B> DEVINL void fft_bwd_b(cd (&v)[B], const FusedLane<N> &L) {
    constexpr int LOG = FusedLane<N>::LOG;
#pragma unroll
    for (int s = LOG - 1; s >= 0; --s) {
        const int h = N >> (s + 1);
        const cd w = (h > 1) ? L.tw(s) : mk(1.0, 0.0);
#pragma unroll
        for (int q = 0; q < B; ++q) {
            cd a = v[q];
            if (h > 1) a = cmulc(a, w);
            v[q] = mk(a.re - shfl_xor_flip(a.re, h, L.sgn[s]), a.im - shfl_xor_flip(a.im, h, L.sgn[s]));
        }
    }
}

// Cooperative deposit of one particle by its N-lane group (compute_rho_m6.F90:89-187).  The position is uniform in the
// group; the 12 weights are computed once (one per lane), the 36 live taps are spread over the lanes, weights travel by
// shuffle.  Every lane of the warp must call this (shuffles inside); `valid` only guards the atomics.
template <int N>
DEVINL void deposit_coop(const MeshDev &m, const RhoAcc &r, const Cell &c, double weight, int j, bool valid) {
    constexpr int S = (12 + N - 1) / N;    // weight slots per lane
    constexpr int R = (36 + N - 1) / N;    // tap rounds
    double wreg[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int idx = j + s * N;                 // 0..5: x offsets -2..3 ; 6..11: y offsets -2..3
        const int axis = idx >= 6;
        const int off = idx - 6 * axis - 2;
        const double dp = axis ? c.dpy : c.dpx;
        const double q = (off <= 0) ? (double)(-off) + dp : (double)off - dp;
        wreg[s] = f_m6_branchless(fabs(q));
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        const int tap = j + rr * N;
        const bool active = tap < 36;
        const int tp = active ? tap : 0;
        const int a = tp / 6, b = tp - 6 * a;
        double cx = 0.0, cy = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double wa = __shfl_sync(kFull, wreg[s], a & (N - 1), N);
            const double wb = __shfl_sync(kFull, wreg[s], (6 + b) & (N - 1), N);
            if (a / N == s) cx = wa;
            if ((6 + b) / N == s) cy = wb;
        }
        if (active && valid) {
            const int idx = wrap_fast(c.i, a - 2, m.nx) + wrap_fast(c.j, b - 2, m.ny) * m.ld;
            rho_add(r, idx, cx * cy * weight);
        }
    }
}

// pl, ql of ua_steps.F90:60-66 with the reciprocals of l precomputed per lane
template <int N> DEVINL void pl_ql_fast(const FusedLane<N> &L, double t, double eps, cd elt, cd &pl, cd &ql) {
    if (L.T.k == 0) {
        pl = mk(t, 0.0);
        ql = mk(0.5 * t * t, 0.0);
    } else {
        const double il = L.inv_l(), il2 = L.inv_l2();
        pl = mk(-eps * elt.im * il, eps * (elt.re - 1.0) * il);
        ql = mk(eps * eps * (1.0 - elt.re) * il2, -eps * fma(eps, elt.im, L.T.lf * t) * il2);
    }
}

// fx = R(-tau) y / b in tau-Fourier space: F[cos(tau) y]_k = (Y_{k-1}+Y_{k+1})/2, F[sin(tau) y]_k = -i (Y_{k-1}-Y_{k+1})/2
// (ua_steps.F90:174-175 followed by the FFT of :187-188); y1h, y2h normalised coefficients in lane order
template <int N> DEVINL void fx_from_yhat(const FusedLane<N> &L, double rb, cd y1h, cd y2h, cd &fx1, cd &fx2) {
    const cd p1 = shfl_idx<N>(y1h, L.src_m1), m1 = shfl_idx<N>(y1h, L.src_p1);
    const cd p2 = shfl_idx<N>(y2h, L.src_m1), m2 = shfl_idx<N>(y2h, L.src_p1);
    const cd a1 = cadd(p1, m1), d1 = csub(p1, m1), a2 = cadd(p2, m2), d2 = csub(p2, m2);
    const double hb = 0.5 * rb;
    fx1 = mk(hb * (a1.re + d2.im), hb * (a1.im - d2.re));
    fx2 = mk(hb * (a2.re - d1.im), hb * (a2.im + d1.re));
}

// time-domain fy of ua_steps.F90:177-183
template <int N>
DEVINL void fy_time(const FusedLane<N> &L, double rb, double interv, cd yt1, cd yt2, double et1, double et2, cd &fy1, cd &fy2) {
    const double ct = L.T.ct, st = L.T.st;
    const cd t1 = mk(fma(ct * yt2.re - st * yt1.re, interv, et1), (ct * yt2.im - st * yt1.im) * interv);
    const cd t2 = mk(fma(-(ct * yt1.re + st * yt2.re), interv, et2), -(ct * yt1.im + st * yt2.im) * interv);
    fy1 = mk((ct * t1.re - st * t2.re) * rb, (ct * t1.im - st * t2.im) * rb);
    fy2 = mk((st * t1.re + ct * t2.re) * rb, (st * t1.im + ct * t2.im) * rb);
}

DEVINL double dot_conj(cd xh, cd elt) { return fma(xh.re, elt.re, xh.im * elt.im); }   // Re(xh * conj(elt))

struct FusedParams {
    PhaseParams p;
    MeshFast f;
    double inv_eps;
};

// =================================================================================================
// The predictor of one particle (N lanes): preparation -> [gather E] -> compute_f -> ua_step1 for x and y.
//   ua_steps.F90:49-113, interpolation_m6.F90:83-189, ua_steps.F90:160-195, :215-234
// GATHER = true : E at the tau samples is interpolated from the mesh and returned in et1, et2 (phase A)
// GATHER = false: et1, et2 are inputs (hybrid storage: phase B recomputes the predictor from x, v, e and the stored et)
// out: b, t; o4 = predicted xt1, xt2, yt1, yt2 in the time domain; fhat_x, fhat_y (normalised, lane order);
//      pos1, pos2 = predicted position at tau* = t/eps (compute_rho_m6.F90:74-87)
template <int N, bool GATHER>
DEVINL void predictor(const FusedLane<N> &L, const FusedParams &F, double x1, double x2, double vx, double vy, double ex, double ey,
                      double &et1, double &et2, double &b, double &t, cd (&o4)[4], cd &fx1, cd &fx2, cd &fy1, cd &fy2,
                      double &pos1, double &pos2) {
    const PhaseParams &P = F.p;
    const double eps = P.eps, inv_eps = F.inv_eps;
    const double ct = L.T.ct, st = L.T.st;
    const int k = L.T.k;

    // ---- preparation (ua_steps.F90:49-113) ----
    b = bfield(x1, x2);                                               // :54
    t = P.dt * b;                                                     // :55
    const double rb = 1.0 / b;
    const double vxb = vx * rb, vyb = vy * rb;                        // :73-74
    const double xt1 = x1 + eps * (st * vxb - ct * vyb) + eps * vyb;  // :78-81
    const double xt2 = x2 + eps * (st * vyb + ct * vxb) - eps * vxb;  // :79-82
    const double interv = (1.0 + 0.5 * sin(xt1) * sin(xt2) - b) * inv_eps;   // :87
    const double exb = ((ct * vy - st * vx) * interv + ex) * rb;      // :89
    const double eyb = ((-ct * vx - st * vy) * interv + ey) * rb;     // :90
    cd z[1] = {mk(ct * exb - st * eyb, st * exb + ct * eyb)};          // r1 + i r2   :92-93
    fft_fwd_b<N, 1>(z, L);                                            // :97-98 (both real signals at once)
    z[0] = rmul(L.s1, z[0]);
    const cd wn = shfl_idx<N>(z[0], L.src_neg);
    cd c[2];   // filtered coefficients, :100-103 ; slot 0 keeps the unnormalised sums
    if (k == 0) {
        c[0] = mk(z[0].re, 0.0); c[1] = mk(z[0].im, 0.0);
    } else {
        const double s = 0.5 * L.inv_lN();
        c[0] = mk(s * (z[0].im - wn.im), -s * (z[0].re + wn.re));
        c[1] = mk(-s * (z[0].re - wn.re), -s * (z[0].im + wn.im));
    }
    const double r10sum = group_bcast0<N>(c[0].re), r20sum = group_bcast0<N>(c[1].re);
    cd rp[2] = {rmul(L.s1, c[0]), rmul(L.s1, c[1])};
    fft_bwd_b<N, 2>(rp, L);                                           // :105-106
    const cd r10 = group_bcast0<N>(rp[0]), r20 = group_bcast0<N>(rp[1]);
    const cd yt1 = mk(vx + (rp[0].re - r10.re) * eps, (rp[0].im - r10.im) * eps);   // :109
    const cd yt2 = mk(vy + (rp[1].re - r20.re) * eps, (rp[1].im - r20.im) * eps);   // :110
    // FFT(yt)/N without a transform: eps*c_k for k != 0, mean value for k = 0
    cd yh1, yh2;
    if (k == 0) {
        yh1 = mk(vx - eps * (r10.re - r10sum), -eps * r10.im);
        yh2 = mk(vy - eps * (r20.re - r20sum), -eps * r20.im);
    } else {
        yh1 = rmul(eps, c[0]); yh2 = rmul(eps, c[1]);
    }

    // ---- gather E at the tau samples (interpolation_m6.F90:83-189) ----
    if (GATHER) {
        double xw, yw;
        const Cell cell = cell_fast(P.m, F.f, xt1, xt2, P.wrap, xw, yw);
        gather_fast(P.m, P.ehalo, cell, et1, et2);
    }

    // ---- compute_f (ua_steps.F90:160-195) ----
    fx_from_yhat<N>(L, rb, yh1, yh2, fx1, fx2);
    cd fy[2];
    fy_time<N>(L, rb, interv, yt1, yt2, et1, et2, fy[0], fy[1]);
    fft_fwd_b<N, 2>(fy, L);
    fy1 = rmul(L.sN, fy[0]); fy2 = rmul(L.sN, fy[1]);

    // ---- ua_step1 for x and y (ua_steps.F90:215-234) ----
    const cd elt = elt_minus<N>(L.T, t, eps);                         // :224
    cd pl, ql;
    pl_ql_fast<N>(L, t, eps, elt, pl, ql);
    // FFT(xt)/N of the first-order profile: modes 0 and +-1 only
    cd xh1 = mk(0.0, 0.0), xh2 = mk(0.0, 0.0);
    {
        const double he = 0.5 * eps;
        if (k == 0) { xh1 = mk(x1 + eps * vyb, 0.0); xh2 = mk(x2 - eps * vxb, 0.0); }
        if (k == (1 & (N - 1)) && N > 1 && k != 0) { xh1 = mk(xh1.re - he * vyb, xh1.im - he * vxb); xh2 = mk(xh2.re + he * vxb, xh2.im - he * vyb); }
        if (k == N - 1 && k != 0) { xh1 = mk(xh1.re - he * vyb, xh1.im + he * vxb); xh2 = mk(xh2.re + he * vxb, xh2.im + he * vyb); }
    }
    // everything below is carried NEGATED ON ODD LANES (input convention of fft_bwd_b); the position sums are
    // products of two such quantities, so they are unaffected
    const cd elts = rmul(L.s1, elt), pls = rmul(L.s1, pl);
    o4[0] = cfma(pls, fx1, cmul(elts, xh1));                          // :226
    o4[1] = cfma(pls, fx2, cmul(elts, xh2));                          // :227
    o4[2] = cfma(pls, fy1, cmul(elts, yh1));
    o4[3] = cfma(pls, fy2, cmul(elts, yh2));
    pos1 = group_sum<N>(dot_conj(o4[0], elts));
    pos2 = group_sum<N>(dot_conj(o4[1], elts));
    fft_bwd_b<N, 4>(o4, L);                                           // :231-232
}

// =================================================================================================
// HYB = false: store-full  -- 128 B per particle-tau cross the barrier (predicted xt, yt and fhat_x, fhat_y)
// HYB = true : hybrid      --  16 B per particle-tau (et only); phase B recomputes the predictor
template <int N, bool HYB>
__global__ void __launch_bounds__(kPhaseBlock, kPhaseMinBlocks) k_phase_a(FusedParams F) {
    const PhaseParams &P = F.p;
    __shared__ double2 sm_tw[4 * 32];
    __shared__ double sm_inv[3 * 32];
    FusedLane<N> L; L.init(threadIdx.x & 31, sm_tw, sm_inv);
    WarpMap<N> W;
    for (int64_t base = W.first; base < P.np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < P.np;
        const int64_t ip = valid ? kraw : P.np - 1;
        const double2 xx = P.x[ip], vv = P.v[ip], ee = P.ep[ip];
        double et1, et2, b, t, pos1, pos2;
        cd o4[4], fx1, fx2, fy1, fy2;
        predictor<N, true>(L, F, xx.x, xx.y, vv.x, vv.y, ee.x, ee.y, et1, et2, b, t, o4, fx1, fx2, fy1, fy2, pos1, pos2);

        if (valid) {
            if (HYB) {
                double *e = P.etstore + (size_t)ip * 2 * N + L.T.j;
                e[0] = et1; e[N] = et2;
            } else {
                double2 *s = P.store + (size_t)ip * 8 * N + L.T.j;
                s[0 * N] = make_double2(o4[0].re, o4[0].im);
                s[1 * N] = make_double2(o4[1].re, o4[1].im);
                s[2 * N] = make_double2(o4[2].re, o4[2].im);
                s[3 * N] = make_double2(o4[3].re, o4[3].im);
                s[4 * N] = make_double2(fx1.re, fx1.im);
                s[5 * N] = make_double2(fx2.re, fx2.im);
                s[6 * N] = make_double2(fy1.re, fy1.im);
                s[7 * N] = make_double2(fy2.re, fy2.im);
                if (L.T.j == 0) P.tb[ip] = make_double2(t, b);
            }
        }
        // ---- predictor deposit (compute_rho_m6.F90:89-187) ----
        double xw, yw;
        const Cell cell = cell_fast(P.m, F.f, pos1, pos2, P.wrap, xw, yw);
        deposit_coop<N>(P.m, P.rho, cell, P.weight, L.T.j, valid);
    }
}

// =================================================================================================
template <int N, bool HYB>
__global__ void __launch_bounds__(kPhaseBlock, kPhaseMinBlocks) k_phase_b(FusedParams F) {
    const PhaseParams &P = F.p;
    __shared__ double2 sm_tw[4 * 32];
    __shared__ double sm_inv[3 * 32];
    FusedLane<N> L; L.init(threadIdx.x & 31, sm_tw, sm_inv);
    WarpMap<N> W;
    const double eps = P.eps, inv_eps = F.inv_eps;
    for (int64_t base = W.first; base < P.np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < P.np;
        const int64_t ip = valid ? kraw : P.np - 1;
        double t, b;
        cd f6[6], fx1, fx2, fy1, fy2;
        if (HYB) {
            const double2 xx = P.x[ip], vv = P.v[ip], ee = P.ep[ip];
            const double *e = P.etstore + (size_t)ip * 2 * N + L.T.j;
            double et1 = e[0], et2 = e[N], p1, p2;
            cd o4[4];
            predictor<N, false>(L, F, xx.x, xx.y, vv.x, vv.y, ee.x, ee.y, et1, et2, b, t, o4, fx1, fx2, fy1, fy2, p1, p2);
#pragma unroll
            for (int q = 0; q < 4; ++q) f6[q] = o4[q];
        } else {
            const double2 *s = P.store + (size_t)ip * 8 * N + L.T.j;
            const double2 tb = P.tb[ip];
            t = tb.x; b = tb.y;
#pragma unroll
            for (int q = 0; q < 4; ++q) f6[q] = mk(s[q * N].x, s[q * N].y);          // predicted xt1, xt2, yt1, yt2
            fx1 = mk(s[4 * N].x, s[4 * N].y); fx2 = mk(s[5 * N].x, s[5 * N].y);
            fy1 = mk(s[6 * N].x, s[6 * N].y); fy2 = mk(s[7 * N].x, s[7 * N].y);
        }
        const double rb = 1.0 / b;

        // ---- gather E_new at the predicted samples, g in the time domain (ua_steps.F90:160-185) ----
        double et1, et2;
        {
            double xw, yw;
            const Cell cell = cell_fast(P.m, F.f, f6[0].re, f6[1].re, P.wrap, xw, yw);
            gather_fast(P.m, P.ehalo, cell, et1, et2);
        }
        const double interv = (1.0 + 0.5 * sin(f6[0].re) * sin(f6[1].re) - b) * inv_eps;   // :177
        fy_time<N>(L, rb, interv, f6[2], f6[3], et1, et2, f6[4], f6[5]);

        // six forward transforms at once: predicted x, y (-> elt/N*xf + pl*fhat) and gy   (:187-195)
        fft_fwd_b<N, 6>(f6, L);
#pragma unroll
        for (int q = 0; q < 6; ++q) f6[q] = rmul(L.sN, f6[q]);
        cd gx1, gx2;
        fx_from_yhat<N>(L, rb, f6[2], f6[3], gx1, gx2);

        // ---- ua_step2 (ua_steps.F90:252-265): + ql*(ghat - fhat)/t ----
        const cd elt = elt_minus<N>(L.T, t, eps);                         // :258
        cd pl, ql;
        pl_ql_fast<N>(L, t, eps, elt, pl, ql);
        const cd qt = rmul(1.0 / t, ql);
        const cd xc1 = cfma(qt, csub(gx1, fx1), f6[0]), xc2 = cfma(qt, csub(gx2, fx2), f6[1]);
        const cd yc1 = cfma(qt, csub(f6[4], fy1), f6[2]), yc2 = cfma(qt, csub(f6[5], fy2), f6[3]);

        // ---- corrector deposit position and compute_v (compute_rho_m6.F90:74-87, ua_steps.F90:293-303) ----
        const double pos1 = group_sum<N>(dot_conj(xc1, elt));
        const double pos2 = group_sum<N>(dot_conj(xc2, elt));
        const double px = group_sum<N>(dot_conj(yc1, elt));
        const double py = group_sum<N>(dot_conj(yc2, elt));
        // cos(t/eps), sin(t/eps) sit in the lane whose slot is |l| = 1: elt there is exp(-+ i t/eps)
        const cd e1 = shfl_idx<N>(elt, FusedLane<N>::lane_of(1));
        const double cs = e1.re, sn = (N > 2) ? -e1.im : e1.im;

        double xw, yw;
        const Cell cell = cell_fast(P.m, F.f, pos1, pos2, P.wrap, xw, yw);
        if (valid && L.T.j == 0) {
            P.x[ip] = make_double2(xw, yw);                                           // compute_rho_m6.F90:86-87
            P.v[ip] = make_double2(cs * px + sn * py, cs * py - sn * px);             // ua_steps.F90:302-303
        }
        deposit_coop<N>(P.m, P.rho, cell, P.weight, L.T.j, valid);
    }
}

inline int phase_grid(const LaunchCtx &c, int64_t np, int N) {
    const int per_block = (kPhaseBlock / 32) * (32 / N);
    int64_t need = (np + per_block - 1) / per_block;
    const int64_t cap = (int64_t)c.sm_count * kPhaseMinBlocks;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

FusedParams make_fused(const PhaseParams &p) {
    FusedParams F;
    F.p = p;
    F.f.inv_dx = 1.0 / p.m.dx; F.f.inv_dy = 1.0 / p.m.dy;
    F.f.inv_nx = 1.0 / (double)p.m.nx; F.f.inv_ny = 1.0 / (double)p.m.ny;
    F.f.inv_dimx = 1.0 / p.m.dimx; F.f.inv_dimy = 1.0 / p.m.dimy;
    F.inv_eps = 1.0 / p.eps;
    return F;
}

}  // namespace

#define UAPIC_DISPATCH_N(ntau, CALL)                      \
    switch (ntau) {                                       \
        case 2:  { constexpr int N = 2;  CALL; } break;   \
        case 4:  { constexpr int N = 4;  CALL; } break;   \
        case 8:  { constexpr int N = 8;  CALL; } break;   \
        case 16: { constexpr int N = 16; CALL; } break;   \
        case 32: { constexpr int N = 32; CALL; } break;   \
        default: return cudaErrorInvalidValue;            \
    }

cudaError_t launch_phase_a(const LaunchCtx &c, const PhaseParams &p) {
    if (p.np <= 0) return cudaSuccess;
    if (p.m.nx < 4 || p.m.ny < 4) return cudaErrorInvalidValue;
    const FusedParams F = make_fused(p);
    if (p.hybrid) { UAPIC_DISPATCH_N(p.ntau, (k_phase_a<N, true><<<phase_grid(c, p.np, N), kPhaseBlock, 0, c.stream>>>(F))); }
    else { UAPIC_DISPATCH_N(p.ntau, (k_phase_a<N, false><<<phase_grid(c, p.np, N), kPhaseBlock, 0, c.stream>>>(F))); }
    if (c.launches) *c.launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_phase_b(const LaunchCtx &c, const PhaseParams &p) {
    if (p.np <= 0) return cudaSuccess;
    if (p.m.nx < 4 || p.m.ny < 4) return cudaErrorInvalidValue;
    const FusedParams F = make_fused(p);
    if (p.hybrid) { UAPIC_DISPATCH_N(p.ntau, (k_phase_b<N, true><<<phase_grid(c, p.np, N), kPhaseBlock, 0, c.stream>>>(F))); }
    else { UAPIC_DISPATCH_N(p.ntau, (k_phase_b<N, false><<<phase_grid(c, p.np, N), kPhaseBlock, 0, c.stream>>>(F))); }
    if (c.launches) *c.launches += 1;
    return cudaGetLastError();
}

}  // namespace uapic

// uapic_generic.cu -- the tau-stage kernels for ANY even ntau <= 256 (sm_100a).
//
// The reference takes any even ntau (FFTW plans of length ntau, ua_type.F90:32-76 / src/ua_type.jl:17-41).  The fast kernels of
// this library map one tau sample to one lane and need ntau to be a power of two <= 32; this file is the general path behind
// the same entry points: ONE WARP PER PARTICLE, samples and modes strided over the lanes (n = lane + 32 r), the length-ntau
// transforms done as direct DFTs out of shared memory (O(ntau^2 / 32) per lane; twiddle table exp(-2 pi i m / ntau) per CTA).
// It is a completeness path, not a fast one: the BASELINE configurations (ntau = 16, 32) never come here.
// Arithmetic follows the same reference lines as the templated kernels in uapic_kernels.cu.
#include "uapic_internal.h"

namespace uapic {

namespace {

constexpr int kGenBlock = 128;                 // 4 warps = 4 particles per CTA sweep
constexpr int kGenMaxR = kGenericMaxNtau / 32; // registers per lane and array

struct Gen {
    int N, R, lane;
    cd *buf;            // per-warp exchange area, N entries
    const cd *tw;       // exp(-2 pi i m / N), m < N
    DEVINL double lf(int k) const { return (double)(k < N / 2 ? k : k - N); }      // ua_type.F90:51-56
};

DEVINL double warp_sum(double v) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) v += __shfl_xor_sync(kFull, v, h);
    return v;
}

// a[r] <-> index lane + 32 r.  sign -1: forward (FFTW_FORWARD), +1: backward; unnormalised.
DEVINL void gen_dft(const Gen &g, cd (&a)[kGenMaxR], int sign) {
    __syncwarp();
    for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; if (n < g.N) g.buf[n] = a[r]; }
    __syncwarp();
    for (int r = 0; r < g.R; ++r) {
        const int k = g.lane + 32 * r;
        if (k >= g.N) { a[r] = mk(0.0, 0.0); continue; }
        cd acc = mk(0.0, 0.0);
        int idx = 0;
        for (int n = 0; n < g.N; ++n) {
            cd w = g.tw[idx];
            if (sign > 0) w.im = -w.im;
            acc = cfma(g.buf[n], w, acc);
            idx += k; if (idx >= g.N) idx -= g.N;
        }
        a[r] = acc;
    }
    __syncwarp();
}

DEVINL void gen_setup(Gen &g, int ntau, cd *smem) {
    g.N = ntau; g.R = (ntau + 31) / 32; g.lane = threadIdx.x & 31;
    cd *tw = smem;
    for (int m = threadIdx.x; m < ntau; m += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)m / (double)ntau, &s, &c);
        tw[m] = mk(c, s);
    }
    g.tw = tw;
    g.buf = smem + ntau + (threadIdx.x >> 5) * ntau;
    __syncthreads();
}

DEVINL cd gen_elt(const Gen &g, int k, double t, double eps) {       // exp(-i l t/eps), phase formed as -(l*t)/eps   ua_steps.F90:64
    double s, c;
    sincos(-(g.lf(k) * t) / eps, &s, &c);
    return mk(c, s);
}

DEVINL void gen_pl_ql(const Gen &g, int k, double t, double eps, cd elt, cd &pl, cd &ql) {      // ua_steps.F90:60-66
    if (k == 0) { pl = mk(t, 0.0); ql = mk(t * t / 2.0, 0.0); return; }
    const double l = g.lf(k), l2 = l * l;
    pl = mk((-eps * elt.im) / l, (eps * (elt.re - 1.0)) / l);
    ql = mk((eps * (eps * (1.0 - elt.re))) / l2, (eps * (-eps * elt.im - l * t)) / l2);
}

#define GEN_PARTICLE_LOOP(np)                                                                                          \
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < ((np) + 3) / 4 * 4;             \
         k += (int64_t)gridDim.x * (kGenBlock / 32))

// preparation                                         ua_steps.F90:15-115
__global__ void __launch_bounds__(kGenBlock) k_gen_preparation(int ntau, double eps, double dt, int64_t np, const double *x, const double *v,
                                                               const double *e, double *b, double *t, double *pl, double *ql, double *xt,
                                                               double *yt) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < np; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        const double2 xx = ld2(x, k), vv = ld2(v, k), ee = ld2(e, k);
        const double bb = bfield(xx.x, xx.y), tt = dt * bb;                       // :54-55
        const double vxb = vv.x / bb, vyb = vv.y / bb;
        cd r1[kGenMaxR], r2[kGenMaxR];
        double x1s[kGenMaxR], x2s[kGenMaxR];
        for (int r = 0; r < g.R; ++r) {
            const int n = g.lane + 32 * r;
            r1[r] = r2[r] = mk(0.0, 0.0); x1s[r] = x2s[r] = 0.0;
            if (n >= N) continue;
            double st, ct;
            sincospi(2.0 * (double)n / (double)N, &st, &ct);
            const double xt1 = xx.x + eps * (st * vxb - ct * vyb) + eps * vyb;    // :78-82
            const double xt2 = xx.y + eps * (st * vyb + ct * vxb) - eps * vxb;
            const double interv = (1.0 + 0.5 * sin(xt1) * sin(xt2) - bb) / eps;   // :87
            const double exb = ((ct * vv.y - st * vv.x) * interv + ee.x) / bb;    // :89-90
            const double eyb = ((-ct * vv.x - st * vv.y) * interv + ee.y) / bb;
            r1[r] = mk(ct * exb - st * eyb, 0.0);                                 // :92-93
            r2[r] = mk(st * exb + ct * eyb, 0.0);
            x1s[r] = xt1; x2s[r] = xt2;
        }
        gen_dft(g, r1, -1); gen_dft(g, r2, -1);                                   // :97-98
        for (int r = 0; r < g.R; ++r) {
            const int kk = g.lane + 32 * r;
            if (kk == 0 || kk >= N) continue;
            const double s = 1.0 / (g.lf(kk) * (double)N);                        // :100-103  rf = -(i/l) rf / N
            r1[r] = mk(r1[r].im * s, -r1[r].re * s);
            r2[r] = mk(r2[r].im * s, -r2[r].re * s);
        }
        gen_dft(g, r1, +1); gen_dft(g, r2, +1);                                   // :105-106
        const cd r10 = mk(__shfl_sync(kFull, r1[0].re, 0), __shfl_sync(kFull, r1[0].im, 0));
        const cd r20 = mk(__shfl_sync(kFull, r2[0].re, 0), __shfl_sync(kFull, r2[0].im, 0));
        if (g.lane == 0) { b[k] = bb; t[k] = tt; }
        for (int r = 0; r < g.R; ++r) {
            const int n = g.lane + 32 * r;
            if (n >= N) continue;
            const cd elt = gen_elt(g, n, tt, eps);
            cd plv, qlv;
            gen_pl_ql(g, n, tt, eps, elt, plv, qlv);
            st2(pl, n + (int64_t)N * k, plv);
            st2(ql, n + (int64_t)N * k, qlv);
            st2(xt, n + (int64_t)N * (0 + 2 * k), mk(x1s[r], 0.0));
            st2(xt, n + (int64_t)N * (1 + 2 * k), mk(x2s[r], 0.0));
            st2(yt, n + (int64_t)N * (0 + 2 * k), mk(vv.x + (r1[r].re - r10.re) * eps, (r1[r].im - r10.im) * eps));   // :109
            st2(yt, n + (int64_t)N * (1 + 2 * k), mk(vv.y + (r2[r].re - r20.re) * eps, (r2[r].im - r20.im) * eps));   // :110
        }
    }
}

// compute_f                                           ua_steps.F90:140-198
__global__ void __launch_bounds__(kGenBlock) k_gen_compute_f(int ntau, double eps, int64_t np, const double *b, const double *xt,
                                                             const double *yt, const double *et, double *fx, double *fy, int normalise) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    const double sc = normalise ? 1.0 / (double)N : 1.0;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < np; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        const double bb = b[k], rb = 1.0 / bb;
        cd f1[kGenMaxR], f2[kGenMaxR], g1[kGenMaxR], g2[kGenMaxR];
        for (int r = 0; r < g.R; ++r) {
            const int n = g.lane + 32 * r;
            f1[r] = f2[r] = g1[r] = g2[r] = mk(0.0, 0.0);
            if (n >= N) continue;
            const int64_t i1 = n + (int64_t)N * (0 + 2 * k), i2 = n + (int64_t)N * (1 + 2 * k);
            double st, ct;
            sincospi(2.0 * (double)n / (double)N, &st, &ct);
            const double x1 = xt[2 * i1], x2 = xt[2 * i2];
            const cd y1 = ldc(yt, i1), y2 = ldc(yt, i2);
            const double interv = (1.0 + 0.5 * sin(x1) * sin(x2) - bb) / eps;                                  // :177
            f1[r] = mk((ct * y1.re + st * y2.re) * rb, (ct * y1.im + st * y2.im) * rb);                        // :174
            f2[r] = mk((ct * y2.re - st * y1.re) * rb, (ct * y2.im - st * y1.im) * rb);                        // :175
            const cd t1 = mk(et[i1] + (ct * y2.re - st * y1.re) * interv, (ct * y2.im - st * y1.im) * interv);   // :179
            const cd t2 = mk(et[i2] - (ct * y1.re + st * y2.re) * interv, -(ct * y1.im + st * y2.im) * interv);  // :180
            g1[r] = mk((ct * t1.re - st * t2.re) * rb, (ct * t1.im - st * t2.im) * rb);                        // :182
            g2[r] = mk((st * t1.re + ct * t2.re) * rb, (st * t1.im + ct * t2.im) * rb);                        // :183
        }
        gen_dft(g, f1, -1); gen_dft(g, f2, -1); gen_dft(g, g1, -1); gen_dft(g, g2, -1);                        // :187-190
        for (int r = 0; r < g.R; ++r) {
            const int kk = g.lane + 32 * r;
            if (kk >= N) continue;
            const int64_t o1 = kk + (int64_t)N * (0 + 2 * k), o2 = kk + (int64_t)N * (1 + 2 * k);
            st2(fx, o1, rmul(sc, f1[r])); st2(fx, o2, rmul(sc, f2[r]));                                        // :194-195
            st2(fy, o1, rmul(sc, g1[r])); st2(fy, o2, rmul(sc, g2[r]));
        }
    }
}

// mul!(x̃t, ftau, xt) / ifft!(xt,1)                    test/bupdate.jl:79,85
__global__ void __launch_bounds__(kGenBlock) k_gen_fft_tau(int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    const double sc = normalise ? 1.0 / (double)N : 1.0;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < nvec; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        cd a[kGenMaxR];
        for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; a[r] = n < N ? ldc(in, n + (int64_t)N * k) : mk(0.0, 0.0); }
        gen_dft(g, a, sign < 0 ? -1 : +1);
        for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; if (n < N) st2(out, n + (int64_t)N * k, rmul(sc, a[r])); }
    }
}

// ua_step1 / ua_step2 (Fortran forms)                  ua_steps.F90:200-272
__global__ void __launch_bounds__(kGenBlock) k_gen_step_fortran(int ntau, double eps, int64_t np, const double *t, const double *pl,
                                                                const double *ql, double *xt, double *xf, const double *fx,
                                                                const double *gx, int corrector) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < np; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        const double tt = t[k];
        for (int c = 0; c < 2; ++c) {
            cd a[kGenMaxR];
            if (!corrector) {
                for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; a[r] = n < N ? ldc(xt, n + (int64_t)N * (c + 2 * k)) : mk(0.0, 0.0); }
                gen_dft(g, a, -1);                                                                             // :217-218
            }
            for (int r = 0; r < g.R; ++r) {
                const int kk = g.lane + 32 * r;
                if (kk >= N) { a[r] = mk(0.0, 0.0); continue; }
                const int64_t is = kk + (int64_t)N * (c + 2 * k);
                cd xfv;
                if (!corrector) { xfv = a[r]; st2(xf, is, xfv); } else xfv = ldc(xf, is);
                const cd elt = rmul(1.0 / (double)N, gen_elt(g, kk, tt, eps));                                 // :224-225, :258-259
                const cd plv = ldc(pl, kk + (int64_t)N * k), f = ldc(fx, is);
                cd rr = cfma(plv, f, cmul(elt, xfv));                                                          // :226, :260
                if (corrector) {
                    const cd q = cmul(ldc(ql, kk + (int64_t)N * k), csub(ldc(gx, is), f));
                    rr = mk(rr.re + q.re / tt, rr.im + q.im / tt);                                             // :261
                }
                a[r] = rr;
            }
            gen_dft(g, a, +1);                                                                                 // :231-232, :267-268
            for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; if (n < N) st2(xt, n + (int64_t)N * (c + 2 * k), a[r]); }
        }
    }
}

// tau* evaluation  sum_k xhat_k exp(+i l_k t/eps) / N  of a time-domain profile   compute_rho_m6.F90:74-84, ua_steps.F90:293-300
DEVINL double gen_tau_star(const Gen &g, cd (&a)[kGenMaxR], double tt, double eps, bool already_fourier) {
    if (!already_fourier) gen_dft(g, a, -1);
    double s = 0.0;
    for (int r = 0; r < g.R; ++r) {
        const int kk = g.lane + 32 * r;
        if (kk >= g.N) continue;
        const cd elt = gen_elt(g, kk, tt, eps);
        const cd f = rmul(1.0 / (double)g.N, a[r]);
        s += f.re * elt.re + f.im * elt.im;                        // Re(f * conj(elt))
    }
    return warp_sum(s);
}

// compute_rho_m6_complex (per-particle part)           compute_rho_m6.F90:70-187
__global__ void __launch_bounds__(kGenBlock) k_gen_deposit_tau(MeshDev m, int ntau, double eps, int64_t np, const double *xt, const double *t,
                                                               double w, RhoAcc acc, double *x, int wrap) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < np; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        const double tt = t[k];
        double pos[2];
        for (int c = 0; c < 2; ++c) {
            cd a[kGenMaxR];
            for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; a[r] = n < N ? ldc(xt, n + (int64_t)N * (c + 2 * k)) : mk(0.0, 0.0); }
            pos[c] = gen_tau_star(g, a, tt, eps, false);
        }
        double xw, yw;
        const Cell c = m6_cell_exact(m, pos[0], pos[1], wrap, xw, yw);
        if (g.lane == 0) reinterpret_cast<double2 *>(x)[k] = make_double2(xw, yw);                             // :86-87
        m6_scatter(m, acc, c, w, g.lane, 32);
    }
}

// compute_v                                            ua_steps.F90:274-307 / src/ua_steps.jl:204-224
__global__ void __launch_bounds__(kGenBlock) k_gen_compute_v(int ntau, double eps, int64_t np, const double *t, const double *yt,
                                                             int yt_is_fourier, double *v) {
    extern __shared__ double2 smem_raw[];
    Gen g; gen_setup(g, ntau, reinterpret_cast<cd *>(smem_raw));
    const int N = ntau;
    for (int64_t k = (int64_t)blockIdx.x * (kGenBlock / 32) + (threadIdx.x >> 5); k < np; k += (int64_t)gridDim.x * (kGenBlock / 32)) {
        const double tt = t[k];
        double p[2];
        for (int c = 0; c < 2; ++c) {
            cd a[kGenMaxR];
            for (int r = 0; r < g.R; ++r) { const int n = g.lane + 32 * r; a[r] = n < N ? ldc(yt, n + (int64_t)N * (c + 2 * k)) : mk(0.0, 0.0); }
            p[c] = gen_tau_star(g, a, tt, eps, yt_is_fourier != 0);
        }
        double sn, cs;
        sincos(tt / eps, &sn, &cs);
        if (g.lane == 0) reinterpret_cast<double2 *>(v)[k] = make_double2(cs * p[0] + sn * p[1], cs * p[1] - sn * p[0]);   // :302-303
    }
}

inline int gen_grid(const LaunchCtx &c, int64_t np) {
    int64_t need = (np + kGenBlock / 32 - 1) / (kGenBlock / 32);
    const int64_t cap = (int64_t)c.sm_count * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
inline size_t gen_smem(int ntau) { return sizeof(double2) * (size_t)ntau * (1 + kGenBlock / 32); }
inline void gcount(const LaunchCtx &c) { if (c.launches) *c.launches += 1; }

}  // namespace

bool generic_ntau_supported(int ntau) { return ntau >= 2 && ntau <= kGenericMaxNtau && (ntau % 2) == 0; }

cudaError_t launch_preparation_generic(const LaunchCtx &c, int ntau, double eps, double dt, int64_t np, const double *x, const double *v,
                                       const double *e, double *b, double *t, double *pl, double *ql, double *xt, double *yt) {
    k_gen_preparation<<<gen_grid(c, np), kGenBlock, gen_smem(ntau), c.stream>>>(ntau, eps, dt, np, x, v, e, b, t, pl, ql, xt, yt);
    gcount(c);
    return cudaGetLastError();
}
cudaError_t launch_compute_f_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *b, const double *xt, const double *yt,
                                     const double *et, double *fx, double *fy, int normalise) {
    k_gen_compute_f<<<gen_grid(c, np), kGenBlock, gen_smem(ntau), c.stream>>>(ntau, eps, np, b, xt, yt, et, fx, fy, normalise);
    gcount(c);
    return cudaGetLastError();
}
cudaError_t launch_fft_tau_generic(const LaunchCtx &c, int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise) {
    k_gen_fft_tau<<<gen_grid(c, nvec), kGenBlock, gen_smem(ntau), c.stream>>>(ntau, nvec, in, out, sign, normalise);
    gcount(c);
    return cudaGetLastError();
}
cudaError_t launch_step_fortran_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *pl,
                                        const double *ql, double *xt, double *xf, const double *fx, const double *gx, int corrector) {
    k_gen_step_fortran<<<gen_grid(c, np), kGenBlock, gen_smem(ntau), c.stream>>>(ntau, eps, np, t, pl, ql, xt, xf, fx, gx, corrector);
    gcount(c);
    return cudaGetLastError();
}
cudaError_t launch_deposit_tau_generic(const LaunchCtx &c, const MeshDev &m, int ntau, double eps, int64_t np, const double *xt,
                                       const double *t, double w, const RhoAcc &acc, double *x, int wrap) {
    k_gen_deposit_tau<<<gen_grid(c, np), kGenBlock, gen_smem(ntau), c.stream>>>(m, ntau, eps, np, xt, t, w, acc, x, wrap);
    gcount(c);
    return cudaGetLastError();
}
cudaError_t launch_compute_v_generic(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *yt,
                                     int yt_is_fourier, double *v) {
    k_gen_compute_v<<<gen_grid(c, np), kGenBlock, gen_smem(ntau), c.stream>>>(ntau, eps, np, t, yt, yt_is_fourier, v);
    gcount(c);
    return cudaGetLastError();
}

}  // namespace uapic

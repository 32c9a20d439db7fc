// uapic_device.cuh -- device-side building blocks shared by the stage kernels and the fused
// phase kernels.  sm_100a only.
//
// Work mapping: one tau sample per lane.  A warp holds G = 32/N particles
// (N = ntau, a power of two <= 32); lane j of an N-lane group owns tau_j in the time domain and
// Fourier slot k = bitrev(j) in the tau-Fourier domain.  Length-N FFTs run across lanes with
// __shfl_xor_sync (DIF forward: natural -> bit-reversed; DIT backward: bit-reversed -> natural), so
// no permutation is ever applied on chip; natural-order Fourier arrays exist only at the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define DEVINL __device__ __forceinline__

namespace uapic {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWrapFortran = 0;
constexpr int kWrapJulia = 1;

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
struct cd { double re, im; };

DEVINL cd mk(double r, double i) { cd z; z.re = r; z.im = i; return z; }
DEVINL cd cadd(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
DEVINL cd csub(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
DEVINL cd cmul(cd a, cd b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
DEVINL cd cmulc(cd a, cd b) { return mk(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im); }  // a * conj(b)
DEVINL cd rmul(double r, cd a) { return mk(r * a.re, r * a.im); }
DEVINL cd cfma(cd a, cd b, cd c) {  // a*b + c
    return mk(fma(a.re, b.re, fma(-a.im, b.im, c.re)), fma(a.re, b.im, fma(a.im, b.re, c.im)));
}

DEVINL cd shfl_xor(cd v, int m) { return mk(__shfl_xor_sync(kFull, v.re, m), __shfl_xor_sync(kFull, v.im, m)); }

template <int N> DEVINL double group_sum(double v) {
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) v += __shfl_xor_sync(kFull, v, h);
    return v;
}
template <int N> DEVINL double group_bcast0(double v) { return __shfl_sync(kFull, v, 0, N); }
template <int N> DEVINL cd group_bcast0(cd v) { return mk(group_bcast0<N>(v.re), group_bcast0<N>(v.im)); }

// ---------------------------------------------------------------------------------------------
// mesh
// ---------------------------------------------------------------------------------------------
struct MeshDev {
    double xmin, ymin, dimx, dimy, dx, dy;
    int nx, ny, ld;  // ld = nx + 1 : meshes keep the reference's ghost row/column
};

// ---------------------------------------------------------------------------------------------
// per-lane constants
// ---------------------------------------------------------------------------------------------
template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

template <int N> struct TauLane {
    static constexpr int LOG = Log2<N>::v;
    int j;        // tau index of this lane inside its group
    int k;        // Fourier slot held by this lane (bit reversal of j)
    double ct, st;  // cos(tau_j), sin(tau_j)                       ua_type.F90:60-62
    double lf;    // ltau[k]                                        ua_type.F90:51-56
    double twr[LOG > 1 ? LOG - 1 : 1], twi[LOG > 1 ? LOG - 1 : 1];  // DIF twiddles, stage s <-> h = N >> (s+1)

    DEVINL void init(int lane) {
        j = lane & (N - 1);
        k = (int)(__brev((unsigned)j) >> (32 - LOG));
        sincospi(2.0 * (double)j / (double)N, &st, &ct);
        lf = (k < N / 2) ? (double)k : (double)(k - N);
#pragma unroll
        for (int s = 0; s < LOG - 1; ++s) {
            const int h = N >> (s + 1);
            if (j & h) {
                double sn, cs;
                sincospi(-(double)(j & (h - 1)) / (double)h, &sn, &cs);  // exp(-2 pi i (j mod h) / (2h))
                twr[s] = cs; twi[s] = sn;
            } else {
                twr[s] = 1.0; twi[s] = 0.0;
            }
        }
    }
};

// forward, unnormalised (FFTW_FORWARD): natural order in, lane j ends with X[bitrev(j)]
template <int N> DEVINL cd fft_fwd(cd v, const TauLane<N> &L) {
    constexpr int LOG = TauLane<N>::LOG;
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int h = N >> (s + 1);
        const cd o = shfl_xor(v, h);
        const double sg = (L.j & h) ? -1.0 : 1.0;
        cd d = mk(fma(sg, v.re, o.re), fma(sg, v.im, o.im));   // upper: v+o ; lower: o-v
        if (h > 1) d = cmul(d, mk(L.twr[s], L.twi[s]));
        v = d;
    }
    return v;
}

// backward, unnormalised (FFTW_BACKWARD): lane j holds X[bitrev(j)] in, natural order out
template <int N> DEVINL cd fft_bwd(cd v, const TauLane<N> &L) {
    constexpr int LOG = TauLane<N>::LOG;
#pragma unroll
    for (int s = LOG - 1; s >= 0; --s) {
        const int h = N >> (s + 1);
        if (h > 1) v = cmulc(v, mk(L.twr[s], L.twi[s]));
        const cd o = shfl_xor(v, h);
        const double sg = (L.j & h) ? -1.0 : 1.0;
        v = mk(fma(sg, v.re, o.re), fma(sg, v.im, o.im));
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// M6 (quintic spline)          compute_rho_m6.F90:28-45, src/compute_rho.jl:10-25
// ---------------------------------------------------------------------------------------------

// reference operation order, never contracted to FMA: bit-identical to the CPU oracle
DEVINL double pow5_rn(double x) {
    const double x2 = __dmul_rn(x, x);
    const double x4 = __dmul_rn(x2, x2);
    return __dmul_rn(x4, x);
}
DEVINL double f_m6_exact(double q) {
    double f;
    if (q < 1.0)
        f = __dadd_rn(__dsub_rn(pow5_rn(__dsub_rn(3.0, q)), __dmul_rn(6.0, pow5_rn(__dsub_rn(2.0, q)))),
                      __dmul_rn(15.0, pow5_rn(__dsub_rn(1.0, q))));
    else if (q < 2.0)
        f = __dsub_rn(pow5_rn(__dsub_rn(3.0, q)), __dmul_rn(6.0, pow5_rn(__dsub_rn(2.0, q))));
    else if (q < 3.0)
        f = pow5_rn(__dsub_rn(3.0, q));
    else
        f = 0.0;
    return __ddiv_rn(f, 120.0);
}
// weight of the node at offset `off` (-2..3) from cell index i, for in-cell position dp in [0,1)
// arguments exactly as compute_rho_m6.F90:118-131 forms them: |off|+dp for off<=0, off-dp for off>0
DEVINL double m6_weight_exact(int off, double dp) {
    const double q = (off <= 0) ? __dadd_rn((double)(-off), dp) : __dsub_rn((double)off, dp);
    return f_m6_exact(q);
}

DEVINL double pow5(double x) { const double x2 = x * x; return x2 * x2 * x; }
// the six non-zero weights for offsets -2..3 (offset -3 is identically zero for dp in [0,1))
DEVINL void m6_weights_fast(double dp, double w[6]) {
    const double a = pow5(dp), b = pow5(1.0 + dp), c = pow5(2.0 + dp);
    const double d = pow5(1.0 - dp), e = pow5(2.0 - dp), f = pow5(3.0 - dp);
    const double k = 1.0 / 120.0;
    w[0] = d * k;
    w[1] = fma(-6.0, d, e) * k;
    w[2] = fma(15.0, d, fma(-6.0, e, f)) * k;
    w[3] = fma(15.0, a, fma(-6.0, b, c)) * k;
    w[4] = fma(-6.0, a, b) * k;
    w[5] = a * k;
}

// Fortran MODULO / Julia mod, exactly as the oracle (fmod is exact)
DEVINL double modulo_exact(double a, double p) {
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r = __dadd_rn(r, p);
    return r;
}
// same value without the slow fmod: q may be off by one, the two corrections repair it; the only
// difference to modulo_exact is a possible double rounding when a/p is within 1 ulp of an integer
DEVINL double modulo_fast(double a, double p, double inv_p) {
    const double q = floor(a * inv_p);
    double r = fma(-q, p, a);
    if (r >= p) r -= p;
    if (r < 0.0) r += p;
    return r;
}

struct Cell { int i, j; double dpx, dpy; };

// compute_rho_m6.F90:89-98 (Fortran) / src/compute_rho.jl:63-76 (Julia); xw,yw = value stored back into particles.x
DEVINL Cell m6_cell_exact(const MeshDev &m, double x, double y, int wrap, double &xw, double &yw) {
    double px, py;
    if (wrap == kWrapJulia) {
        const double xn = modulo_exact(__dsub_rn(x, m.xmin), m.dimx);
        const double yn = modulo_exact(__dsub_rn(y, m.ymin), m.dimy);
        px = __ddiv_rn(xn, m.dx); py = __ddiv_rn(yn, m.dy);
        xw = __dadd_rn(xn, m.xmin); yw = __dadd_rn(yn, m.ymin);
    } else {
        px = modulo_exact(__ddiv_rn(x, m.dx), (double)m.nx);
        py = modulo_exact(__ddiv_rn(y, m.dy), (double)m.ny);
        xw = x; yw = y;
    }
    Cell c;
    c.i = (int)floor(px); c.dpx = __dsub_rn(px, (double)c.i);
    c.j = (int)floor(py); c.dpy = __dsub_rn(py, (double)c.j);
    return c;
}

DEVINL Cell m6_cell_fast(const MeshDev &m, double x, double y, int wrap, double &xw, double &yw) {
    double px, py;
    if (wrap == kWrapJulia) {
        const double xn = modulo_fast(x - m.xmin, m.dimx, 1.0 / m.dimx);
        const double yn = modulo_fast(y - m.ymin, m.dimy, 1.0 / m.dimy);
        px = xn / m.dx; py = yn / m.dy;
        xw = xn + m.xmin; yw = yn + m.ymin;
    } else {
        px = modulo_fast(x / m.dx, (double)m.nx, 1.0 / (double)m.nx);
        py = modulo_fast(y / m.dy, (double)m.ny, 1.0 / (double)m.ny);
        xw = x; yw = y;
    }
    Cell c;
    c.i = (int)floor(px); c.dpx = px - (double)c.i;
    c.j = (int)floor(py); c.dpy = py - (double)c.j;
    return c;
}

// node index for offset `off`; the centre (off = 0) is NOT wrapped, as in the reference (i = i+1 after the
// modulo of the neighbours, compute_rho_m6.F90:102-116): it may address the ghost row when px == nx.
DEVINL int wrap_index(int i, int off, int n) {
    if (off == 0) return i;
    int r = (i + off) % n;
    return r < 0 ? r + n : r;
}

// 36 live taps of the 49-term sum of interpolation_m6.F90:130-183 in the reference's order and without
// FMA contraction: bit-identical to the CPU oracle (the 13 dropped terms are exactly +-0)
DEVINL void m6_gather_exact(const MeshDev &m, const double2 *__restrict__ e, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    int ix[6], jy[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        cx[a] = m6_weight_exact(a - 2, c.dpx);
        cy[a] = m6_weight_exact(a - 2, c.dpy);
        ix[a] = wrap_index(c.i, a - 2, m.nx);
        jy[a] = wrap_index(c.j, a - 2, m.ny) * m.ld;
    }
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const double2 ev = __ldg(&e[ix[a] + jy[b]]);
            const double w = __dmul_rn(cx[a], cy[b]);
            s1 = __dadd_rn(s1, __dmul_rn(w, ev.x));
            s2 = __dadd_rn(s2, __dmul_rn(w, ev.y));
        }
    }
    e1 = s1; e2 = s2;
}

// separable form for the fused kernels: rows first (6 FMA per row per component), then the 6 rows
DEVINL void m6_gather_fast(const MeshDev &m, const double2 *__restrict__ e, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    int ix[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) ix[a] = wrap_index(c.i, a - 2, m.nx);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const double2 *row = e + wrap_index(c.j, b - 2, m.ny) * m.ld;
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double2 ev = __ldg(&row[ix[a]]);
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
    }
    e1 = s1; e2 = s2;
}

// ---------------------------------------------------------------------------------------------
// charge accumulation: fp64 atomics (RED.E.ADD.F64 at L2) or int64 fixed point (order independent)
// ---------------------------------------------------------------------------------------------
struct RhoAcc {
    double *f64;               // (nx+1)*(ny+1) doubles, or null
    unsigned long long *i64;   // same extent, fixed point, or null
    double scale;              // 2^S
};

DEVINL void rho_add(const RhoAcc &r, int idx, double val) {
    if (r.i64) {
        const long long q = __double2ll_rn(__dmul_rn(val, r.scale));
        atomicAdd(r.i64 + idx, (unsigned long long)q);
    } else {
        atomicAdd(r.f64 + idx, val);
    }
}

// scatter the 36 live taps of compute_rho_m6.F90:133-187; `part`/`nparts` split the taps over cooperating lanes.
// Weight products follow the reference order (cx*cy)*w without contraction so that fixed-point deposits are
// bit-identical to the oracle's for identical positions.
DEVINL void m6_scatter(const MeshDev &m, const RhoAcc &r, const Cell &c, double weight, int part, int nparts) {
    for (int tap = part; tap < 36; tap += nparts) {
        const int a = tap / 6, b = tap - 6 * a;
        const double cx = m6_weight_exact(a - 2, c.dpx);
        const double cy = m6_weight_exact(b - 2, c.dpy);
        const int idx = wrap_index(c.i, a - 2, m.nx) + wrap_index(c.j, b - 2, m.ny) * m.ld;
        rho_add(r, idx, __dmul_rn(__dmul_rn(cx, cy), weight));
    }
}

// ---------------------------------------------------------------------------------------------
// UA building blocks (all "lane = tau sample")
// ---------------------------------------------------------------------------------------------
struct Pcl { double x1, x2, vx, vy, ex, ey, b, t; };

DEVINL double bfield(double x1, double x2) { return 1.0 + 0.5 * sin(x1) * sin(x2); }   // ua_steps.F90:54

// elt = exp(-i*l*t/eps) for this lane's Fourier slot, phase formed as the reference does: -(l*t)/eps
template <int N> DEVINL cd elt_minus(const TauLane<N> &L, double t, double eps) {
    double s, c;
    sincos(-(L.lf * t) / eps, &s, &c);                                                  // ua_steps.F90:64,224,258
    return mk(c, s);
}

// pl, ql of ua_steps.F90:60-66 for this lane's Fourier slot
template <int N> DEVINL void pl_ql(const TauLane<N> &L, double t, double eps, cd elt, cd &pl, cd &ql) {
    if (L.k == 0) {
        pl = mk(t, 0.0);
        ql = mk(t * t / 2.0, 0.0);
    } else {
        const double l = L.lf;
        // pl = eps*i*(elt-1)/l
        pl = mk((-eps * elt.im) / l, (eps * (elt.re - 1.0)) / l);
        // ql = eps*(eps*(1-elt) - i*l*t)/l^2
        const double l2 = l * l;
        ql = mk((eps * (eps * (1.0 - elt.re))) / l2, (eps * (-eps * elt.im - l * t)) / l2);
    }
}

// preparation, ua_steps.F90:49-113 for one particle spread over N lanes.
// out: p.b, p.t, xt (real), yt (complex), interv (reused by the predictor's compute_f)
template <int N>
DEVINL void prep_particle(const TauLane<N> &L, double eps, double dt, Pcl &p, double &xt1, double &xt2, cd &yt1, cd &yt2,
                          double &interv) {
    p.b = bfield(p.x1, p.x2);
    p.t = dt * p.b;
    const double vxb = p.vx / p.b, vyb = p.vy / p.b;
    const double h1 = eps * (L.st * vxb - L.ct * vyb);
    const double h2 = eps * (L.st * vyb + L.ct * vxb);
    xt1 = p.x1 + h1 + eps * vyb;
    xt2 = p.x2 + h2 - eps * vxb;
    interv = (1.0 + 0.5 * sin(xt1) * sin(xt2) - p.b) / eps;
    const double exb = ((L.ct * p.vy - L.st * p.vx) * interv + p.ex) / p.b;
    const double eyb = ((-L.ct * p.vx - L.st * p.vy) * interv + p.ey) / p.b;
    cd r1 = mk(L.ct * exb - L.st * eyb, 0.0);
    cd r2 = mk(L.st * exb + L.ct * eyb, 0.0);
    r1 = fft_fwd<N>(r1, L);
    r2 = fft_fwd<N>(r2, L);
    if (L.k != 0) {
        // rf = -(i/l) * rf / N        ua_steps.F90:100-103
        const double s = 1.0 / (L.lf * (double)N);
        r1 = mk(r1.im * s, -r1.re * s);
        r2 = mk(r2.im * s, -r2.re * s);
    }
    r1 = fft_bwd<N>(r1, L);
    r2 = fft_bwd<N>(r2, L);
    const cd r10 = group_bcast0<N>(r1), r20 = group_bcast0<N>(r2);
    yt1 = mk(p.vx + (r1.re - r10.re) * eps, (r1.im - r10.im) * eps);                    // :109
    yt2 = mk(p.vy + (r2.re - r20.re) * eps, (r2.im - r20.im) * eps);                    // :110
}

// compute_f before the tau FFT, ua_steps.F90:164-185
template <int N>
DEVINL void force_terms(const TauLane<N> &L, double rb, double interv, cd yt1, cd yt2, double et1, double et2,
                        cd &fx1, cd &fx2, cd &fy1, cd &fy2) {
    const double ct = L.ct, st = L.st;
    fx1 = mk((ct * yt1.re + st * yt2.re) * rb, (ct * yt1.im + st * yt2.im) * rb);      // :174
    fx2 = mk((ct * yt2.re - st * yt1.re) * rb, (ct * yt2.im - st * yt1.im) * rb);      // :175
    const cd t1 = mk(et1 + (ct * yt2.re - st * yt1.re) * interv, (ct * yt2.im - st * yt1.im) * interv);     // :179
    const cd t2 = mk(et2 - (ct * yt1.re + st * yt2.re) * interv, -(ct * yt1.im + st * yt2.im) * interv);    // :180
    fy1 = mk((ct * t1.re - st * t2.re) * rb, (ct * t1.im - st * t2.im) * rb);          // :182
    fy2 = mk((st * t1.re + ct * t2.re) * rb, (st * t1.im + ct * t2.im) * rb);          // :183
}

// tau* evaluation of compute_rho_m6.F90:74-84 / ua_steps.F90:293-300 from *normalised* Fourier coefficients:
// returns sum_k xhat_k * exp(+i l_k t/eps) (complex), identical on all lanes of the group
template <int N> DEVINL cd eval_tau_star(cd xhat, cd elt /* exp(-i l t/eps) */) {
    const cd z = cmulc(xhat, elt);   // xhat * conj(elt) = xhat * exp(+i l t/eps)
    return mk(group_sum<N>(z.re), group_sum<N>(z.im));
}

// warp/particle bookkeeping for "lane = tau sample" kernels
template <int N> struct WarpMap {
    static constexpr int G = 32 / N;   // particles per warp
    int lane, g;
    int64_t first, stride;             // first particle of this warp, particles per grid sweep
    DEVINL WarpMap() {
        lane = threadIdx.x & 31;
        g = lane / N;
        const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        first = warp * G;
        stride = nwarps * G;
    }
};

DEVINL double2 ld2(const double *p, int64_t i) { return reinterpret_cast<const double2 *>(p)[i]; }
DEVINL void st2(double *p, int64_t i, cd z) { reinterpret_cast<double2 *>(p)[i] = make_double2(z.re, z.im); }
DEVINL cd ldc(const double *p, int64_t i) { const double2 d = ld2(p, i); return mk(d.x, d.y); }


}  // namespace uapic

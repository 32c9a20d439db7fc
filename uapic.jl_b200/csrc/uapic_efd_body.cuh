// uapic_efd_body.cuh -- the per-particle arithmetic of the external-field program (fortran/efd.f90:133-478), written once
// against a small "tau policy" P that says where the tau samples of a particle live and how a length-ntau transform is done:
//
//     P::SPL                 samples (= Fourier slots) held by this thread
//     T.ct(j), T.st(j)       cos / sin of the tau sample in slot j                       efd.f90:113-116
//     T.lmode(j)             wavenumber of Fourier slot j, T.mode_live(j): slot in use    efd.f90:111-112
//     T.fwd(a), T.inv(a)     forward transform carrying 1/ntau, unnormalised backward    fft.f90:44-81
//     T.fwd2(a, b), T.inv2(a, b)   the same on the two components at once (fft_2d / ifft_2d, fft.f90:37-59)
//     T.first(a)             slot 0 of the particle (tau = 0, or mode 0), on every thread of the particle
//     T.sum(z)               sum of z over all the threads of the particle
//
// uapic_efd.cu instantiates it with the two device policies (one sample per lane / one warp per particle).  The file uses
// nothing CUDA-specific beyond DEVINL, the complex helpers cd / mk / cmul and the transcendentals efd_sin / efd_sincos (out-of-line
// wrappers on the device: inlined, the ~100 expansions of sin/cos made the kernel 110 KB of SASS, see uapic_efd.cu), so tests/efd_host_body.cpp can compile the very
// same text for the host with a one-thread policy and hold it to the oracle without a GPU.
#pragma once

namespace uapic {

namespace {

struct EfdScalars {
    double eps, dt, tfinal, xmin, xmax, ymin, ymax;
    int nstep;
};

DEVINL cd operator+(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
DEVINL cd operator-(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
DEVINL cd operator+(double a, cd b) { return mk(a + b.re, b.im); }
DEVINL cd operator+(cd b, double a) { return mk(b.re + a, b.im); }
DEVINL cd operator-(cd b, double a) { return mk(b.re - a, b.im); }
DEVINL cd operator-(double a, cd b) { return mk(a - b.re, -b.im); }
DEVINL cd operator*(double a, cd b) { return mk(a * b.re, a * b.im); }
DEVINL cd operator*(cd b, double a) { return mk(a * b.re, a * b.im); }
DEVINL cd operator/(cd b, double a) { return mk(b.re / a, b.im / a); }
DEVINL cd operator-(cd b) { return mk(-b.re, -b.im); }
DEVINL cd operator*(cd a, cd b) { return cmul(a, b); }
DEVINL cd mul_mi(cd a) { return mk(a.im, -a.re); }                       // -i a
DEVINL cd over_1pia(cd z, double a, double rd) {                         // z / (1 + i a), rd = 1 / (1 + a^2)
    return mk((z.re + z.im * a) * rd, (z.im - z.re * a) * rd);
}
DEVINL double bfun(double a, double b) { return 1.0 + 0.5 * efd_sin(a) * efd_sin(b); }    // efd.f90:139

#define EACH _Pragma("unroll") for (int j = 0; j < S; ++j)

// tilde(n) = -i tilde(n) / ltau(n) for n >= 2, tilde(1) = 0                       efd.f90:183-187 and its repeats
// (il = 1 / ltau, 0 for mode 0 and for unused slots)
template <int S> DEVINL void primitive_multiplier(const double (&il)[S], cd (&a)[S]) {
    EACH a[j] = mul_mi(a[j]) * il[j];
}
// zero-mean tau-primitive of both components: forward, multiplier, backward
template <class P, int S> DEVINL void primitive2(const P &T, const double (&il)[S], cd (&a)[S], cd (&b)[S]) {
    T.fwd2(a, b); primitive_multiplier<S>(il, a); primitive_multiplier<S>(il, b); T.inv2(a, b);
}
// out(:) = base + a(:) - a(1)                                                     efd.f90:163-164,194-195,221-222,...
template <class P, int S> DEVINL void rebase(const P &T, cd (&out)[S], double base, const cd (&a)[S]) {
    const cd a0 = T.first(a);
    EACH out[j] = base + a[j] - a0;
}
// compute_fy                                                                      efd.f90:509-524
// (ibx = 1/b, ibe = 1/b/eps: the loop-invariant divisions of the reference are hoisted into reciprocals, and sin(a) comes from the
//  half-angle pair the field needs anyway -- each a <= 1 ulp change, far inside the conditioning of the map, see tests/test_gpu_efd.py)
template <class P, int S>
DEVINL void force(const P &T, double ibx, double ibe, double bx, double time, const cd (&X1)[S], const cd (&X2)[S], const cd (&Y1)[S],
                  const cd (&Y2)[S], cd (&f1)[S], cd (&f2)[S]) {
    const double amp = 1.0 + 0.5 * efd_sin(time);
    EACH {
        const double a = X1[j].re, b = X2[j].re;
        double sh, ch, sb, cb;
        efd_sincos(a / 2.0, &sh, &ch);
        efd_sincos(b, &sb, &cb);
        const double e1 = (0.5 * ch * sb) * amp;
        const double e2 = (cb * sh) * amp;
        const double interv = (1.0 + 0.5 * (2.0 * sh * ch) * sb - bx) * ibe;
        const double t1 = (T.ct(j) * e1 - T.st(j) * e2) * ibx;
        const double t2 = (T.ct(j) * e2 + T.st(j) * e1) * ibx;
        f1[j] = t1 + interv * Y2[j];
        f2[j] = t2 - interv * Y1[j];
    }
}

template <class P>
DEVINL void efd_particle(const P &T, const EfdScalars &q, double x1, double x2, double v1, double v2, double (&xo)[2], double (&vo)[2]) {
    constexpr int S = P::SPL;
    const double eps = q.eps;
    double time = 0.0;
    // the transcendentals of the particle's own position, once (the reference re-evaluates them where they appear), and
    // sin(time), cos(time) of the preparation, which runs at time = 0 (efd.f90:137)
    double sx1, cx1, sx2, cx2, sh1, ch1;
    efd_sincos(x1, &sx1, &cx1);
    efd_sincos(x2, &sx2, &cx2);
    efd_sincos(x1 / 2.0, &sh1, &ch1);
    const double st0 = 0.0, ct0 = 1.0;
    const double bx = 1.0 + 0.5 * sx1 * sx2;                                                 // efd.f90:138-140
    const double ds = q.dt * bx;
    // Reciprocals formed once: the reference divides by b, eps and ltau wherever they appear (about eighty fp64 divisions per
    // particle and tau sample, ~25 instructions each); multiplying by 1/b, 1/eps, 1/ltau instead moves a result by <= 1 ulp,
    // against a conditioning of the map that moves v by 1e-11 per input ulp at eps = 1e-3 (tests/test_gpu_efd.py).
    const double ibx = 1.0 / bx, ieps = 1.0 / eps, ibe = ibx * ieps;
    const double eb = eps * ibx, e2b = eps * eps * ibx;
    cd xt1[S], xt2[S], yt1[S], yt2[S], r1[S], r2[S], t1[S], t2[S], f1[S], f2[S];
    double il_[S];                                       // 1 / ltau, 0 for mode 0 and for unused slots
    EACH { const double l = T.lmode(j); il_[j] = (T.mode_live(j) && l != 0.0) ? 1.0 / l : 0.0; }

    // ---- first-order datum (efd.f90:157-195) --------------------------------------------------------------------
    EACH {
        t1[j] = mk(eps * (T.st(j) * (v1 * ibx) - T.ct(j) * (v2 * ibx)), 0.0);
        t2[j] = mk(eps * (T.st(j) * (v2 * ibx) + T.ct(j) * (v1 * ibx)), 0.0);
    }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);
    double e1 = (0.5 * ch1 * sx2) * (1.0 + 0.5 * st0);
    double e2 = (sh1 * cx2) * (1.0 + 0.5 * st0);
    EACH {
        const double interv = (bfun(xt1[j].re, xt2[j].re) - bx) * ibx;
        r1[j] = mk(interv * v2, 0.0);
        r2[j] = mk(-interv * v1, 0.0);
    }
    T.fwd2(r1, r2);
    const double ave1 = T.first(r1).re * ieps, ave2 = T.first(r2).re * ieps;                   // efd.f90:180
    primitive_multiplier<S>(il_, r1); primitive_multiplier<S>(il_, r2);
    T.inv2(r1, r2);
    EACH {
        r1[j] = eb * (T.st(j) * e1 + T.ct(j) * e2) + r1[j];
        r2[j] = eb * (T.st(j) * e2 - T.ct(j) * e1) + r2[j];
    }
    rebase<P, S>(T, yt1, v1, r1); rebase<P, S>(T, yt2, v2, r2);

    // ---- second-order position (efd.f90:200-222) ----------------------------------------------------------------
    EACH {
        t1[j] = eb * (T.ct(j) * yt1[j] + T.st(j) * yt2[j]);
        t2[j] = eb * (T.ct(j) * yt2[j] - T.st(j) * yt1[j]);
    }
    primitive2<P, S>(T, il_, t1, t2);
    EACH {
        t1[j] = t1[j] - e2b * (-T.ct(j) * ave1 - T.st(j) * ave2);
        t2[j] = t2[j] - e2b * (-T.ct(j) * ave2 + T.st(j) * ave1);
    }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);

    // ---- second-order velocity (efd.f90:226-310): the time derivative of E enters here ---------------------------
    e1 = (0.5 * ch1 * sx2) * 0.5 * ct0;
    e2 = (sh1 * cx2) * 0.5 * ct0;
    EACH {
        const double interv = (bfun(xt1[j].re, xt2[j].re) - bx) * ibx;
        double fx1 = interv * ave2, fx2 = -interv * ave1;
        double fy1 = eb * (T.st(j) * ave1 - T.ct(j) * ave2);
        double fy2 = eb * (T.ct(j) * ave1 + T.st(j) * ave2);
        const double w = cx1 * sx2 * fy1 + sx1 * cx2 * fy2;
        fy1 = w * ibx * 0.5 * v2 + fx1;
        fy2 = -w * ibx * 0.5 * v1 + fx2;
        fx1 = eb * ibx * (-T.st(j) * e2 + T.ct(j) * e1);
        fx2 = eb * ibx * (T.st(j) * e1 + T.ct(j) * e2);
        t1[j] = mk(fy1 + fx1, 0.0);
        t2[j] = mk(fy2 + fx2, 0.0);
    }
    T.fwd2(t1, t2);
    EACH {                                                                                   // efd.f90:260-266
        f1[j] = mul_mi(t1[j]) * il_[j];
        f2[j] = mul_mi(t2[j]) * il_[j];
        t1[j] = -t1[j] * (il_[j] * il_[j]);
        t2[j] = -t2[j] * (il_[j] * il_[j]);
    }
    T.inv2(t1, t2);
    EACH { r1[j] = -eps * t1[j]; r2[j] = -eps * t2[j]; }
    T.inv2(f1, f2);                                                                    // fy of efd.f90:273
    EACH {
        const double a = xt1[j].re, b = xt2[j].re;
        double sh, ch, sb, cb;
        efd_sincos(a / 2.0, &sh, &ch);
        efd_sincos(b, &sb, &cb);
        const double en1 = (0.5 * ch * sb) * (1.0 + 0.5 * st0);
        const double en2 = (sh * cb) * (1.0 + 0.5 * st0);
        const double interv = (1.0 + 0.5 * (2.0 * sh * ch) * sb - bx) * ibx;
        t1[j] = interv * yt2[j] + eb * (-T.st(j) * en2 + T.ct(j) * en1);
        t2[j] = -interv * yt1[j] + eb * (T.st(j) * en1 + T.ct(j) * en2);
    }
    T.fwd2(t1, t2);
    const cd yd1 = T.first(t1) * ieps, yd2 = T.first(t2) * ieps;                               // xf(1,:), efd.f90:299
    primitive_multiplier<S>(il_, t1); primitive_multiplier<S>(il_, t2);
    T.inv2(t1, t2);
    EACH { r1[j] = r1[j] + t1[j]; r2[j] = r2[j] + t2[j]; }
    rebase<P, S>(T, yt1, v1, r1); rebase<P, S>(T, yt2, v2, r2);

    // ---- third-order position (efd.f90:315-383) -----------------------------------------------------------------
    EACH {
        t1[j] = (T.ct(j) * r1[j] + T.st(j) * r2[j]) * ibx;
        t2[j] = (T.ct(j) * r2[j] - T.st(j) * r1[j]) * ibx;
    }
    T.fwd2(t1, t2);
    const double w0 = cx1 * sx2 * T.first(t1).re + sx1 * cx2 * T.first(t2).re;   // `interv` is real(8): real part
    cd g1 = mk(w0 * ibe * v2 * 0.5, 0.0), g2 = mk(-w0 * ibe * v1 * 0.5, 0.0);
    EACH t1[j] = mk((bfun(xt1[j].re, xt2[j].re) - bx) * ibx, 0.0);
    T.fwd(t1);
    const cd q0 = T.first(t1);
    g1 = g1 + q0 * ieps * ave2;
    g2 = g2 - q0 * ieps * ave1;
    EACH {
        const cd yf1 = yd1 + f1[j], yf2 = yd2 + f2[j];
        t1[j] = T.ct(j) * yf1 + T.st(j) * yf2;
        t2[j] = T.ct(j) * yf2 - T.st(j) * yf1;
    }
    primitive2<P, S>(T, il_, t1, t2);
    EACH {
        f1[j] = t1[j] * eb - e2b * (-T.ct(j) * g1 - T.st(j) * g2);
        f2[j] = t2[j] * eb - e2b * (-T.ct(j) * g2 + T.st(j) * g1);
    }
    primitive2<P, S>(T, il_, f1, f2);
    EACH {
        t1[j] = eb * (T.ct(j) * yt1[j] + T.st(j) * yt2[j]);
        t2[j] = eb * (T.ct(j) * yt2[j] - T.st(j) * yt1[j]);
    }
    primitive2<P, S>(T, il_, t1, t2);
    EACH { t1[j] = -eps * f1[j] + t1[j]; t2[j] = -eps * f2[j] + t2[j]; }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);

    // ---- IMEX2 steps (efd.f90:388-454) --------------------------------------------------------------------------
    // per-slot constants of the two spectral operators: z / (1 + i a) = z (1 - i a) / (1 + a^2) and (1 - i an)
    double a_[S], rd_[S], an_[S];
    EACH {
        a_[j] = ds / 2.0 * T.lmode(j) / eps;
        an_[j] = ds / eps / 2.0 * T.lmode(j);
        rd_[j] = 1.0 / (1.0 + a_[j] * a_[j]);
    }
    for (int istep = 0; istep < q.nstep; ++istep) {
        force<P, S>(T, ibx, ibe, bx, time, xt1, xt2, yt1, yt2, f1, f2);
        EACH { r1[j] = yt1[j] + ds / 2.0 * f1[j]; r2[j] = yt2[j] + ds / 2.0 * f2[j]; }
        T.fwd2(r1, r2);
        EACH { r1[j] = over_1pia(r1[j], a_[j], rd_[j]); r2[j] = over_1pia(r2[j], a_[j], rd_[j]); }
        T.inv2(r1, r2);                                                                // yt(tn+1/2)
        EACH {
            t1[j] = xt1[j] + ds / 2.0 * ((T.ct(j) * r1[j] + T.st(j) * r2[j]) * ibx);
            t2[j] = xt2[j] + ds / 2.0 * ((T.ct(j) * r2[j] - T.st(j) * r1[j]) * ibx);
        }
        T.fwd2(t1, t2);
        EACH { t1[j] = over_1pia(t1[j], a_[j], rd_[j]); t2[j] = over_1pia(t2[j], a_[j], rd_[j]); }
        T.inv2(t1, t2);                                                                // xt(tn+1/2)
        time = time + q.dt / 2.0;
        force<P, S>(T, ibx, ibe, bx, time, t1, t2, r1, r2, f1, f2);
        T.fwd2(f1, f2);
        EACH { r1[j] = yt1[j]; r2[j] = yt2[j]; }
        T.fwd2(r1, r2);
        EACH {
            const cd nm = mk(1.0, -an_[j]);
            r1[j] = over_1pia(r1[j] * nm + ds * f1[j], a_[j], rd_[j]);
            r2[j] = over_1pia(r2[j] * nm + ds * f2[j], a_[j], rd_[j]);
        }
        T.inv2(r1, r2);                                                                // yt(tn+1)
        EACH {
            const cd m1 = (r1[j] + yt1[j]) * 0.5, m2 = (r2[j] + yt2[j]) * 0.5;
            yt1[j] = r1[j]; yt2[j] = r2[j];
            f1[j] = (T.ct(j) * m1 + T.st(j) * m2) * ibx;
            f2[j] = (T.ct(j) * m2 - T.st(j) * m1) * ibx;
        }
        T.fwd2(f1, f2);
        T.fwd2(xt1, xt2);
        EACH {
            const cd nm = mk(1.0, -an_[j]);
            xt1[j] = over_1pia(xt1[j] * nm + ds * f1[j], a_[j], rd_[j]);
            xt2[j] = over_1pia(xt2[j] * nm + ds * f2[j], a_[j], rd_[j]);
        }
        T.inv2(xt1, xt2);                                                              // xt(tn+1)
        time = time + q.dt / 2.0;
    }

    // ---- physical state at tau = tfinal b / eps (efd.f90:456-478), wrapped as apply_bc (efd.f90:526-544) ----------
    T.fwd2(xt1, xt2); T.fwd2(yt1, yt2);
    cd ax1 = mk(0.0, 0.0), ax2 = ax1, ay1 = ax1, ay2 = ax1;
    EACH {
        if (T.mode_live(j)) {
            double sp, cp;
            efd_sincos(T.lmode(j) * q.tfinal * bx / eps, &sp, &cp);
            const cd ph = mk(cp, sp);
            ax1 = ax1 + xt1[j] * ph; ax2 = ax2 + xt2[j] * ph;
            ay1 = ay1 + yt1[j] * ph; ay2 = ay2 + yt2[j] * ph;
        }
    }
    ax1 = T.sum(ax1); ax2 = T.sum(ax2); ay1 = T.sum(ay1); ay2 = T.sum(ay2);
    double xx = ax1.re, yy = ax2.re;
    const double dimx = q.xmax - q.xmin, dimy = q.ymax - q.ymin;
    for (int it = 0; it < 1024 && xx > q.xmax; ++it) xx -= dimx;
    for (int it = 0; it < 1024 && xx < q.xmin; ++it) xx += dimx;
    for (int it = 0; it < 1024 && yy > q.ymax; ++it) yy -= dimy;
    for (int it = 0; it < 1024 && yy < q.ymin; ++it) yy += dimy;
    double sb, cb;
    efd_sincos(q.tfinal * bx / eps, &sb, &cb);
    xo[0] = xx; xo[1] = yy;
    vo[0] = cb * ay1.re + sb * ay2.re;
    vo[1] = cb * ay2.re - sb * ay1.re;
}

#undef EACH

}  // namespace

}  // namespace uapic

// uapic_efd_body.cuh -- the per-particle arithmetic of the external-field program (fortran/efd.f90:133-478), written once
// against a small "tau policy" P that says where the tau samples of a particle live and how a length-ntau transform is done:
//
//     P::SPL                 samples (= Fourier slots) held by this thread
//     T.ct(j), T.st(j)       cos / sin of the tau sample in slot j                       efd.f90:113-116
//     T.lmode(j)             wavenumber of Fourier slot j, T.mode_live(j): slot in use    efd.f90:111-112
//     T.fwd(a), T.inv(a)     forward transform carrying 1/ntau, unnormalised backward    fft.f90:44-81
//     T.first(a)             slot 0 of the particle (tau = 0, or mode 0), on every thread of the particle
//     T.sum(z)               sum of z over all the threads of the particle
//
// uapic_efd.cu instantiates it with the two device policies (one sample per lane / one warp per particle).  The file uses
// nothing CUDA-specific beyond DEVINL and the complex helpers cd / mk / cmul, so tests/efd_host_body.cpp can compile the very
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
DEVINL double bfun(double a, double b) { return 1.0 + 0.5 * sin(a) * sin(b); }    // efd.f90:139

#define EACH _Pragma("unroll") for (int j = 0; j < S; ++j)

// tilde(n) = -i tilde(n) / ltau(n) for n >= 2, tilde(1) = 0                       efd.f90:183-187 and its repeats
template <class P, int S> DEVINL void primitive_multiplier(const P &T, cd (&a)[S]) {
    EACH {
        const double l = T.lmode(j);
        a[j] = (T.mode_live(j) && l != 0.0) ? mul_mi(a[j]) / l : mk(0.0, 0.0);
    }
}
// zero-mean tau-primitive: forward, multiplier, backward
template <class P, int S> DEVINL void primitive(const P &T, cd (&a)[S]) {
    T.fwd(a); primitive_multiplier<P, S>(T, a); T.inv(a);
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
    const double amp = 1.0 + 0.5 * sin(time);
    EACH {
        const double a = X1[j].re, b = X2[j].re;
        double sh, ch, sb, cb;
        sincos(a / 2.0, &sh, &ch);
        sincos(b, &sb, &cb);
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
    const double bx = bfun(x1, x2);                                                          // efd.f90:138-140
    const double ds = q.dt * bx;
    cd xt1[S], xt2[S], yt1[S], yt2[S], r1[S], r2[S], t1[S], t2[S], f1[S], f2[S];

    // ---- first-order datum (efd.f90:157-195) --------------------------------------------------------------------
    EACH {
        t1[j] = mk(eps * (T.st(j) * (v1 / bx) - T.ct(j) * (v2 / bx)), 0.0);
        t2[j] = mk(eps * (T.st(j) * (v2 / bx) + T.ct(j) * (v1 / bx)), 0.0);
    }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);
    double e1 = (0.5 * cos(x1 / 2.0) * sin(x2)) * (1.0 + 0.5 * sin(time));
    double e2 = (sin(x1 / 2.0) * cos(x2)) * (1.0 + 0.5 * sin(time));
    EACH {
        const double interv = (bfun(xt1[j].re, xt2[j].re) - bx) / bx;
        r1[j] = mk(interv * v2, 0.0);
        r2[j] = mk(-interv * v1, 0.0);
    }
    T.fwd(r1); T.fwd(r2);
    const double ave1 = T.first(r1).re / eps, ave2 = T.first(r2).re / eps;                   // efd.f90:180
    primitive_multiplier<P, S>(T, r1); primitive_multiplier<P, S>(T, r2);
    T.inv(r1); T.inv(r2);
    EACH {
        r1[j] = eps * (T.st(j) * e1 + T.ct(j) * e2) / bx + r1[j];
        r2[j] = eps * (T.st(j) * e2 - T.ct(j) * e1) / bx + r2[j];
    }
    rebase<P, S>(T, yt1, v1, r1); rebase<P, S>(T, yt2, v2, r2);

    // ---- second-order position (efd.f90:200-222) ----------------------------------------------------------------
    EACH {
        t1[j] = eps * (T.ct(j) * yt1[j] + T.st(j) * yt2[j]) / bx;
        t2[j] = eps * (T.ct(j) * yt2[j] - T.st(j) * yt1[j]) / bx;
    }
    primitive<P, S>(T, t1); primitive<P, S>(T, t2);
    EACH {
        t1[j] = t1[j] - eps * eps / bx * (-T.ct(j) * ave1 - T.st(j) * ave2);
        t2[j] = t2[j] - eps * eps / bx * (-T.ct(j) * ave2 + T.st(j) * ave1);
    }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);

    // ---- second-order velocity (efd.f90:226-310): the time derivative of E enters here ---------------------------
    e1 = (0.5 * cos(x1 / 2.0) * sin(x2)) * 0.5 * cos(time);
    e2 = (sin(x1 / 2.0) * cos(x2)) * 0.5 * cos(time);
    EACH {
        const double interv = (bfun(xt1[j].re, xt2[j].re) - bx) / bx;
        double fx1 = interv * ave2, fx2 = -interv * ave1;
        double fy1 = eps / bx * (T.st(j) * ave1 - T.ct(j) * ave2);
        double fy2 = eps / bx * (T.ct(j) * ave1 + T.st(j) * ave2);
        const double w = cos(x1) * sin(x2) * fy1 + sin(x1) * cos(x2) * fy2;
        fy1 = w / bx / 2.0 * v2 + fx1;
        fy2 = -w / bx / 2.0 * v1 + fx2;
        fx1 = eps / (bx * bx) * (-T.st(j) * e2 + T.ct(j) * e1);
        fx2 = eps / (bx * bx) * (T.st(j) * e1 + T.ct(j) * e2);
        t1[j] = mk(fy1 + fx1, 0.0);
        t2[j] = mk(fy2 + fx2, 0.0);
    }
    T.fwd(t1); T.fwd(t2);
    EACH {                                                                                   // efd.f90:260-266
        const double l = T.lmode(j);
        const bool on = T.mode_live(j) && l != 0.0;
        f1[j] = on ? mul_mi(t1[j]) / l : mk(0.0, 0.0);
        f2[j] = on ? mul_mi(t2[j]) / l : mk(0.0, 0.0);
        t1[j] = on ? -t1[j] / (l * l) : mk(0.0, 0.0);
        t2[j] = on ? -t2[j] / (l * l) : mk(0.0, 0.0);
    }
    T.inv(t1); T.inv(t2);
    EACH { r1[j] = -eps * t1[j]; r2[j] = -eps * t2[j]; }
    T.inv(f1); T.inv(f2);                                                                    // fy of efd.f90:273
    EACH {
        const double a = xt1[j].re, b = xt2[j].re;
        const double en1 = (0.5 * cos(a / 2.0) * sin(b)) * (1.0 + 0.5 * sin(time));
        const double en2 = (sin(a / 2.0) * cos(b)) * (1.0 + 0.5 * sin(time));
        const double interv = (bfun(a, b) - bx) / bx;
        t1[j] = interv * yt2[j] + eps / bx * (-T.st(j) * en2 + T.ct(j) * en1);
        t2[j] = -interv * yt1[j] + eps / bx * (T.st(j) * en1 + T.ct(j) * en2);
    }
    T.fwd(t1); T.fwd(t2);
    const cd yd1 = T.first(t1) / eps, yd2 = T.first(t2) / eps;                               // xf(1,:), efd.f90:299
    primitive_multiplier<P, S>(T, t1); primitive_multiplier<P, S>(T, t2);
    T.inv(t1); T.inv(t2);
    EACH { r1[j] = r1[j] + t1[j]; r2[j] = r2[j] + t2[j]; }
    rebase<P, S>(T, yt1, v1, r1); rebase<P, S>(T, yt2, v2, r2);

    // ---- third-order position (efd.f90:315-383) -----------------------------------------------------------------
    EACH {
        t1[j] = (T.ct(j) * r1[j] + T.st(j) * r2[j]) / bx;
        t2[j] = (T.ct(j) * r2[j] - T.st(j) * r1[j]) / bx;
    }
    T.fwd(t1); T.fwd(t2);
    const double w0 = cos(x1) * sin(x2) * T.first(t1).re + sin(x1) * cos(x2) * T.first(t2).re;   // `interv` is real(8): real part
    cd g1 = mk(w0 / eps / bx * v2 / 2.0, 0.0), g2 = mk(-w0 / eps / bx * v1 / 2.0, 0.0);
    EACH t1[j] = mk((bfun(xt1[j].re, xt2[j].re) - bx) / bx, 0.0);
    T.fwd(t1);
    const cd q0 = T.first(t1);
    g1 = g1 + q0 / eps * ave2;
    g2 = g2 - q0 / eps * ave1;
    EACH {
        const cd yf1 = yd1 + f1[j], yf2 = yd2 + f2[j];
        t1[j] = T.ct(j) * yf1 + T.st(j) * yf2;
        t2[j] = T.ct(j) * yf2 - T.st(j) * yf1;
    }
    primitive<P, S>(T, t1); primitive<P, S>(T, t2);
    EACH {
        f1[j] = t1[j] * eps / bx - eps * eps / bx * (-T.ct(j) * g1 - T.st(j) * g2);
        f2[j] = t2[j] * eps / bx - eps * eps / bx * (-T.ct(j) * g2 + T.st(j) * g1);
    }
    primitive<P, S>(T, f1); primitive<P, S>(T, f2);
    EACH {
        t1[j] = eps * (T.ct(j) * yt1[j] + T.st(j) * yt2[j]) / bx;
        t2[j] = eps * (T.ct(j) * yt2[j] - T.st(j) * yt1[j]) / bx;
    }
    primitive<P, S>(T, t1); primitive<P, S>(T, t2);
    EACH { t1[j] = -eps * f1[j] + t1[j]; t2[j] = -eps * f2[j] + t2[j]; }
    rebase<P, S>(T, xt1, x1, t1); rebase<P, S>(T, xt2, x2, t2);

    // ---- IMEX2 steps (efd.f90:388-454) --------------------------------------------------------------------------
    // per-slot constants of the two spectral operators: z / (1 + i a) = z (1 - i a) / (1 + a^2) and (1 - i an)
    const double ibx = 1.0 / bx, ibe = 1.0 / bx / eps;
    double a_[S], rd_[S], an_[S];
    EACH {
        a_[j] = ds / 2.0 * T.lmode(j) / eps;
        an_[j] = ds / eps / 2.0 * T.lmode(j);
        rd_[j] = 1.0 / (1.0 + a_[j] * a_[j]);
    }
    for (int istep = 0; istep < q.nstep; ++istep) {
        force<P, S>(T, ibx, ibe, bx, time, xt1, xt2, yt1, yt2, f1, f2);
        EACH { r1[j] = yt1[j] + ds / 2.0 * f1[j]; r2[j] = yt2[j] + ds / 2.0 * f2[j]; }
        T.fwd(r1); T.fwd(r2);
        EACH { r1[j] = over_1pia(r1[j], a_[j], rd_[j]); r2[j] = over_1pia(r2[j], a_[j], rd_[j]); }
        T.inv(r1); T.inv(r2);                                                                // yt(tn+1/2)
        EACH {
            t1[j] = xt1[j] + ds / 2.0 * ((T.ct(j) * r1[j] + T.st(j) * r2[j]) * ibx);
            t2[j] = xt2[j] + ds / 2.0 * ((T.ct(j) * r2[j] - T.st(j) * r1[j]) * ibx);
        }
        T.fwd(t1); T.fwd(t2);
        EACH { t1[j] = over_1pia(t1[j], a_[j], rd_[j]); t2[j] = over_1pia(t2[j], a_[j], rd_[j]); }
        T.inv(t1); T.inv(t2);                                                                // xt(tn+1/2)
        time = time + q.dt / 2.0;
        force<P, S>(T, ibx, ibe, bx, time, t1, t2, r1, r2, f1, f2);
        T.fwd(f1); T.fwd(f2);
        EACH { r1[j] = yt1[j]; r2[j] = yt2[j]; }
        T.fwd(r1); T.fwd(r2);
        EACH {
            const cd nm = mk(1.0, -an_[j]);
            r1[j] = over_1pia(r1[j] * nm + ds * f1[j], a_[j], rd_[j]);
            r2[j] = over_1pia(r2[j] * nm + ds * f2[j], a_[j], rd_[j]);
        }
        T.inv(r1); T.inv(r2);                                                                // yt(tn+1)
        EACH {
            const cd m1 = (r1[j] + yt1[j]) * 0.5, m2 = (r2[j] + yt2[j]) * 0.5;
            yt1[j] = r1[j]; yt2[j] = r2[j];
            f1[j] = (T.ct(j) * m1 + T.st(j) * m2) * ibx;
            f2[j] = (T.ct(j) * m2 - T.st(j) * m1) * ibx;
        }
        T.fwd(f1); T.fwd(f2);
        T.fwd(xt1); T.fwd(xt2);
        EACH {
            const cd nm = mk(1.0, -an_[j]);
            xt1[j] = over_1pia(xt1[j] * nm + ds * f1[j], a_[j], rd_[j]);
            xt2[j] = over_1pia(xt2[j] * nm + ds * f2[j], a_[j], rd_[j]);
        }
        T.inv(xt1); T.inv(xt2);                                                              // xt(tn+1)
        time = time + q.dt / 2.0;
    }

    // ---- physical state at tau = tfinal b / eps (efd.f90:456-478), wrapped as apply_bc (efd.f90:526-544) ----------
    T.fwd(xt1); T.fwd(xt2); T.fwd(yt1); T.fwd(yt2);
    cd sx1 = mk(0.0, 0.0), sx2 = sx1, sy1 = sx1, sy2 = sx1;
    EACH {
        if (T.mode_live(j)) {
            double sp, cp;
            sincos(T.lmode(j) * q.tfinal * bx / eps, &sp, &cp);
            const cd ph = mk(cp, sp);
            sx1 = sx1 + xt1[j] * ph; sx2 = sx2 + xt2[j] * ph;
            sy1 = sy1 + yt1[j] * ph; sy2 = sy2 + yt2[j] * ph;
        }
    }
    sx1 = T.sum(sx1); sx2 = T.sum(sx2); sy1 = T.sum(sy1); sy2 = T.sum(sy2);
    double xx = sx1.re, yy = sx2.re;
    const double dimx = q.xmax - q.xmin, dimy = q.ymax - q.ymin;
    for (int it = 0; it < 1024 && xx > q.xmax; ++it) xx -= dimx;
    for (int it = 0; it < 1024 && xx < q.xmin; ++it) xx += dimx;
    for (int it = 0; it < 1024 && yy > q.ymax; ++it) yy -= dimy;
    for (int it = 0; it < 1024 && yy < q.ymin; ++it) yy += dimy;
    double sb, cb;
    sincos(q.tfinal * bx / eps, &sb, &cb);
    xo[0] = xx; xo[1] = yy;
    vo[0] = cb * sy1.re + sb * sy2.re;
    vo[1] = cb * sy2.re - sb * sy1.re;
}

#undef EACH

}  // namespace

}  // namespace uapic

// efd_host_body.cpp -- TEST HARNESS (never part of the product): compiles the per-particle body of the CUDA external-field
// kernel (uapic.jl_b200/csrc/uapic_efd_body.cuh, the text the device runs) for the HOST with a one-thread tau policy, so that
// tests/test_efd_host_body.py can hold that text to the oracle without a GPU.  Built by the test itself with g++.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cmath>
#include <cstdint>

#define DEVINL inline
namespace uapic {
struct cd { double re, im; };
DEVINL cd mk(double r, double i) { cd z; z.re = r; z.im = i; return z; }
DEVINL cd cmul(cd a, cd b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline double efd_sin(double x) { return std::sin(x); }
inline void efd_sincos(double x, double *s, double *c) { *s = std::sin(x); *c = std::cos(x); }
}  // namespace uapic
#include "../uapic.jl_b200/csrc/uapic_efd_body.cuh"

namespace {

// all N samples of the particle in one thread; direct DFTs
template <int N> struct HostTau {
    static constexpr int SPL = N;
    double c_[N], s_[N];
    uapic::cd tw[N];
    HostTau() {
        const double pi = 4.0 * std::atan(1.0);
        for (int n = 0; n < N; ++n) {
            c_[n] = std::cos(2.0 * pi * n / N); s_[n] = std::sin(2.0 * pi * n / N);
            tw[n] = uapic::mk(std::cos(-2.0 * pi * n / N), std::sin(-2.0 * pi * n / N));
        }
    }
    double ct(int j) const { return c_[j]; }
    double st(int j) const { return s_[j]; }
    bool mode_live(int) const { return true; }
    double lmode(int k) const { return (double)(k < N / 2 ? k : k - N); }
    void dft(uapic::cd (&a)[N], bool forward) const {
        uapic::cd out[N];
        for (int k = 0; k < N; ++k) {
            uapic::cd acc = uapic::mk(0.0, 0.0);
            for (int n = 0; n < N; ++n) {
                uapic::cd w = tw[(k * n) % N];
                if (!forward) w.im = -w.im;
                const uapic::cd t = uapic::cmul(a[n], w);
                acc = uapic::mk(acc.re + t.re, acc.im + t.im);
            }
            out[k] = forward ? uapic::mk(acc.re / N, acc.im / N) : acc;
        }
        for (int k = 0; k < N; ++k) a[k] = out[k];
    }
    void fwd(uapic::cd (&a)[N]) const { dft(a, true); }
    void inv(uapic::cd (&a)[N]) const { dft(a, false); }
    void fwd2(uapic::cd (&a)[N], uapic::cd (&b)[N]) const { dft(a, true); dft(b, true); }
    void inv2(uapic::cd (&a)[N], uapic::cd (&b)[N]) const { dft(a, false); dft(b, false); }
    uapic::cd first(const uapic::cd (&a)[N]) const { return a[0]; }
    uapic::cd sum(uapic::cd v) const { return v; }
};

template <int N> void run(const uapic::EfdScalars &q, int64_t np, double *x, double *v) {
    HostTau<N> T;
    for (int64_t m = 0; m < np; ++m) {
        double xo[2], vo[2];
        uapic::efd_particle<HostTau<N>>(T, q, x[2 * m], x[2 * m + 1], v[2 * m], v[2 * m + 1], xo, vo);
        x[2 * m] = xo[0]; x[2 * m + 1] = xo[1]; v[2 * m] = vo[0]; v[2 * m + 1] = vo[1];
    }
}

}  // namespace

extern "C" int efd_host_body(int ntau, int64_t np, double eps, double dt, double tfinal, int nstep, const double *box, double *x, double *v) {
    uapic::EfdScalars q;
    q.eps = eps; q.dt = dt; q.tfinal = tfinal; q.nstep = nstep;
    q.xmin = box[0]; q.xmax = box[1]; q.ymin = box[2]; q.ymax = box[3];
    switch (ntau) {
        case 4: run<4>(q, np, x, v); return 0;
        case 12: run<12>(q, np, x, v); return 0;
        case 16: run<16>(q, np, x, v); return 0;
        case 32: run<32>(q, np, x, v); return 0;
        default: return -1;
    }
}

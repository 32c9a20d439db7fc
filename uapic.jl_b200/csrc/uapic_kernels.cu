// uapic_kernels.cu -- all CUDA kernels of libuapic_b200 (sm_100a).
//
//  1. stage kernels: one per reference routine, reference-shaped arrays (parity + the Julia stage API)
//  2. mesh kernels : rho epilogue, Poisson (shared-memory FFTs), energy
//  3. fused phase kernels: the session (performance) path
//  4. loaders / diagnostics
//
// Reference citations are to /root/reference (fortran/*.F90, src/*.jl); see DESIGN.md for the mapping.
#include <cooperative_groups.h>

#include "uapic_internal.h"
#include "uapic_mesh.cuh"

namespace uapic {

namespace {

constexpr int kBlock = 256;

inline int grid_for(const LaunchCtx &c, int64_t work_items, int items_per_block, int max_blocks_per_sm = 8) {
    int64_t need = (work_items + items_per_block - 1) / items_per_block;
    int64_t cap = (int64_t)c.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
inline void count(const LaunchCtx &c, int n = 1) { if (c.launches) *c.launches += n; }

#define UAPIC_DISPATCH_N(ntau, CALL)                      \
    switch (ntau) {                                       \
        case 2:  { constexpr int N = 2;  CALL; } break;   \
        case 4:  { constexpr int N = 4;  CALL; } break;   \
        case 8:  { constexpr int N = 8;  CALL; } break;   \
        case 16: { constexpr int N = 16; CALL; } break;   \
        case 32: { constexpr int N = 32; CALL; } break;   \
        default: return cudaErrorInvalidValue;            \
    }

// =================================================================================================
// 1. stage kernels
// =================================================================================================

// preparation                                         ua_steps.F90:15-115 / src/ua_steps.jl:3-78
template <int N>
__global__ void __launch_bounds__(kBlock) k_preparation(double eps, double dt, int64_t np, const double *x, const double *v,
                                                        const double *e, double *b, double *t, double *pl, double *ql,
                                                        double *xt, double *yt) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    for (int64_t base = W.first; base < np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < np;
        const int64_t k = valid ? kraw : np - 1;
        Pcl p;
        const double2 xx = ld2(x, k), vv = ld2(v, k), ee = ld2(e, k);
        p.x1 = xx.x; p.x2 = xx.y; p.vx = vv.x; p.vy = vv.y; p.ex = ee.x; p.ey = ee.y;
        double xt1, xt2, interv; cd yt1, yt2;
        prep_particle<N>(L, eps, dt, p, xt1, xt2, yt1, yt2, interv);
        const cd elt = elt_minus<N>(L, p.t, eps);
        cd plv, qlv;
        pl_ql<N>(L, p.t, eps, elt, plv, qlv);
        if (valid) {
            if (L.j == 0) { b[k] = p.b; t[k] = p.t; }
            st2(pl, L.k + (int64_t)N * k, plv);
            st2(ql, L.k + (int64_t)N * k, qlv);
            st2(xt, L.j + (int64_t)N * (0 + 2 * k), mk(xt1, 0.0));
            st2(xt, L.j + (int64_t)N * (1 + 2 * k), mk(xt2, 0.0));
            st2(yt, L.j + (int64_t)N * (0 + 2 * k), yt1);
            st2(yt, L.j + (int64_t)N * (1 + 2 * k), yt2);
        }
    }
}

// interpolate_eb_m6_complex                           interpolation_m6.F90:40-191 / src/interpolation.jl:3-123
__global__ void __launch_bounds__(kBlock) k_gather_tau(MeshDev m, const double2 *__restrict__ emesh, int ntau, int64_t np,
                                                       const double *__restrict__ xt, double *__restrict__ et, int wrap) {
    const int64_t total = (int64_t)ntau * np;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = s / ntau;
        const int n = (int)(s - k * ntau);
        const int64_t i1 = n + (int64_t)ntau * (0 + 2 * k), i2 = n + (int64_t)ntau * (1 + 2 * k);
        double xw, yw;
        const Cell c = m6_cell_exact(m, xt[2 * i1], xt[2 * i2], wrap, xw, yw);   // real(x(n,1,k)), real(x(n,2,k))
        double e1, e2;
        m6_gather_exact(m, emesh, c, e1, e2);
        et[i1] = e1; et[i2] = e2;
    }
}

// compute_f                                           ua_steps.F90:140-198 / src/ua_steps.jl:105-145
template <int N>
__global__ void __launch_bounds__(kBlock) k_compute_f(double eps, int64_t np, const double *b, const double *xt, const double *yt,
                                                      const double *et, double *fx, double *fy, int normalise) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    const double sc = normalise ? 1.0 / (double)N : 1.0;
    for (int64_t base = W.first; base < np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < np;
        const int64_t k = valid ? kraw : np - 1;
        const int64_t i1 = L.j + (int64_t)N * (0 + 2 * k), i2 = L.j + (int64_t)N * (1 + 2 * k);
        const double bb = b[k];
        const double x1 = xt[2 * i1], x2 = xt[2 * i2];
        const cd y1 = ldc(yt, i1), y2 = ldc(yt, i2);
        const double interv = (1.0 + 0.5 * sin(x1) * sin(x2) - bb) / eps;                 // :177
        cd f1, f2, g1, g2;
        force_terms<N>(L, 1.0 / bb, interv, y1, y2, et[i1], et[i2], f1, f2, g1, g2);
        f1 = rmul(sc, fft_fwd<N>(f1, L)); f2 = rmul(sc, fft_fwd<N>(f2, L));               // :187-195
        g1 = rmul(sc, fft_fwd<N>(g1, L)); g2 = rmul(sc, fft_fwd<N>(g2, L));
        if (valid) {
            const int64_t o1 = L.k + (int64_t)N * (0 + 2 * k), o2 = L.k + (int64_t)N * (1 + 2 * k);
            st2(fx, o1, f1); st2(fx, o2, f2); st2(fy, o1, g1); st2(fy, o2, g2);
        }
    }
}

// mul!(x̃t, ftau, xt) / ifft!(xt,1)                    test/bupdate.jl:79,85
template <int N>
__global__ void __launch_bounds__(kBlock) k_fft_tau(int64_t nvec, const double *in, double *out, int sign, int normalise) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    const double sc = normalise ? 1.0 / (double)N : 1.0;
    for (int64_t base = W.first; base < nvec; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < nvec;
        const int64_t k = valid ? kraw : nvec - 1;
        if (sign < 0) {
            cd z = rmul(sc, fft_fwd<N>(ldc(in, L.j + (int64_t)N * k), L));
            if (valid) st2(out, L.k + (int64_t)N * k, z);
        } else {
            cd z = rmul(sc, fft_bwd<N>(ldc(in, L.k + (int64_t)N * k), L));
            if (valid) st2(out, L.j + (int64_t)N * k, z);
        }
    }
}

// ua_step! (Julia, pointwise in tau-Fourier space)     src/ua_steps.jl:149-200 ; gx == null -> predictor
__global__ void __launch_bounds__(kBlock) k_step_pointwise(int ntau, double eps, int64_t np, const double *t, const double *pl,
                                                           const double *ql, const double *xf, const double *fx,
                                                           const double *gx, double *out) {
    const int64_t total = (int64_t)ntau * np;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = s / ntau;
        const int n = (int)(s - k * ntau);
        const double l = (n < ntau / 2) ? (double)n : (double)(n - ntau);
        const double tt = t[k];
        double sn, cs;
        sincos(-(l * tt) / eps, &sn, &cs);                                               // :162, :185
        const cd elt = mk(cs, sn);
        const cd plv = ldc(pl, n + (int64_t)ntau * k);
        for (int c = 0; c < 2; ++c) {
            const int64_t i = n + (int64_t)ntau * (c + 2 * k);
            const cd f = ldc(fx, i);
            cd r = cfma(plv, f, cmul(elt, ldc(xf, i)));                                  // :163-164
            if (gx) {
                const cd qlv = ldc(ql, n + (int64_t)ntau * k);
                const cd d = csub(ldc(gx, i), f);
                const cd q = cmul(qlv, d);
                r = mk(r.re + q.re / tt, r.im + q.im / tt);                              // :190-191
            }
            st2(out, i, r);
        }
    }
}

// ua_step1 / ua_step2 (Fortran forms)                  ua_steps.F90:200-272
template <int N>
__global__ void __launch_bounds__(kBlock) k_step_fortran(double eps, int64_t np, const double *t, const double *pl, const double *ql,
                                                         double *xt, double *xf, const double *fx, const double *gx,
                                                         int corrector) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    for (int64_t base = W.first; base < np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < np;
        const int64_t k = valid ? kraw : np - 1;
        const double tt = t[k];
        const cd elt = rmul(1.0 / (double)N, elt_minus<N>(L, tt, eps));                  // :224-225, :258-259
        const cd plv = ldc(pl, L.k + (int64_t)N * k);
        cd out[2];
        for (int c = 0; c < 2; ++c) {
            const int64_t it = L.j + (int64_t)N * (c + 2 * k), is = L.k + (int64_t)N * (c + 2 * k);
            cd xfv;
            if (!corrector) {
                xfv = fft_fwd<N>(ldc(xt, it), L);                                        // :217-218
                if (valid) st2(xf, is, xfv);
            } else {
                xfv = ldc(xf, is);
            }
            const cd f = ldc(fx, is);
            cd r = cfma(plv, f, cmul(elt, xfv));                                         // :226, :260
            if (corrector) {
                const cd q = cmul(ldc(ql, L.k + (int64_t)N * k), csub(ldc(gx, is), f));
                r = mk(r.re + q.re / tt, r.im + q.im / tt);                              // :261
            }
            out[c] = fft_bwd<N>(r, L);                                                   // :231-232, :267-268
        }
        if (valid) {
            st2(xt, L.j + (int64_t)N * (0 + 2 * k), out[0]);
            st2(xt, L.j + (int64_t)N * (1 + 2 * k), out[1]);
        }
    }
}

// compute_rho_m6_complex (per-particle part)           compute_rho_m6.F90:70-187 / src/compute_rho.jl:43-165
template <int N>
__global__ void __launch_bounds__(kBlock) k_deposit_tau(MeshDev m, double eps, int64_t np, const double *xt, const double *t, double w,
                                                        RhoAcc acc, double *x, int wrap) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    for (int64_t base = W.first; base < np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < np;
        const int64_t k = valid ? kraw : np - 1;
        const double tt = t[k];
        const cd elt = elt_minus<N>(L, tt, eps);
        const double sc = 1.0 / (double)N;
        const cd f1 = rmul(sc, fft_fwd<N>(ldc(xt, L.j + (int64_t)N * (0 + 2 * k)), L));   // :74
        const cd f2 = rmul(sc, fft_fwd<N>(ldc(xt, L.j + (int64_t)N * (1 + 2 * k)), L));   // :80
        const double p1 = eval_tau_star<N>(f1, elt).re;                                  // :76-78
        const double p2 = eval_tau_star<N>(f2, elt).re;                                  // :82-84
        double xw, yw;
        const Cell c = m6_cell_exact(m, p1, p2, wrap, xw, yw);
        if (valid) {
            if (L.j == 0) reinterpret_cast<double2 *>(x)[k] = make_double2(xw, yw);      // :86-87
            m6_scatter(m, acc, c, w, L.j, N);
        }
    }
}

// compute_v                                            ua_steps.F90:274-307 / src/ua_steps.jl:204-224
template <int N>
__global__ void __launch_bounds__(kBlock) k_compute_v(double eps, int64_t np, const double *t, const double *yt, int yt_is_fourier,
                                                      double *v) {
    TauLane<N> L; L.init(threadIdx.x & 31);
    WarpMap<N> W;
    for (int64_t base = W.first; base < np; base += W.stride) {
        const int64_t kraw = base + W.g;
        const bool valid = kraw < np;
        const int64_t k = valid ? kraw : np - 1;
        const double tt = t[k];
        const cd elt = elt_minus<N>(L, tt, eps);
        cd y1, y2;
        if (yt_is_fourier) {
            y1 = ldc(yt, L.k + (int64_t)N * (0 + 2 * k));
            y2 = ldc(yt, L.k + (int64_t)N * (1 + 2 * k));
        } else {
            y1 = fft_fwd<N>(ldc(yt, L.j + (int64_t)N * (0 + 2 * k)), L);                 // :290
            y2 = fft_fwd<N>(ldc(yt, L.j + (int64_t)N * (1 + 2 * k)), L);                 // :291
        }
        const double sc = 1.0 / (double)N;
        const double px = eval_tau_star<N>(rmul(sc, y1), elt).re;                        // :296-300
        const double py = eval_tau_star<N>(rmul(sc, y2), elt).re;
        double sn, cs;
        sincos(tt / eps, &sn, &cs);
        if (valid && L.j == 0)
            reinterpret_cast<double2 *>(v)[k] = make_double2(cs * px + sn * py, cs * py - sn * px);   // :302-303
    }
}

// compute_rho_m6_real (per-particle part)              compute_rho_m6.F90:221-327 / src/compute_rho.jl:195-300
__global__ void __launch_bounds__(kBlock) k_deposit(MeshDev m, int64_t np, double *x, double w, RhoAcc acc, int wrap, int scheme) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 xx = ld2(x, k);
        double xw, yw;
        const Cell c = m6_cell_exact(m, xx.x, xx.y, wrap, xw, yw);
        if (wrap == kWrapJulia) reinterpret_cast<double2 *>(x)[k] = make_double2(xw, yw);
        if (scheme == 1) {      // build-defined CIC: weights of performance/test_cic.F90:73-76, products in the oracle's order
            for (int q = 0; q < 4; ++q) {
                const int ox = q >> 1, oy = q & 1;
                const double wx = ox ? c.dpx : __dsub_rn(1.0, c.dpx), wy = oy ? c.dpy : __dsub_rn(1.0, c.dpy);
                rho_add(acc, wrap_index(c.i, ox, m.nx) + wrap_index(c.j, oy, m.ny) * m.ld, __dmul_rn(__dmul_rn(wx, wy), w));
            }
        } else {
            m6_scatter(m, acc, c, w, 0, 1);
        }
    }
}

// interpolate_eb_m6_real                               interpolation_m6.F90:193-327 / src/interpolation.jl:125-247
__global__ void __launch_bounds__(kBlock) k_gather(MeshDev m, const double2 *__restrict__ emesh, int64_t np, double *x, double *ep,
                                                   int wrap, int scheme) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 xx = ld2(x, k);
        double xw, yw;
        const Cell c = m6_cell_exact(m, xx.x, xx.y, wrap, xw, yw);
        if (wrap == kWrapJulia) reinterpret_cast<double2 *>(x)[k] = make_double2(xw, yw);
        double e1, e2;
        if (scheme == 1) {      // build-defined CIC, summed in the oracle's order: x offset outer, y offset inner
            e1 = 0.0; e2 = 0.0;
            for (int q = 0; q < 4; ++q) {
                const int ox = q >> 1, oy = q & 1;
                const double wx = ox ? c.dpx : __dsub_rn(1.0, c.dpx), wy = oy ? c.dpy : __dsub_rn(1.0, c.dpy);
                const double2 ev = emesh[wrap_index(c.i, ox, m.nx) + wrap_index(c.j, oy, m.ny) * m.ld];
                const double wq = __dmul_rn(wx, wy);
                e1 = __dadd_rn(e1, __dmul_rn(wq, ev.x));
                e2 = __dadd_rn(e2, __dmul_rn(wq, ev.y));
            }
        } else
        m6_gather_exact(m, emesh, c, e1, e2);
        reinterpret_cast<double2 *>(ep)[k] = make_double2(e1, e2);
    }
}

// =================================================================================================
// 2. mesh kernels
// =================================================================================================

// compute_rho_m6.F90:191-200 : ghost copy -> /(dx*dy) -> subtract the mean.  Single CTA: the mesh is <= 0.5 MB and
// the sum must have a fixed order.
__global__ void __launch_bounds__(kMeshBlock) k_rho_epilogue(MeshDev m, RhoAcc acc, double *rho, double *rho_total) {
    __shared__ double sh[kMeshBlock];
    const int nx = m.nx, ny = m.ny, ld = m.ld;
    const double dxdy = m.dx * m.dy;
    const double inv_scale = acc.i64 ? 1.0 / acc.scale : 1.0;
    double part = 0.0;
    long long ipart = 0;
    for (int idx = threadIdx.x; idx < nx * ny; idx += blockDim.x) {
        const int j = idx / nx, i = idx - j * nx;
        const int q = i + ld * j;
        double raw;
        if (acc.i64) { const long long r = (long long)acc.i64[q]; ipart += r; raw = (double)r * inv_scale; }
        else raw = acc.f64[q];
        const double val = raw / dxdy;
        rho[q] = val;
        part += val;
    }
    double total;
    if (acc.i64) {
        // fixed point: sum(rho)*dx*dy == sum(raw) exactly, and the integer sum has no order -> bit-reproducible
        long long *ish = reinterpret_cast<long long *>(sh);
        ish[threadIdx.x] = ipart;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) ish[threadIdx.x] += ish[threadIdx.x + s];
            __syncthreads();
        }
        total = (double)ish[0] * inv_scale;
        __syncthreads();
    } else {
        total = block_sum(part, sh) * m.dx * m.dy;
    }
    const double sub = total / m.dimx / m.dimy;
    for (int idx = threadIdx.x; idx < nx * ny; idx += blockDim.x) {
        const int j = idx / nx, i = idx - j * nx;
        rho[i + ld * j] -= sub;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nx; i += blockDim.x) rho[i + ld * ny] = rho[i];
    for (int j = threadIdx.x; j < ny; j += blockDim.x) rho[nx + ld * j] = rho[ld * j];
    if (threadIdx.x == 0) {
        rho[nx + ld * ny] = rho[0];
        if (rho_total) *rho_total = total;
    }
}

// r2c along x for row j                                poisson_2d.f90:96 (first half)
__global__ void k_poisson_rows_fwd(MeshDev m, const double *__restrict__ rho, double2 *__restrict__ rk) {
    extern __shared__ double2 smem_raw[];
    cd *a = reinterpret_cast<cd *>(smem_raw);
    cd *tmp = a + m.nx;
    cd *tw = tmp + m.nx;
    const int j = blockIdx.x, nx = m.nx, nh = nx / 2 + 1;
    line_twiddles(tw, nx);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) a[i] = mk(rho[i + m.ld * j], 0.0);
    line_fft(a, tmp, tw, nx, -1);
    for (int i = threadIdx.x; i < nh; i += blockDim.x) rk[i + (size_t)nh * j] = make_double2(a[i].re, a[i].im);
}

// along y for column ik: forward, multiply by -i*k/k^2, backward for both components     poisson_2d.f90:96-102
__global__ void k_poisson_cols(MeshDev m, const double2 *__restrict__ rk, double2 *__restrict__ ek) {
    extern __shared__ double2 smem_raw[];
    const int ny = m.ny, nx = m.nx, nh = nx / 2 + 1;
    cd *a = reinterpret_cast<cd *>(smem_raw);
    cd *bx = a + ny;
    cd *by = bx + ny;
    cd *tmp = by + ny;
    cd *tw = tmp + ny;
    const int ik = blockIdx.x;
    const double pi = 3.14159265358979323846;
    const double kx0 = 2.0 * pi / m.dimx, ky0 = 2.0 * pi / m.dimy;                        // :48-49
    line_twiddles(tw, ny);
    for (int jk = threadIdx.x; jk < ny; jk += blockDim.x) {
        const double2 r = rk[ik + (size_t)nh * jk];
        a[jk] = mk(r.x, r.y);
    }
    line_fft(a, tmp, tw, ny, -1);
    for (int jk = threadIdx.x; jk < ny; jk += blockDim.x) {
        double kx = (double)ik * kx0;                                                    // :59
        const double ky = (jk < ny / 2) ? (double)jk * ky0 : (double)(jk - ny) * ky0;    // :62, :66
        if (ik == 0 && jk == 0) kx = 1.0;                                                // :70
        const double k2 = kx * kx + ky * ky;                                             // :71
        const double kkx = kx / k2, kky = ky / k2;                                       // :72-73
        const cd r = a[jk];
        bx[jk] = mk(kkx * r.im, -kkx * r.re);                                            // (0,-1)*kx*rho   :98
        by[jk] = mk(kky * r.im, -kky * r.re);                                            // :99
    }
    line_fft(bx, tmp, tw, ny, +1);
    line_fft(by, tmp, tw, ny, +1);
    const size_t plane = (size_t)nh * ny;
    for (int jk = threadIdx.x; jk < ny; jk += blockDim.x) {
        ek[ik + (size_t)nh * jk] = make_double2(bx[jk].re, bx[jk].im);
        ek[plane + ik + (size_t)nh * jk] = make_double2(by[jk].re, by[jk].im);
    }
}

// c2r along x for row j, component blockIdx.y; FFTW c2r semantics: Im of the DC / Nyquist bins ignored.
// Also writes the ghost column, the ghost row (from row 0) and the 1/(nx*ny) scale      poisson_2d.f90:101-109
__global__ void k_poisson_rows_bwd(MeshDev m, const double2 *__restrict__ ek, double *__restrict__ emesh) {
    extern __shared__ double2 smem_raw[];
    cd *a = reinterpret_cast<cd *>(smem_raw);
    cd *tmp = a + m.nx;
    cd *tw = tmp + m.nx;
    const int j = blockIdx.x, comp = blockIdx.y, nx = m.nx, ny = m.ny, nh = nx / 2 + 1;
    const double2 *src = ek + (size_t)comp * nh * ny + (size_t)nh * j;
    line_twiddles(tw, nx);
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
        cd z;
        if (i < nh) {
            z = mk(src[i].x, src[i].y);
            if (i == 0 || (2 * i == nx)) z.im = 0.0;
        } else {
            z = mk(src[nx - i].x, -src[nx - i].y);
        }
        a[i] = z;
    }
    line_fft(a, tmp, tw, nx, +1);
    const double sc = 1.0 / (double)(nx * ny);
    for (int i = threadIdx.x; i <= nx; i += blockDim.x) {
        const double val = a[i == nx ? 0 : i].re * sc;
        emesh[comp + 2 * ((size_t)i + (size_t)m.ld * j)] = val;
        if (j == 0) emesh[comp + 2 * ((size_t)i + (size_t)m.ld * ny)] = val;
    }
}

// periodic halo copy of E: lets the fused gathers address the 6x6 stencil without any index wrap
// (the ghost row/column of the reference layout already is the i = nx / j = ny image; this extends it to [-2, n+3])
__global__ void __launch_bounds__(kBlock) k_extend_emesh(MeshDev m, const double2 *__restrict__ emesh, double2 *__restrict__ ehalo) {
    const int lx = m.nx + 6, ly = m.ny + 6;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < lx * ly; q += gridDim.x * blockDim.x) {
        const int je = q / lx, ie = q - je * lx;
        int i = ie - 2, j = je - 2;
        i += (i < 0) ? m.nx : 0; i -= (i >= m.nx) ? m.nx : 0;
        j += (j < 0) ? m.ny : 0; j -= (j >= m.ny) ? m.ny : 0;
        ehalo[q] = emesh[i + m.ld * j];
    }
}

// the same halo copy in the 2 x 4-node tiled layout of uapic_fast.cuh (gather_tiled)
__global__ void __launch_bounds__(kBlock) k_extend_emesh_tiled(MeshDev m, const double2 *__restrict__ emesh, double2 *__restrict__ ehalo) {
    const int ntx = (m.nx + 6 + 1) >> 1, nty = (m.ny + 6 + 3) >> 2;
    const int n = ntx * nty * 8;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const int tile = q >> 3, w = q & 7;
        const int tj = tile / ntx, ti = tile - tj * ntx;
        int i = 2 * ti + (w & 1) - 2, j = 4 * tj + (w >> 1) - 2;
        // nodes of the padding beyond [-2, n+3] wrap as well: never read, but defined
        i %= m.nx; i += (i < 0) ? m.nx : 0;
        j %= m.ny; j += (j < 0) ? m.ny : 0;
        ehalo[q] = emesh[i + m.ld * j];
    }
}

// src/poisson.jl:80-81 : sum over the ghosted array of e1^2+e2^2, times dx*dy (fixed summation order)
__global__ void __launch_bounds__(kMeshBlock) k_energy(MeshDev m, const double2 *__restrict__ emesh, double *energy) {
    __shared__ double sh[kMeshBlock];
    const int n = m.ld * (m.ny + 1);
    double part = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const double2 ev = emesh[q];
        part += ev.x * ev.x + ev.y * ev.y;
    }
    const double tot = block_sum(part, sh);
    if (threadIdx.x == 0) *energy = tot * m.dx * m.dy;
}


// =================================================================================================
// 2b. the whole field solve of the session path in ONE cooperative launch, for up to two meshes at once
// =================================================================================================
// raw deposits (already summed over the ranks) -> fold of the CTA-private copies -> /(dx dy), mean, ghosts
// (compute_rho_m6.F90:191-200) -> r2c along x, FFT along y, -i k/k^2, two inverse FFTs, c2r along x (poisson_2d.f90:85-111)
// -> ghost copies, 1/(nx ny) -> electric energy (src/poisson.jl:80-81) -> tiled / linear halo copy for the gathers.
// The separate kernels above did this in 6 launches per mesh, two of them single-CTA (9.7 + 5.2 us); a one-pass step needs
// two solves back to back, i.e. 12 launches and ~150 us of fixed cost -- 7 % of a BASELINE config-2 step.  Here the phases
// are separated by grid.sync() and both meshes of a step go through together.  All sums run over kSolveParts FIXED chunks in
// a fixed order, so the result does not depend on the grid size (same bits on any GPU, any SM count).
constexpr int kSolveParts = 64;
constexpr int kSolveBlock = 256;

DEVINL double block_sum_256(double v, double *sh) {      // fixed tree over kSolveBlock threads, valid on every thread
    const int tid = threadIdx.x;
    __syncthreads();
    sh[tid] = v;
    __syncthreads();
    for (int s = kSolveBlock / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    return sh[0];
}
DEVINL long long block_isum_256(long long v, long long *sh) {
    const int tid = threadIdx.x;
    __syncthreads();
    sh[tid] = v;
    __syncthreads();
    for (int s = kSolveBlock / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    return sh[0];
}

__global__ void __launch_bounds__(kSolveBlock) k_field_solve(MeshDev m, SolveBatch B) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double2 smem_raw[];
    __shared__ double shd[kSolveBlock];
    long long *shi = reinterpret_cast<long long *>(shd);
    cd *buf = reinterpret_cast<cd *>(smem_raw);
    const int nx = m.nx, ny = m.ny, ld = m.ld, nh = nx / 2 + 1, nb = B.nb;
    const double dxdy = m.dx * m.dy;
    const int ncell = nx * ny, per = (ncell + kSolveParts - 1) / kSolveParts;

    // ---- peer mode: wait until every rank has published exchange number B.peer_seq (its folded deposits sit in ITS exchange
    //      buffer, mapped here over NVLink), then phase 0 sums the ranks' buffers in rank order -- the all-reduce lives inside the
    //      solve, every rank adds in the same order, so the sum is the same bits everywhere (fp64 included)
    if (B.npeers > 0) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            for (int r = 0; r < B.npeers; ++r) {
                unsigned long long seen;
                do {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(B.peer_flag[r]) : "memory");
                    if (seen < B.peer_seq && clock64() - t0 > (1ll << 33)) { atomicExch(B.peer_error, 1); break; }   // ~4 s: a rank is gone
                } while (seen < B.peer_seq);
            }
        }
        __syncthreads();
    }
    // ---- phase 0: fold the copies, scale, partial sums over fixed chunks ----
    for (int task = blockIdx.x; task < nb * kSolveParts; task += gridDim.x) {
        const int b = task / kSolveParts, c = task - b * kSolveParts;
        const RhoAcc acc = B.acc[b];
        const double inv_scale = acc.i64 ? 1.0 / acc.scale : 1.0;
        const int lo = c * per, hi = min(lo + per, ncell);
        double part = 0.0;
        long long ipart = 0;
        for (int idx = lo + threadIdx.x; idx < hi; idx += kSolveBlock) {
            const int j = idx / nx, i = idx - j * nx, q = i + ld * j;
            double raw;
            if (B.npeers > 0) {
                const size_t off = (size_t)b * B.peer_mesh_stride + q;
                if (acc.i64) {
                    unsigned long long t = 0;
                    for (int r = 0; r < B.npeers; ++r) {
                        unsigned long long u;
                        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(u) : "l"(B.peer_data[r] + off));
                        t += u;
                    }
                    ipart += (long long)t;
                    raw = (double)(long long)t * inv_scale;
                } else {
                    double t = 0.0;
                    for (int r = 0; r < B.npeers; ++r) {
                        double u;
                        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(u) : "l"(B.peer_data[r] + off));
                        t += u;
                    }
                    raw = t;
                }
            } else if (acc.i64) {
                unsigned long long t = acc.i64[q];
                for (int k = 1; k < B.fold_copies; ++k) t += acc.i64[(size_t)k * B.fold_stride + q];
                ipart += (long long)t;
                raw = (double)(long long)t * inv_scale;
            } else {
                double t = acc.f64[q];
                for (int k = 1; k < B.fold_copies; ++k) t += acc.f64[(size_t)k * B.fold_stride + q];
                raw = t;
            }
            const double val = raw / dxdy;
            B.rho[b][q] = val;
            part += val;
        }
        if (acc.i64) {
            const long long tot = block_isum_256(ipart, shi);
            if (threadIdx.x == 0) reinterpret_cast<long long *>(B.partial)[task] = tot;
        } else {
            const double tot = block_sum_256(part, shd);
            if (threadIdx.x == 0) B.partial[task] = tot;
        }
    }
    grid.sync();

    // ---- phase 1: subtract the mean, ghosts of rho, r2c along x ----
    {
        cd *a = buf, *tmp = a + nx, *tw = tmp + nx;
        line_twiddles(tw, nx);
        double sub[2] = {0.0, 0.0};
        for (int b = 0; b < nb; ++b) {
            double total;
            if (B.acc[b].i64) {
                long long t = 0;
                for (int c = 0; c < kSolveParts; ++c) t += reinterpret_cast<const long long *>(B.partial)[b * kSolveParts + c];
                total = (double)t / B.acc[b].scale;                // sum(rho)*dx*dy == sum(raw) exactly in fixed point
            } else {
                double t = 0.0;
                for (int c = 0; c < kSolveParts; ++c) t += B.partial[b * kSolveParts + c];
                total = t * m.dx * m.dy;
            }
            sub[b] = total / m.dimx / m.dimy;
        }
        for (int task = blockIdx.x; task < nb * ny; task += gridDim.x) {
            const int b = task / ny, j = task - b * ny;
            double *rho = B.rho[b];
            __syncthreads();
            for (int i = threadIdx.x; i < nx; i += kSolveBlock) {
                const double val = rho[i + ld * j] - sub[b];
                rho[i + ld * j] = val;
                if (j == 0) rho[i + ld * ny] = val;                 // ghost row
                if (i == 0) { rho[nx + ld * j] = val; if (j == 0) rho[nx + ld * ny] = val; }   // ghost column, corner
                a[i] = mk(val, 0.0);
            }
            line_fft(a, tmp, tw, nx, -1);
            for (int i = threadIdx.x; i < nh; i += kSolveBlock) B.rk[b][i + (size_t)nh * j] = make_double2(a[i].re, a[i].im);
        }
    }
    grid.sync();

    // ---- phase 2: along y, multiply by -i k / k^2, back along y for both components ----
    {
        cd *a = buf, *bx = a + ny, *by = bx + ny, *tmp = by + ny, *tw = tmp + ny;
        const double pi = 3.14159265358979323846;
        const double kx0 = 2.0 * pi / m.dimx, ky0 = 2.0 * pi / m.dimy;
        __syncthreads();
        line_twiddles(tw, ny);
        for (int task = blockIdx.x; task < nb * nh; task += gridDim.x) {
            const int b = task / nh, ik = task - b * nh;
            __syncthreads();
            for (int jk = threadIdx.x; jk < ny; jk += kSolveBlock) {
                const double2 r = B.rk[b][ik + (size_t)nh * jk];
                a[jk] = mk(r.x, r.y);
            }
            line_fft(a, tmp, tw, ny, -1);
            for (int jk = threadIdx.x; jk < ny; jk += kSolveBlock) {
                double kx = (double)ik * kx0;
                const double ky = (jk < ny / 2) ? (double)jk * ky0 : (double)(jk - ny) * ky0;
                if (ik == 0 && jk == 0) kx = 1.0;
                const double k2 = kx * kx + ky * ky;
                const double kkx = kx / k2, kky = ky / k2;
                const cd r = a[jk];
                bx[jk] = mk(kkx * r.im, -kkx * r.re);
                by[jk] = mk(kky * r.im, -kky * r.re);
            }
            line_fft(bx, tmp, tw, ny, +1);
            line_fft(by, tmp, tw, ny, +1);
            const size_t plane = (size_t)nh * ny;
            for (int jk = threadIdx.x; jk < ny; jk += kSolveBlock) {
                B.ek[b][ik + (size_t)nh * jk] = make_double2(bx[jk].re, bx[jk].im);
                B.ek[b][plane + ik + (size_t)nh * jk] = make_double2(by[jk].re, by[jk].im);
            }
        }
    }
    grid.sync();

    // ---- phase 3: c2r along x (FFTW semantics), ghosts, 1/(nx ny) ----
    {
        cd *a = buf, *tmp = a + nx, *tw = tmp + nx;
        __syncthreads();
        line_twiddles(tw, nx);
        const double sc = 1.0 / (double)(nx * ny);
        for (int task = blockIdx.x; task < nb * 2 * ny; task += gridDim.x) {
            const int b = task / (2 * ny), r = task - b * 2 * ny, comp = r / ny, j = r - comp * ny;
            const double2 *src = B.ek[b] + (size_t)comp * nh * ny + (size_t)nh * j;
            double *emesh = reinterpret_cast<double *>(B.emesh[b]);
            __syncthreads();
            for (int i = threadIdx.x; i < nx; i += kSolveBlock) {
                cd z;
                if (i < nh) {
                    z = mk(src[i].x, src[i].y);
                    if (i == 0 || (2 * i == nx)) z.im = 0.0;
                } else {
                    z = mk(src[nx - i].x, -src[nx - i].y);
                }
                a[i] = z;
            }
            line_fft(a, tmp, tw, nx, +1);
            for (int i = threadIdx.x; i <= nx; i += kSolveBlock) {
                const double val = a[i == nx ? 0 : i].re * sc;
                emesh[comp + 2 * ((size_t)i + (size_t)ld * j)] = val;
                if (j == 0) emesh[comp + 2 * ((size_t)i + (size_t)ld * ny)] = val;
            }
        }
    }
    grid.sync();

    // ---- phase 4: energy partials over fixed chunks; halo copies ----
    {
        const int n = ld * (ny + 1), eper = (n + kSolveParts - 1) / kSolveParts;
        for (int task = blockIdx.x; task < nb * kSolveParts; task += gridDim.x) {
            const int b = task / kSolveParts, c = task - b * kSolveParts;
            const int lo = c * eper, hi = min(lo + eper, n);
            double part = 0.0;
            for (int q = lo + threadIdx.x; q < hi; q += kSolveBlock) {
                const double2 ev = B.emesh[b][q];
                part += ev.x * ev.x + ev.y * ev.y;
            }
            const double tot = block_sum_256(part, shd);
            if (threadIdx.x == 0) B.partial[2 * kSolveParts + task] = tot;
        }
        for (int b = 0; b < nb; ++b) {
            const double2 *emesh = B.emesh[b];
            double2 *ehalo = B.ehalo[b];
            if (!ehalo) continue;
            if (B.halo_tiled) {
                const int ntx = (nx + 6 + 1) >> 1, nty = (ny + 6 + 3) >> 2, nn = ntx * nty * 8;
                for (int q = blockIdx.x * kSolveBlock + threadIdx.x; q < nn; q += gridDim.x * kSolveBlock) {
                    const int tile = q >> 3, w = q & 7;
                    const int tj = tile / ntx, ti = tile - tj * ntx;
                    int i = 2 * ti + (w & 1) - 2, j = 4 * tj + (w >> 1) - 2;
                    i %= nx; i += (i < 0) ? nx : 0;
                    j %= ny; j += (j < 0) ? ny : 0;
                    ehalo[q] = emesh[i + ld * j];
                }
            } else {
                const int lx = nx + 6, ly = ny + 6;
                for (int q = blockIdx.x * kSolveBlock + threadIdx.x; q < lx * ly; q += gridDim.x * kSolveBlock) {
                    const int je = q / lx, ie = q - je * lx;
                    int i = ie - 2, j = je - 2;
                    i += (i < 0) ? nx : 0; i -= (i >= nx) ? nx : 0;
                    j += (j < 0) ? ny : 0; j -= (j >= ny) ? ny : 0;
                    ehalo[q] = emesh[i + ld * j];
                }
            }
        }
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < nb) {
        const int b = threadIdx.x;
        double t = 0.0;
        for (int c = 0; c < kSolveParts; ++c) t += B.partial[2 * kSolveParts + b * kSolveParts + c];
        if (B.energy[b]) *B.energy[b] = t * m.dx * m.dy;
    }
}

// =================================================================================================
// 4. loaders / diagnostics
// =================================================================================================

DEVINL uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// counter-based uniform in [0,1): keyed by (seed, particle, stream, draw)
DEVINL double uniform01(uint64_t seed, uint64_t particle, uint32_t stream, uint32_t draw) {
    uint64_t h = splitmix64(seed ^ splitmix64(particle * 0xD1342543DE82EF95ull + stream));
    h = splitmix64(h + draw);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

// kind 0: rejection sampling of particles.F90:68-103 / src/plasma.jl:17-48 ; kind 1: Landau (src/landau.jl:19-43 intent)
__global__ void __launch_bounds__(kBlock) k_generate(MeshDev m, int kind, uint64_t seed, int64_t first, int64_t stride, int64_t np,
                                                     int64_t np_global, double alpha, double kx, double *x, double *v) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = (uint64_t)(first + k * stride);
        double x1, x2, v1, v2;
        if (kind == 0) {
            uint32_t d = 0;
            for (;;) {
                const double xi = uniform01(seed, id, 0, d) * m.dimx;
                const double yi = uniform01(seed, id, 0, d + 1) * m.dimy;
                const double zi = (2.0 + alpha) * uniform01(seed, id, 0, d + 2);
                d += 3;
                if (1.0 + sin(yi) + alpha * cos(kx * xi) >= zi) { x1 = xi; x2 = yi; break; }
            }
            d = 0;
            for (;;) {
                const double xi = (uniform01(seed, id, 1, d) - 0.5) * 10.0;
                const double yi = (uniform01(seed, id, 1, d + 1) - 0.5) * 10.0;
                const double zi = uniform01(seed, id, 1, d + 2);
                d += 3;
                const double temm = (exp(-((xi - 2.0) * (xi - 2.0) + yi * yi) / 2.0) + exp(-((xi + 2.0) * (xi + 2.0) + yi * yi) / 2.0)) / 2.0;
                if (temm >= zi) { v1 = xi; v2 = yi; break; }
            }
            x1 += m.xmin; x2 += m.ymin;
        } else {
            const double r1 = uniform01(seed, id, 2, 0), r2 = uniform01(seed, id, 2, 1), r3 = uniform01(seed, id, 2, 2);
            // Newton on  x + alpha*sin(kx*x)/kx = r2 * 2*pi/kx          src/landau.jl:19-28
            const double target = r2 * (2.0 * 3.14159265358979323846 / kx);
            double x0 = target;
            for (int it = 0; it < 50; ++it) {
                const double pfun = x0 + alpha * sin(kx * x0) / kx;
                const double f = 1.0 + alpha * cos(kx * x0);
                const double xn = x0 - (pfun - target) / f;
                const bool done = fabs(xn - x0) <= 1e-12;
                x0 = xn;
                if (done) break;
            }
            x1 = m.xmin + x0;
            x2 = m.ymin + r3 * m.dimy;
            const double vv = sqrt(-2.0 * log(((double)id + 0.5) / (double)np_global));  // landau.jl:35
            double sn, cs;
            sincospi(2.0 * r1, &sn, &cs);
            v1 = vv * cs; v2 = vv * sn;
        }
        reinterpret_cast<double2 *>(x)[k] = make_double2(x1, x2);
        reinterpret_cast<double2 *>(v)[k] = make_double2(v1, v2);
    }
}

// sum(v[1,:]), sum(v[2,:])  (bupdate.F90:125) -- two-level reduction in a FIXED order that does not depend on the GPU:
// kSumParts CTAs each reduce one contiguous slice (per-thread strided partial, then the fixed tree of block_sum), a second
// single-CTA launch reduces the kSumParts partials the same way.  bupdate prints this number every step, so at 1e8 particles
// the old single-CTA kernel (1.6 GB through one SM) would have dominated a faithful driver loop.
constexpr int kSumParts = 1024;
__global__ void __launch_bounds__(kMeshBlock) k_sum_v_partial(int64_t np, const double2 *__restrict__ v, double2 *__restrict__ part) {
    __shared__ double sh[kMeshBlock];
    const int64_t per = (np + kSumParts - 1) / kSumParts;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < np ? lo + per : np;
    double a = 0.0, b = 0.0;
    for (int64_t k = lo + threadIdx.x; k < hi; k += blockDim.x) { const double2 vv = v[k]; a += vv.x; b += vv.y; }
    const double sa = block_sum(a, sh);
    const double sb = block_sum(b, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(sa, sb);
}
__global__ void __launch_bounds__(kMeshBlock) k_sum_v_final(const double2 *__restrict__ part, double *out2) {
    __shared__ double sh[kMeshBlock];
    const double2 p = threadIdx.x < kSumParts ? part[threadIdx.x] : make_double2(0.0, 0.0);
    const double sa = block_sum(p.x, sh);
    const double sb = block_sum(p.y, sh);
    if (threadIdx.x == 0) { out2[0] = sa; out2[1] = sb; }
}

// fp64 pipe probe: kProbeChains independent DFMA chains per thread, nothing else in the loop (uapic_probe_fp64_peak)
constexpr int kProbeChains = 8;
__global__ void __launch_bounds__(256) k_probe_dfma(int iters, double seed, double *sink) {
    double a[kProbeChains];
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i) a[i] = seed + (double)(threadIdx.x + i);
    const double m = 0.999999999, c = 1e-9 * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kProbeChains; ++i) a[i] = fma(a[i], m, c);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i) t += a[i];
    if (t == 12345.678) *sink = t;      // never true: keeps the chains alive
}

}  // namespace

// =================================================================================================
// launchers
// =================================================================================================

cudaError_t launch_preparation(const LaunchCtx &c, int ntau, double eps, double dt, int64_t np, const double *x,
                               const double *v, const double *e, double *b, double *t, double *pl, double *ql,
                               double *xt, double *yt) {
    if (np <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_preparation_generic(c, ntau, eps, dt, np, x, v, e, b, t, pl, ql, xt, yt) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_preparation<N><<<grid_for(c, np, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(
                               eps, dt, np, x, v, e, b, t, pl, ql, xt, yt)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_gather_tau(const LaunchCtx &c, const MeshDev &m, const double *emesh, int ntau, int64_t np,
                              const double *xt, double *et, int wrap) {
    if (np <= 0) return cudaSuccess;
    k_gather_tau<<<grid_for(c, np * ntau, kBlock), kBlock, 0, c.stream>>>(m, reinterpret_cast<const double2 *>(emesh), ntau, np, xt,
                                                                          et, wrap);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_compute_f(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *b, const double *xt,
                             const double *yt, const double *et, double *fx, double *fy, int normalise) {
    if (np <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_compute_f_generic(c, ntau, eps, np, b, xt, yt, et, fx, fy, normalise) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_compute_f<N><<<grid_for(c, np, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(
                               eps, np, b, xt, yt, et, fx, fy, normalise)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_fft_tau(const LaunchCtx &c, int ntau, int64_t nvec, const double *in, double *out, int sign,
                           int normalise) {
    if (nvec <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_fft_tau_generic(c, ntau, nvec, in, out, sign, normalise) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_fft_tau<N><<<grid_for(c, nvec, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(nvec, in, out, sign,
                                                                                                           normalise)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_step_pointwise(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t,
                                  const double *pl, const double *ql, const double *xf, const double *fx,
                                  const double *gx, double *out) {
    if (np <= 0) return cudaSuccess;
    k_step_pointwise<<<grid_for(c, np * ntau, kBlock), kBlock, 0, c.stream>>>(ntau, eps, np, t, pl, ql, xf, fx, gx, out);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_step_fortran(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t,
                                const double *pl, const double *ql, double *xt, double *xf, const double *fx,
                                const double *gx, int corrector) {
    if (np <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_step_fortran_generic(c, ntau, eps, np, t, pl, ql, xt, xf, fx, gx, corrector) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_step_fortran<N><<<grid_for(c, np, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(
                               eps, np, t, pl, ql, xt, xf, fx, gx, corrector)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_deposit_tau(const LaunchCtx &c, const MeshDev &m, int ntau, double eps, int64_t np,
                               const double *xt, const double *t, double w, const RhoAcc &acc, double *x, int wrap) {
    if (np <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_deposit_tau_generic(c, m, ntau, eps, np, xt, t, w, acc, x, wrap) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_deposit_tau<N><<<grid_for(c, np, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(m, eps, np, xt, t, w,
                                                                                                           acc, x, wrap)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_compute_v(const LaunchCtx &c, int ntau, double eps, int64_t np, const double *t, const double *yt,
                             int yt_is_fourier, double *v) {
    if (np <= 0) return cudaSuccess;
    if (!ntau_supported(ntau)) return generic_ntau_supported(ntau) ? launch_compute_v_generic(c, ntau, eps, np, t, yt, yt_is_fourier, v) : cudaErrorInvalidValue;
    UAPIC_DISPATCH_N(ntau, (k_compute_v<N><<<grid_for(c, np, (kBlock / 32) * (32 / N)), kBlock, 0, c.stream>>>(eps, np, t, yt,
                                                                                                         yt_is_fourier, v)));
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_deposit(const LaunchCtx &c, const MeshDev &m, int64_t np, double *x, double w, const RhoAcc &acc,
                           int wrap, int scheme) {
    if (np <= 0) return cudaSuccess;
    k_deposit<<<grid_for(c, np, kBlock), kBlock, 0, c.stream>>>(m, np, x, w, acc, wrap, scheme);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_gather(const LaunchCtx &c, const MeshDev &m, const double *emesh, int64_t np, double *x, double *ep,
                          int wrap, int scheme) {
    if (np <= 0) return cudaSuccess;
    k_gather<<<grid_for(c, np, kBlock), kBlock, 0, c.stream>>>(m, reinterpret_cast<const double2 *>(emesh), np, x, ep, wrap, scheme);
    count(c);
    return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(kBlock) k_fold_raw(RhoAcc acc, int64_t n, int copies) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (acc.i64) {
            unsigned long long t = acc.i64[i];
            for (int c = 1; c < copies; ++c) t += acc.i64[(size_t)c * n + i];       // fixed order; integer sum is exact anyway
            acc.i64[i] = t;
        } else {
            double t = acc.f64[i];
            for (int c = 1; c < copies; ++c) t += acc.f64[(size_t)c * n + i];
            acc.f64[i] = t;
        }
    }
}
}  // namespace

namespace {
// fold the CTA-private copies of `acc` (n elements each, `copies` of them) into `dst` (the exchange buffer peers read)
__global__ void __launch_bounds__(kBlock) k_fold_to(RhoAcc acc, int64_t n, int copies, unsigned long long *dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (acc.i64) {
            unsigned long long t = acc.i64[i];
            for (int c = 1; c < copies; ++c) t += acc.i64[(size_t)c * n + i];
            dst[i] = t;
        } else {
            double t = acc.f64[i];
            for (int c = 1; c < copies; ++c) t += acc.f64[(size_t)c * n + i];
            reinterpret_cast<double *>(dst)[i] = t;
        }
    }
}
// "my deposits of exchange number `seq` are in place": release at system scope, peers acquire it over NVLink
__global__ void k_publish(unsigned long long *flag, unsigned long long seq) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
}
}  // namespace

cudaError_t launch_fold_publish(const LaunchCtx &c, const RhoAcc &acc, int64_t n, int copies, void *dst, unsigned long long *flag,
                                unsigned long long seq) {
    if (n <= 0) return cudaErrorInvalidValue;
    k_fold_to<<<grid_for(c, n, kBlock), kBlock, 0, c.stream>>>(acc, n, copies < 1 ? 1 : copies, reinterpret_cast<unsigned long long *>(dst));
    k_publish<<<1, 1, 0, c.stream>>>(flag, seq);
    count(c, 2);
    return cudaGetLastError();
}

cudaError_t launch_fold_raw(const LaunchCtx &c, const RhoAcc &acc, int64_t n, int copies) {
    if (copies <= 1 || n <= 0) return cudaSuccess;
    k_fold_raw<<<grid_for(c, n, kBlock), kBlock, 0, c.stream>>>(acc, n, copies);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_rho_epilogue(const LaunchCtx &c, const MeshDev &m, const RhoAcc &acc, double *rho, double *rho_total) {
    k_rho_epilogue<<<1, kMeshBlock, 0, c.stream>>>(m, acc, rho, rho_total);
    count(c);
    return cudaGetLastError();
}

bool poisson_size_supported(int n) {
    if (n < 2) return false;
    const bool pow2 = (n & (n - 1)) == 0;
    return pow2 ? n <= 1024 : n <= 512;
}

cudaError_t launch_poisson(const LaunchCtx &c, const MeshDev &m, const PoissonWork &w, const double *rho, double *emesh,
                           double *energy) {
    if (!poisson_size_supported(m.nx) || !poisson_size_supported(m.ny)) return cudaErrorInvalidValue;
    const int tx = m.nx >= 512 ? 256 : (m.nx >= 128 ? 128 : 64);
    const int ty = m.ny >= 512 ? 256 : (m.ny >= 128 ? 128 : 64);
    const size_t shx = sizeof(double2) * (size_t)m.nx * 3;
    const size_t shy = sizeof(double2) * (size_t)m.ny * 5;
    // above the 48 KB default (ny = 1024 needs 80 KB, nx = 1024 exactly 48 KB) the kernels must opt in
    if (shx > 48 * 1024 - 256) {
        cudaError_t e = cudaFuncSetAttribute(k_poisson_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shx);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_poisson_rows_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shx);
        if (e != cudaSuccess) return e;
    }
    if (shy > 48 * 1024 - 256) {
        cudaError_t e = cudaFuncSetAttribute(k_poisson_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shy);
        if (e != cudaSuccess) return e;
    }
    k_poisson_rows_fwd<<<m.ny, tx, shx, c.stream>>>(m, rho, w.rk);
    k_poisson_cols<<<m.nx / 2 + 1, ty, shy, c.stream>>>(m, w.rk, w.ek);
    k_poisson_rows_bwd<<<dim3(m.ny, 2), tx, shx, c.stream>>>(m, w.ek, emesh);
    count(c, 3);
    if (energy) {
        k_energy<<<1, kMeshBlock, 0, c.stream>>>(m, reinterpret_cast<const double2 *>(emesh), energy);
        count(c);
    }
    return cudaGetLastError();
}

size_t field_solve_scratch_bytes() { return sizeof(double) * 4 * kSolveParts; }

cudaError_t launch_field_solve(const LaunchCtx &c, const MeshDev &m, const SolveBatch &B) {
    if (B.nb < 1 || B.nb > 2) return cudaErrorInvalidValue;
    if (!poisson_size_supported(m.nx) || !poisson_size_supported(m.ny) || m.nx < 4 || m.ny < 4) return cudaErrorInvalidValue;
    const size_t smem = sizeof(double2) * (size_t)max(3 * m.nx, 5 * m.ny);
    cudaError_t e;
    if (smem > 48 * 1024 - 4096) {      // per device and per context: set it every time it is needed (a host-side table write)
        e = cudaFuncSetAttribute(k_field_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_field_solve, kSolveBlock, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const int want = B.nb * 2 * m.ny;                      // the phase with the most tasks
    int grid = c.sm_count * (per_sm > 2 ? 2 : per_sm);
    if (grid > want) grid = want;
    MeshDev mm = m;
    SolveBatch bb = B;
    void *args[] = {&mm, &bb};
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_field_solve), dim3(grid), dim3(kSolveBlock), args, smem, c.stream);
    if (e != cudaSuccess) return e;
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_generate(const LaunchCtx &c, const MeshDev &m, int kind, uint64_t seed, int64_t first, int64_t stride, int64_t np,
                            int64_t np_global, double alpha, double kx, double *x, double *v) {
    if (np <= 0) return cudaSuccess;
    k_generate<<<grid_for(c, np, kBlock), kBlock, 0, c.stream>>>(m, kind, seed, first, stride, np, np_global, alpha, kx, x, v);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_extend_emesh(const LaunchCtx &c, const MeshDev &m, const double *emesh, double2 *ehalo) {
    if (m.nx < 4 || m.ny < 4) return cudaErrorInvalidValue;
    const int n = (m.nx + 6) * (m.ny + 6);
    k_extend_emesh<<<grid_for(c, n, kBlock), kBlock, 0, c.stream>>>(m, reinterpret_cast<const double2 *>(emesh), ehalo);
    count(c);
    return cudaGetLastError();
}

cudaError_t launch_extend_emesh_tiled(const LaunchCtx &c, const MeshDev &m, const double *emesh, double2 *ehalo) {
    if (m.nx < 4 || m.ny < 4) return cudaErrorInvalidValue;
    const int n = (int)ehalo_tiled_nodes(m);
    k_extend_emesh_tiled<<<grid_for(c, n, kBlock), kBlock, 0, c.stream>>>(m, reinterpret_cast<const double2 *>(emesh), ehalo);
    count(c);
    return cudaGetLastError();
}

size_t sum_v_scratch_bytes() { return sizeof(double2) * (size_t)kSumParts; }

cudaError_t launch_sum_v(const LaunchCtx &c, int64_t np, const double *v, double *scratch, double *out2) {
    k_sum_v_partial<<<kSumParts, kMeshBlock, 0, c.stream>>>(np, reinterpret_cast<const double2 *>(v), reinterpret_cast<double2 *>(scratch));
    k_sum_v_final<<<1, kMeshBlock, 0, c.stream>>>(reinterpret_cast<const double2 *>(scratch), out2);
    count(c, 2);
    return cudaGetLastError();
}

// DFMA issue rate of the whole chip: `launches` back-to-back launches of a pure DFMA loop, timed with CUDA events
cudaError_t probe_fp64_peak(const LaunchCtx &c, int launches, double *dfma_per_s, double *ms_per_launch) {
    double *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, 8);
    if (e != cudaSuccess) return e;
    const int iters = 1 << 16, grid = c.sm_count * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe_dfma<<<grid, block, 0, c.stream>>>(iters, 1.0, sink);          // warm-up
    cudaEventRecord(e0, c.stream);
    for (int l = 0; l < launches; ++l) k_probe_dfma<<<grid, block, 0, c.stream>>>(iters, 1.0 + l, sink);
    cudaEventRecord(e1, c.stream);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess) return e;
    const double n = (double)grid * block * kProbeChains * (double)iters * launches;
    if (dfma_per_s) *dfma_per_s = n / (ms * 1e-3);
    if (ms_per_launch) *ms_per_launch = ms / launches;
    return cudaGetLastError();
}

}  // namespace uapic

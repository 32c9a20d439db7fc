/*
 * uapic_oracle.c  --  CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the UA-PIC hot path of JuliaVlasov/UAPIC.jl, following the
 * Fortran `bupdate` program statement by statement (operation order included).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this file's shared object; the CUDA product path never does.
 *
 * PARITY STATUS: the reference holds no golden vector for the end-to-end run and neither
 * a Fortran compiler, Julia nor FFTW exist in this image, so positions / velocities /
 * energy history are "parity unpinned" by the reference itself.  What *is* pinned:
 *   - test/test_poisson.jl:28,47       (Poisson vs analytic, atol 1e-14)
 *   - test/test_particles.jl:45        (integral of rho ~ 0)
 *   - test/test_particles.jl:73-74     (M6 interpolation reproduces a linear field)
 *   - efd.f90:481                      (the external-field program's printed sum(v), 13 digits: orc_efd_run below, run on
 *                                       init_particles_2d's load regenerated from the libgfortran stream)
 * and an independently written numpy twin (oracle/uapic_oracle_np.py, following the Julia
 * sources) must agree with this file to <= 1e-12 (tests/test_oracle.py).
 *
 * Every function cites the reference lines it follows.  FFTs: FFTW is a third-party
 * dependency absent from /root/reference (Project.toml:8 `FFTW`, fortran/Makefile:7
 * `-lfftw3`, version unpinned); a DFT is mathematically unique, so an own radix-2 /
 * direct DFT is used, and FFTW's c2r treatment of non-Hermitian input is reproduced
 * explicitly (see orc_poisson).
 *
 * Arrays use the reference's column-major layouts:
 *   xt, yt, fx, ... : complex(8) (ntau, 2, nbpart)  -> interleaved re,im, tau fastest
 *   et              : real(8)    (ntau, 2, nbpart)
 *   x, v, e         : real(8)    (2, nbpart)
 *   mesh e          : real(8)    (2, nx+1, ny+1)
 *   mesh rho        : real(8)    (nx+1, ny+1)
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp; no fast-math, no FMA
 * contraction: gfortran -O3 on baseline x86-64 does not fuse either).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NTAU 256   /* stack arrays of the per-particle tau loops */

typedef struct { double re, im; } cplx;

static inline cplx c_make(double re, double im) { cplx z = { re, im }; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
/* complex * complex as gfortran emits it (no NaN recovery needed for finite data) */
static inline cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cplx c_rmul(double r, cplx a) { return c_make(r * a.re, r * a.im); }
static inline cplx c_rdiv(cplx a, double r) { return c_make(a.re / r, a.im / r); }
/* exp((0,phi)) -> cexp: exp(0)*(cos, sin) */
static inline cplx c_expi(double phi) { return c_make(cos(phi), sin(phi)); }

#define ORC_WRAP_FORTRAN 0   /* px = x/dx ; px = modulo(px, nx)          (compute_rho_m6.F90:89-93)  */
#define ORC_WRAP_JULIA   1   /* x  = mod(x-xmin, dimx) ; px = x/dx        (src/compute_rho.jl:63-67)  */

typedef struct {
    double xmin, xmax, ymin, ymax;
    int32_t nx, ny;
} orc_mesh;

static int g_threads = 1;

void orc_set_threads(int n)
{
    if (n < 1) n = 1;
    g_threads = n;
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* FFT (stand-in for FFTW's unnormalised complex DFT, sign -1 forward / +1 backward)     */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int n;
    int pow2;
    cplx *w;       /* w[k] = exp(-2 pi i k / n), k = 0..n-1 */
    int *brev;
} orc_fft_plan;

static orc_fft_plan *orc_fft_new(int n)
{
    orc_fft_plan *p = (orc_fft_plan *)malloc(sizeof(*p));
    p->n = n;
    p->pow2 = (n > 0) && ((n & (n - 1)) == 0);
    p->w = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    p->brev = (int *)malloc(sizeof(int) * (size_t)n);
    const double pi = 4.0 * atan(1.0);
    for (int k = 0; k < n; k++) {
        double a = -2.0 * pi * (double)k / (double)n;
        p->w[k] = c_make(cos(a), sin(a));
    }
    /* exact values on the axes / diagonals keep the transform of smooth data clean */
    if (n % 4 == 0) { p->w[n / 4] = c_make(0.0, -1.0); p->w[3 * n / 4] = c_make(0.0, 1.0); }
    if (n % 2 == 0) { p->w[n / 2] = c_make(-1.0, 0.0); }
    int bits = 0;
    while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        p->brev[i] = r;
    }
    return p;
}

static void orc_fft_free(orc_fft_plan *p) { if (p) { free(p->w); free(p->brev); free(p); } }

/* out-of-place capable (in may equal out); stride in elements; sign = -1 fwd, +1 bwd */
static void orc_fft_exec(const orc_fft_plan *p, const cplx *in, int istride, cplx *out, int ostride, int sign)
{
    const int n = p->n;
    cplx stackbuf[512];
    cplx *buf = (n <= 512) ? stackbuf : (cplx *)malloc(sizeof(cplx) * (size_t)n);
    if (p->pow2) {
        for (int i = 0; i < n; i++) buf[p->brev[i]] = in[(size_t)i * istride];
        for (int len = 2; len <= n; len <<= 1) {
            int half = len >> 1, step = n / len;
            for (int s = 0; s < n; s += len) {
                for (int k = 0; k < half; k++) {
                    cplx w = p->w[k * step];
                    if (sign > 0) w.im = -w.im;
                    cplx a = buf[s + k];
                    cplx b = c_mul(w, buf[s + k + half]);
                    buf[s + k] = c_add(a, b);
                    buf[s + k + half] = c_sub(a, b);
                }
            }
        }
        for (int i = 0; i < n; i++) out[(size_t)i * ostride] = buf[i];
    } else {
        cplx tmp2[512];
        cplx *src = (n <= 512) ? tmp2 : (cplx *)malloc(sizeof(cplx) * (size_t)n);
        for (int i = 0; i < n; i++) src[i] = in[(size_t)i * istride];
        for (int k = 0; k < n; k++) {
            cplx acc = c_make(0.0, 0.0);
            for (int j = 0; j < n; j++) {
                cplx w = p->w[(int)(((int64_t)k * j) % n)];
                if (sign > 0) w.im = -w.im;
                acc = c_add(acc, c_mul(w, src[j]));
            }
            buf[k] = acc;
        }
        for (int i = 0; i < n; i++) out[(size_t)i * ostride] = buf[i];
        if (n > 512) free(src);
    }
    if (n > 512) free(buf);
}

/* exported for tests of the tau FFT (mul!(x̃t, ftau, xt) / ifft!(xt,1) in test/bupdate.jl:79,85) */
void orc_fft_tau(int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise)
{
    orc_fft_plan *p = orc_fft_new(ntau);
    for (int64_t k = 0; k < nvec; k++) {
        orc_fft_exec(p, (const cplx *)in + k * ntau, 1, (cplx *)out + k * ntau, 1, sign);
        if (normalise) {
            cplx *o = (cplx *)out + k * ntau;
            for (int n = 0; n < ntau; n++) o[n] = c_rdiv(o[n], (double)ntau);
        }
    }
    orc_fft_free(p);
}

/* ------------------------------------------------------------------------------------ */
/* ua_t : tau grid and wavenumbers      fortran/ua_type.F90:32-76, src/ua_type.jl:17-41  */
/* ------------------------------------------------------------------------------------ */

void orc_ua_tables(int ntau, double *tau, double *ltau)
{
    const double pi = 4.0 * atan(1.0);
    double dtau = 2.0 * pi / (double)ntau;                      /* ua_type.F90:47 */
    for (int n = 1; n <= ntau / 2; n++) ltau[n - 1] = (double)(n - 1);          /* :51-53 */
    for (int n = ntau / 2 + 1; n <= ntau; n++) ltau[n - 1] = (double)(n - 1 - ntau); /* :54-56 */
    for (int n = 1; n <= ntau; n++) tau[n - 1] = (double)(n - 1) * dtau;         /* :60-62 */
}

/* ------------------------------------------------------------------------------------ */
/* M6 kernel                       fortran/compute_rho_m6.F90:28-45, src/compute_rho.jl:10-25 */
/* ------------------------------------------------------------------------------------ */

static inline double pow5(double x) { double x2 = x * x; double x4 = x2 * x2; return x4 * x; }

static inline double f_m6(double q)
{
    double f;
    if (q < 1.0)                      f = pow5(3.0 - q) - 6.0 * pow5(2.0 - q) + 15.0 * pow5(1.0 - q);
    else if (q >= 1.0 && q < 2.0)     f = pow5(3.0 - q) - 6.0 * pow5(2.0 - q);
    else if (q >= 2.0 && q < 3.0)     f = pow5(3.0 - q);
    else                              f = 0.0;
    return f / 120.0;
}

double orc_f_m6(double q) { return f_m6(q); }

/* Fortran MODULO / Julia mod for reals: result has the sign of p */
static inline double f_modulo(double a, double p)
{
    double r = fmod(a, p);
    if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
    return r;
}
static inline int i_modulo(int a, int p) { int r = a % p; if (r != 0 && ((r < 0) != (p < 0))) r += p; return r; }

typedef struct {
    int ix[7], jy[7];   /* 0-based wrapped node indices for offsets -3..3 (centre = i, which may be nx: edge case) */
    double cx[7], cy[7];
} m6_stencil;

/* Shape function.  M6 is what the reference ships on this path.  CIC is BUILD-DEFINED (the reference has no 2D CIC
   deposit and no CIC inside the UA loop): the bilinear weights of performance/test_cic.F90:73-76
   (= the 2D restriction of fortran/compute_rho_cic.f90:46-53), a1 = (1-dpx)(1-dpy) on (i,j), a2 = dpx(1-dpy) on (i+1,j),
   a3 = dpx dpy on (i+1,j+1), a4 = (1-dpx) dpy on (i,j+1), with the periodic index wrap, the ghost copy, the /(dx dy) and
   the neutralisation of the M6 path.  Here they are the offsets 0 and +1 of the same 7 x 7 stencil structure. */
#define ORC_SCHEME_M6 0
#define ORC_SCHEME_CIC 1
static int g_scheme = ORC_SCHEME_M6;
void orc_set_scheme(int scheme) { g_scheme = (scheme == ORC_SCHEME_CIC) ? ORC_SCHEME_CIC : ORC_SCHEME_M6; }
int orc_get_scheme(void) { return g_scheme; }

/* cell + weights: compute_rho_m6.F90:89-131 / interpolation_m6.F90:87-127 (Fortran wrap);
   src/compute_rho.jl:63-73 / src/interpolation.jl:19-29 (Julia wrap).  xw/yw return the position
   the Julia variant stores back into particles.x */
static inline void m6_setup(const orc_mesh *m, double dx, double dy, double x, double y, int wrap,
                            m6_stencil *s, double *xw, double *yw)
{
    const int nx = m->nx, ny = m->ny;
    double px, py;
    if (wrap == ORC_WRAP_JULIA) {
        double dimx = m->xmax - m->xmin, dimy = m->ymax - m->ymin;
        double xn = f_modulo(x - m->xmin, dimx);
        double yn = f_modulo(y - m->ymin, dimy);
        px = xn / dx; py = yn / dy;
        if (xw) { *xw = xn + m->xmin; *yw = yn + m->ymin; }
    } else {
        px = x / dx; py = y / dy;
        px = f_modulo(px, (double)nx);
        py = f_modulo(py, (double)ny);
        if (xw) { *xw = x; *yw = y; }
    }
    int i = (int)floor(px); double dpx = px - (double)i;
    int j = (int)floor(py); double dpy = py - (double)j;
    for (int a = -3; a <= 3; a++) {
        s->ix[a + 3] = (a == 0) ? i : i_modulo(i + a, nx);   /* centre index is i+1 (1-based), NOT wrapped */
        s->jy[a + 3] = (a == 0) ? j : i_modulo(j + a, ny);
    }
    if (g_scheme == ORC_SCHEME_CIC) {
        for (int a = 0; a < 7; a++) { s->cx[a] = 0.0; s->cy[a] = 0.0; }
        s->cx[3] = 1.0 - dpx; s->cx[4] = dpx;          /* test_cic.F90:73-76 */
        s->cy[3] = 1.0 - dpy; s->cy[4] = dpy;
        return;
    }
    s->cx[0] = f_m6(3.0 + dpx); s->cx[6] = f_m6(3.0 - dpx);
    s->cx[1] = f_m6(2.0 + dpx); s->cx[5] = f_m6(2.0 - dpx);
    s->cx[2] = f_m6(1.0 + dpx); s->cx[4] = f_m6(1.0 - dpx);
    s->cx[3] = f_m6(dpx);
    s->cy[0] = f_m6(3.0 + dpy); s->cy[6] = f_m6(3.0 - dpy);
    s->cy[1] = f_m6(2.0 + dpy); s->cy[5] = f_m6(2.0 - dpy);
    s->cy[2] = f_m6(1.0 + dpy); s->cy[4] = f_m6(1.0 - dpy);
    s->cy[3] = f_m6(dpy);
}

/* 49-term sequential sum, x offset outer, y offset inner: interpolation_m6.F90:130-183 */
static inline void m6_gather(const orc_mesh *m, const double *emesh, const m6_stencil *s, double *e1, double *e2)
{
    const size_t ld = (size_t)(m->nx + 1);
    for (int l = 0; l < 2; l++) {
        double acc = 0.0;
        for (int a = 0; a < 7; a++)
            for (int b = 0; b < 7; b++)
                acc = acc + s->cx[a] * s->cy[b] * emesh[l + 2 * ((size_t)s->ix[a] + ld * (size_t)s->jy[b])];
        if (l == 0) *e1 = acc; else *e2 = acc;
    }
}

/* 49 read-modify-writes: compute_rho_m6.F90:133-187 */
static inline void m6_scatter(const orc_mesh *m, double *rho, const m6_stencil *s, double weight)
{
    const size_t ld = (size_t)(m->nx + 1);
    for (int a = 0; a < 7; a++)
        for (int b = 0; b < 7; b++)
            rho[(size_t)s->ix[a] + ld * (size_t)s->jy[b]] += s->cx[a] * s->cy[b] * weight;
}

/* epilogue: ghost copy -> /(dx dy) -> subtract mean   compute_rho_m6.F90:191-200 */
static double rho_epilogue(const orc_mesh *m, double *rho)
{
    const int nx = m->nx, ny = m->ny;
    const size_t ld = (size_t)(nx + 1);
    const double dx = (m->xmax - m->xmin) / (double)nx, dy = (m->ymax - m->ymin) / (double)ny;
    const double dimx = m->xmax - m->xmin, dimy = m->ymax - m->ymin;
    for (int i = 0; i < nx; i++) rho[i + ld * ny] = rho[i];
    for (int j = 0; j < ny; j++) rho[nx + ld * j] = rho[ld * j];
    rho[nx + ld * ny] = rho[0];
    const double dxdy = dx * dy;
    for (size_t k = 0; k < ld * (size_t)(ny + 1); k++) rho[k] = rho[k] / dxdy;
    double tot = 0.0;
    for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) tot += rho[i + ld * j];  /* column-major sum */
    double rho_total = tot * dx * dy;
    double sub = rho_total / dimx / dimy;
    for (size_t k = 0; k < ld * (size_t)(ny + 1); k++) rho[k] = rho[k] - sub;
    return rho_total;
}

static inline double mesh_dx(const orc_mesh *m) { return (m->xmax - m->xmin) / (double)m->nx; }  /* meshfields.F90:71 */
static inline double mesh_dy(const orc_mesh *m) { return (m->ymax - m->ymin) / (double)m->ny; }  /* meshfields.F90:72 */

/* private-rho helper for the threaded deposit: thread-ordered reduction keeps results deterministic
   for a fixed thread count; with 1 thread the summation order is exactly the reference's */
typedef void (*deposit_body)(int64_t k, double *rho, void *ctx);

static void deposit_driver(const orc_mesh *m, int64_t np, double *rho, deposit_body body, void *ctx)
{
    const size_t nrho = (size_t)(m->nx + 1) * (size_t)(m->ny + 1);
    memset(rho, 0, sizeof(double) * nrho);
    if (g_threads <= 1) {
        for (int64_t k = 0; k < np; k++) body(k, rho, ctx);
        return;
    }
#ifdef _OPENMP
    int nt = g_threads;
    double *priv = (double *)calloc(nrho * (size_t)nt, sizeof(double));
    #pragma omp parallel num_threads(nt)
    {
        int tid = omp_get_thread_num();
        int64_t lo = np * tid / nt, hi = np * (tid + 1) / nt;
        double *r = priv + nrho * (size_t)tid;
        for (int64_t k = lo; k < hi; k++) body(k, r, ctx);
    }
    for (int t = 0; t < nt; t++) for (size_t q = 0; q < nrho; q++) rho[q] += priv[nrho * (size_t)t + q];
    free(priv);
#else
    for (int64_t k = 0; k < np; k++) body(k, rho, ctx);
#endif
}

/* ------------------------------------------------------------------------------------ */
/* compute_rho_m6_real        fortran/compute_rho_m6.F90:205-335, src/compute_rho.jl:181-316 */
/* ------------------------------------------------------------------------------------ */

typedef struct { const orc_mesh *m; double *x; double w; int wrap; double dx, dy; } dep_real_ctx;

static void dep_real_body(int64_t k, double *rho, void *vctx)
{
    dep_real_ctx *c = (dep_real_ctx *)vctx;
    m6_stencil s; double xw, yw;
    m6_setup(c->m, c->dx, c->dy, c->x[2 * k], c->x[2 * k + 1], c->wrap, &s, &xw, &yw);
    if (c->wrap == ORC_WRAP_JULIA) { c->x[2 * k] = xw; c->x[2 * k + 1] = yw; }
    m6_scatter(c->m, rho, &s, c->w);
}

/* returns rho_total (the value both languages print).  NOTE the Julia method swaps xmin/xmax
   (src/compute_rho.jl:190-191) which only shifts the stored x by one period; the oracle's Julia
   variant stores the in-box position instead (documented deviation, periodic-equivalent). */
double orc_compute_rho_m6(const orc_mesh *m, int64_t np, double *x, double w, double *rho, int wrap)
{
    dep_real_ctx c = { m, x, w, wrap, mesh_dx(m), mesh_dy(m) };
    deposit_driver(m, np, rho, dep_real_body, &c);
    return rho_epilogue(m, rho);
}

/* Deterministic (fixed-point) deposit: NOT in the reference.  It restates the product's build-defined
   UAPIC_DEPOSIT_FIXED_POINT mode so that mode can be checked bit for bit: every tap cx*cy*w (reference
   operation order) is rounded to a multiple of 1/scale and summed in int64; rho_total is the exact integer
   sum; the epilogue is otherwise compute_rho_m6.F90:191-200.  pos = positions (2,np) already evaluated. */
double orc_compute_rho_m6_fixed(const orc_mesh *m, int64_t np, const double *x, double w, double scale, double *rho, int wrap)
{
    const int nx = m->nx, ny = m->ny;
    const size_t ld = (size_t)(nx + 1), nrho = ld * (size_t)(ny + 1);
    const double dx = mesh_dx(m), dy = mesh_dy(m);
    int64_t *acc = (int64_t *)calloc(nrho, sizeof(int64_t));
    for (int64_t k = 0; k < np; k++) {
        m6_stencil s;
        m6_setup(m, dx, dy, x[2 * k], x[2 * k + 1], wrap, &s, NULL, NULL);
        for (int a = 0; a < 7; a++)
            for (int b = 0; b < 7; b++)
                acc[(size_t)s.ix[a] + ld * (size_t)s.jy[b]] += (int64_t)llrint(s.cx[a] * s.cy[b] * w * scale);
    }
    int64_t itot = 0;
    for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) itot += acc[i + ld * j];
    const double total = (double)itot * (1.0 / scale);
    const double sub = total / (m->xmax - m->xmin) / (m->ymax - m->ymin);
    const double dxdy = dx * dy;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            double raw = (double)acc[i + ld * j] * (1.0 / scale);
            double val = raw / dxdy;
            rho[i + ld * j] = val - sub;
        }
    for (int i = 0; i < nx; i++) rho[i + ld * ny] = rho[i];
    for (int j = 0; j < ny; j++) rho[nx + ld * j] = rho[ld * j];
    rho[nx + ld * ny] = rho[0];
    free(acc);
    return total;
}

/* ------------------------------------------------------------------------------------ */
/* interpolate_eb_m6_real     fortran/interpolation_m6.F90:193-327, src/interpolation.jl:125-247 */
/* ------------------------------------------------------------------------------------ */

void orc_interpol_eb_m6(const orc_mesh *m, const double *emesh, int64_t np, double *x, double *ep, int wrap)
{
    const double dx = mesh_dx(m), dy = mesh_dy(m);
    #pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t k = 0; k < np; k++) {
        m6_stencil s; double xw, yw;
        m6_setup(m, dx, dy, x[2 * k], x[2 * k + 1], wrap, &s, &xw, &yw);
        if (wrap == ORC_WRAP_JULIA) { x[2 * k] = xw; x[2 * k + 1] = yw; }   /* src/interpolation.jl:152-153 */
        m6_gather(m, emesh, &s, &ep[2 * k], &ep[2 * k + 1]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* interpolate_eb_m6_complex  fortran/interpolation_m6.F90:40-191, src/interpolation.jl:3-123 */
/* ------------------------------------------------------------------------------------ */

void orc_interpol_eb_m6_tau(const orc_mesh *m, const double *emesh, int ntau, int64_t np,
                            const double *xt, double *et, int wrap)
{
    const double dx = mesh_dx(m), dy = mesh_dy(m);
    const cplx *X = (const cplx *)xt;
    #pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t k = 0; k < np; k++) {
        for (int n = 0; n < ntau; n++) {
            m6_stencil s;
            double xr = X[n + (size_t)ntau * (0 + 2 * k)].re;          /* real(x(n,1,k)) :87 */
            double yr = X[n + (size_t)ntau * (1 + 2 * k)].re;          /* real(x(n,2,k)) :88 */
            m6_setup(m, dx, dy, xr, yr, wrap, &s, NULL, NULL);
            m6_gather(m, emesh, &s, &et[n + (size_t)ntau * (0 + 2 * k)], &et[n + (size_t)ntau * (1 + 2 * k)]);
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* compute_rho_m6_complex     fortran/compute_rho_m6.F90:47-203, src/compute_rho.jl:29-179   */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    const orc_mesh *m; int ntau; double eps; const cplx *xt; const double *t; double *x; double w; int wrap;
    const double *ltau; orc_fft_plan *plan; double dx, dy;
} dep_tau_ctx;

static void dep_tau_body(int64_t k, double *rho, void *vctx)
{
    dep_tau_ctx *c = (dep_tau_ctx *)vctx;
    const int N = c->ntau;
    cplx ft[ORC_MAX_NTAU];
    double pos[2];
    const double t = c->t[k];
    for (int comp = 0; comp < 2; comp++) {
        orc_fft_exec(c->plan, c->xt + (size_t)N * (comp + 2 * k), 1, ft, 1, -1);        /* :74, :80 */
        cplx sum = c_make(0.0, 0.0);
        for (int n = 0; n < N; n++) {
            /* exp(cmplx(0,1)*ltau*t/eps)/cmplx(ntau,0)    :76, :82 */
            cplx el = c_expi(c->ltau[n] * t / c->eps);
            ft[n] = c_rdiv(c_mul(ft[n], el), (double)N);     /* (ft*exp)/ntau, left to right */
        }
        for (int n = 0; n < N; n++) sum = c_add(sum, ft[n]);                             /* sum(ua%ft) :78 */
        pos[comp] = sum.re;
    }
    m6_stencil s; double xw, yw;
    m6_setup(c->m, c->dx, c->dy, pos[0], pos[1], c->wrap, &s, &xw, &yw);
    c->x[2 * k] = xw; c->x[2 * k + 1] = yw;       /* Fortran: unwrapped (:86-87); Julia: wrapped (src/compute_rho.jl:69-70) */
    m6_scatter(c->m, rho, &s, c->w);
}

double orc_compute_rho_m6_tau(const orc_mesh *m, int ntau, double eps, int64_t np, const double *xt,
                              const double *t, double w, double *rho, double *x, int wrap)
{
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    if (ntau > ORC_MAX_NTAU) return NAN;
    orc_ua_tables(ntau, tau, ltau);
    orc_fft_plan *plan = orc_fft_new(ntau);
    dep_tau_ctx c = { m, ntau, eps, (const cplx *)xt, t, x, w, wrap, ltau, plan, mesh_dx(m), mesh_dy(m) };
    deposit_driver(m, np, rho, dep_tau_body, &c);
    orc_fft_free(plan);
    return rho_epilogue(m, rho);
}

/* ------------------------------------------------------------------------------------ */
/* Poisson                         fortran/poisson_2d.f90:30-111, src/poisson.jl:14-83        */
/* ------------------------------------------------------------------------------------ */
/* FFTW semantics reproduced: r2c with x (first Fortran dim) halved; c2r = complex inverse DFT
   along y for every kx column, then a 1-D c2r along x that ignores Im of the kx=0 and kx=nx/2
   bins (what FFTW's rdft2 and pocketfft's irfftn do on non-Hermitian input).                 */

double orc_poisson(const orc_mesh *m, const double *rho, double *e)
{
    const int nx = m->nx, ny = m->ny, nh = nx / 2 + 1;
    const size_t ld = (size_t)(nx + 1);
    const double pi = 4.0 * atan(1.0);
    const double dx = mesh_dx(m), dy = mesh_dy(m);
    const double kx0 = 2.0 * pi / (m->xmax - m->xmin);     /* poisson_2d.f90:48 */
    const double ky0 = 2.0 * pi / (m->ymax - m->ymin);     /* :49 */

    orc_fft_plan *px = orc_fft_new(nx), *py = orc_fft_new(ny);
    cplx *rk = (cplx *)malloc(sizeof(cplx) * (size_t)nh * ny);
    cplx *ek = (cplx *)malloc(sizeof(cplx) * (size_t)nh * ny);
    cplx *row = (cplx *)malloc(sizeof(cplx) * (size_t)(nx > ny ? nx : ny));
    cplx *rowo = (cplx *)malloc(sizeof(cplx) * (size_t)(nx > ny ? nx : ny));

    /* forward r2c: x first, then y   (:96) */
    for (int j = 0; j < ny; j++) {
        for (int i = 0; i < nx; i++) row[i] = c_make(rho[i + ld * j], 0.0);
        orc_fft_exec(px, row, 1, rowo, 1, -1);
        for (int i = 0; i < nh; i++) rk[i + (size_t)nh * j] = rowo[i];
    }
    for (int i = 0; i < nh; i++) orc_fft_exec(py, rk + i, nh, rk + i, nh, -1);

    for (int comp = 0; comp < 2; comp++) {
        for (int jk = 0; jk < ny; jk++) {
            for (int ik = 0; ik < nh; ik++) {
                double kx = (double)ik * kx0;                                        /* :59 */
                double ky = (jk < ny / 2) ? (double)jk * ky0 : (double)(jk - ny) * ky0; /* :62, :66 */
                if (ik == 0 && jk == 0) kx = 1.0;                                    /* :70 */
                double k2 = kx * kx + ky * ky;                                       /* :71 */
                double kk = (comp == 0 ? kx : ky) / k2;                              /* :72-73 */
                /* (0,-1) * kk * rho_hat    :98-99 */
                cplx r = rk[ik + (size_t)nh * jk];
                cplx mk = c_make(0.0 * kk, -1.0 * kk);          /* cmplx(0,-1)*k  (k complex with zero imag) */
                ek[ik + (size_t)nh * jk] = c_mul(mk, r);
            }
        }
        /* c2r: inverse along y per kx column, then c2r along x   (:101-102) */
        for (int i = 0; i < nh; i++) orc_fft_exec(py, ek + i, nh, ek + i, nh, +1);
        for (int j = 0; j < ny; j++) {
            for (int i = 0; i < nh; i++) row[i] = ek[i + (size_t)nh * j];
            row[0].im = 0.0;
            if ((nx & 1) == 0) row[nh - 1].im = 0.0;
            for (int i = nh; i < nx; i++) row[i] = c_make(row[nx - i].re, -row[nx - i].im);
            orc_fft_exec(px, row, 1, rowo, 1, +1);
            for (int i = 0; i < nx; i++) e[comp + 2 * ((size_t)i + ld * j)] = rowo[i].re;
        }
    }
    /* ghosts (:104-107) in the reference's statement order, then /(nx*ny) (:109) */
    for (int comp = 0; comp < 2; comp++) {
        for (int j = 0; j <= ny; j++) e[comp + 2 * ((size_t)nx + ld * j)] = e[comp + 2 * (0 + ld * j)];
        for (int i = 0; i <= nx; i++) e[comp + 2 * ((size_t)i + ld * ny)] = e[comp + 2 * ((size_t)i + ld * 0)];
    }
    const double nn = (double)(nx * ny);
    /* interior was produced unnormalised; ghosts were copied from unnormalised interior */
    for (size_t k = 0; k < 2 * ld * (size_t)(ny + 1); k++) e[k] = e[k] / nn;

    /* energy  src/poisson.jl:80-81 : sum over the full ghosted array of e1^2+e2^2, times dx*dy */
    double nrj = 0.0;
    for (size_t k = 0; k < ld * (size_t)(ny + 1); k++) nrj += e[2 * k] * e[2 * k] + e[2 * k + 1] * e[2 * k + 1];
    nrj = nrj * dx * dy;

    free(rk); free(ek); free(row); free(rowo); orc_fft_free(px); orc_fft_free(py);
    return nrj;
}

/* ------------------------------------------------------------------------------------ */
/* preparation                      fortran/ua_steps.F90:15-115, src/ua_steps.jl:3-78       */
/* ------------------------------------------------------------------------------------ */

void orc_preparation(int ntau, double eps, double dt, int64_t np, const double *x, const double *v, const double *e,
                     double *b_out, double *t_out, double *pl_, double *ql_, double *xt_, double *yt_)
{
    const int N = ntau;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    orc_ua_tables(N, tau, ltau);
    cplx *PL = (cplx *)pl_, *QL = (cplx *)ql_, *XT = (cplx *)xt_, *YT = (cplx *)yt_;

    #pragma omp parallel if (g_threads > 1)
    {
    orc_fft_plan *plan = orc_fft_new(N);
    cplx rt[2][ORC_MAX_NTAU], rf[2][ORC_MAX_NTAU];
    #pragma omp for schedule(static)
    for (int64_t m = 0; m < np; m++) {
        double x1 = x[2 * m], x2 = x[2 * m + 1];                                  /* :51-52 */
        double b = 1.0 + 0.5 * sin(x1) * sin(x2);                                 /* :54 */
        double t = dt * b;                                                        /* :55 */
        b_out[m] = b; t_out[m] = t;                                               /* :57-58 */

        PL[(size_t)N * m] = c_make(t, 0.0);                                       /* :60 */
        QL[(size_t)N * m] = c_make(t * t / 2.0, 0.0);                             /* :61 */
        for (int n = 1; n < N; n++) {
            double l = ltau[n];
            /* elt = exp((0,-1)*ltau(n)*t/eps)    :64 */
            cplx elt = c_expi(((-1.0 * l) * t) / eps);
            /* pl = eps*(0,1)*(elt-1)/ltau(n)     :65 */
            cplx em1 = c_make(elt.re - 1.0, elt.im);
            cplx ie = c_make(eps * 0.0, eps * 1.0);                              /* eps * cmplx(0,1) */
            PL[n + (size_t)N * m] = c_rdiv(c_mul(ie, em1), l);
            /* ql = eps*(eps*(1-elt) - (0,1)*ltau(n)*t)/ltau(n)**2     :66 */
            cplx ome = c_make(1.0 - elt.re, -elt.im);
            cplx a = c_rmul(eps, ome);
            cplx ilt = c_make(0.0 * l * t, (1.0 * l) * t);
            cplx d = c_sub(a, ilt);
            QL[n + (size_t)N * m] = c_rdiv(c_rmul(eps, d), l * l);
        }

        double ex = e[2 * m], ey = e[2 * m + 1];                                  /* :69-70 */
        double vx = v[2 * m], vy = v[2 * m + 1];                                  /* :71-72 */
        double vxb = vx / b, vyb = vy / b;                                        /* :73-74 */

        for (int n = 0; n < N; n++) {
            double st = sin(tau[n]), ct = cos(tau[n]);
            double h1 = eps * (st * vxb - ct * vyb);                              /* :78 */
            double h2 = eps * (st * vyb + ct * vxb);                              /* :79 */
            double xt1 = x1 + h1 + eps * vyb;                                     /* :81 */
            double xt2 = x2 + h2 - eps * vxb;                                     /* :82 */
            XT[n + (size_t)N * (0 + 2 * m)] = c_make(xt1, 0.0);                   /* :84 */
            XT[n + (size_t)N * (1 + 2 * m)] = c_make(xt2, 0.0);                   /* :85 */
            double interv = (1.0 + 0.5 * sin(xt1) * sin(xt2) - b) / eps;          /* :87 */
            double exb = ((ct * vy - st * vx) * interv + ex) / b;                 /* :89 */
            double eyb = ((-ct * vx - st * vy) * interv + ey) / b;                /* :90 */
            rt[0][n] = c_make(ct * exb - st * eyb, 0.0);                          /* :92 */
            rt[1][n] = c_make(st * exb + ct * eyb, 0.0);                          /* :93 */
        }
        orc_fft_exec(plan, rt[0], 1, rf[0], 1, -1);                               /* :97 */
        orc_fft_exec(plan, rt[1], 1, rf[1], 1, -1);                               /* :98 */
        for (int n = 1; n < N; n++) {
            /* rf = -(0,1)/ltau(n) * rf / ntau     :101-102 */
            cplx mil = c_make(-0.0 / ltau[n], -1.0 / ltau[n]);
            rf[0][n] = c_rdiv(c_mul(mil, rf[0][n]), (double)N);
            rf[1][n] = c_rdiv(c_mul(mil, rf[1][n]), (double)N);
        }
        orc_fft_exec(plan, rf[0], 1, rt[0], 1, +1);                               /* :105 */
        orc_fft_exec(plan, rf[1], 1, rt[1], 1, +1);                               /* :106 */
        for (int n = 0; n < N; n++) {
            /* yt = v + (rt(n) - rt(1)) * eps      :109-110 */
            cplx d0 = c_rmul(eps, c_sub(rt[0][n], rt[0][0]));
            cplx d1 = c_rmul(eps, c_sub(rt[1][n], rt[1][0]));
            YT[n + (size_t)N * (0 + 2 * m)] = c_make(vx + d0.re, d0.im);
            YT[n + (size_t)N * (1 + 2 * m)] = c_make(vy + d1.re, d1.im);
        }
    }
    orc_fft_free(plan);
    }
}

/* ------------------------------------------------------------------------------------ */
/* compute_f                        fortran/ua_steps.F90:140-198, src/ua_steps.jl:105-145   */
/* normalise=1: Fortran (FFT then /ntau, :194-195); normalise=0: Julia (fft! only, :142-143) */
/* ------------------------------------------------------------------------------------ */

void orc_compute_f(int ntau, double eps, int64_t np, const double *b_, const double *xt_, const double *yt_,
                   const double *et, double *fx_, double *fy_, int normalise)
{
    const int N = ntau;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    orc_ua_tables(N, tau, ltau);
    const cplx *XT = (const cplx *)xt_, *YT = (const cplx *)yt_;
    cplx *FX = (cplx *)fx_, *FY = (cplx *)fy_;
    #pragma omp parallel if (g_threads > 1)
    {
    orc_fft_plan *plan = orc_fft_new(N);
    #pragma omp for schedule(static)
    for (int64_t m = 0; m < np; m++) {
        double b = b_[m];
        for (int n = 0; n < N; n++) {
            size_t i1 = n + (size_t)N * (0 + 2 * m), i2 = n + (size_t)N * (1 + 2 * m);
            double xt1 = XT[i1].re, xt2 = XT[i2].re;                              /* :166-167 */
            cplx yt1 = YT[i1], yt2 = YT[i2];                                      /* :169-170 */
            double ct = cos(tau[n]), st = sin(tau[n]);
            FX[i1] = c_rdiv(c_add(c_rmul(ct, yt1), c_rmul(st, yt2)), b);          /* :174 */
            FX[i2] = c_rdiv(c_add(c_rmul(-st, yt1), c_rmul(ct, yt2)), b);         /* :175 */
            double interv = (1.0 + 0.5 * sin(xt1) * sin(xt2) - b) / eps;          /* :177 */
            cplx q1 = c_rmul(interv, c_sub(c_rmul(ct, yt2), c_rmul(st, yt1)));    /* :179 */
            cplx q2 = c_rmul(interv, c_sub(c_rmul(-ct, yt1), c_rmul(st, yt2)));   /* :180 */
            cplx tmp1 = c_make(et[i1] + q1.re, q1.im);
            cplx tmp2 = c_make(et[i2] + q2.re, q2.im);
            FY[i1] = c_rdiv(c_sub(c_rmul(ct, tmp1), c_rmul(st, tmp2)), b);        /* :182 */
            FY[i2] = c_rdiv(c_add(c_rmul(st, tmp1), c_rmul(ct, tmp2)), b);        /* :183 */
        }
        for (int comp = 0; comp < 2; comp++) {
            cplx *p1 = FX + (size_t)N * (comp + 2 * m), *p2 = FY + (size_t)N * (comp + 2 * m);
            orc_fft_exec(plan, p1, 1, p1, 1, -1);                                 /* :187-188 */
            orc_fft_exec(plan, p2, 1, p2, 1, -1);                                 /* :189-190 */
            if (normalise)
                for (int n = 0; n < N; n++) { p1[n] = c_rdiv(p1[n], (double)N); p2[n] = c_rdiv(p2[n], (double)N); } /* :194-195 */
        }
    }
    orc_fft_free(plan);
    }
}

/* ------------------------------------------------------------------------------------ */
/* ua_step1 (Fortran form: FFT, exponential Euler, inverse FFT)   fortran/ua_steps.F90:200-236 */
/* ------------------------------------------------------------------------------------ */

void orc_ua_step1(int ntau, double eps, int64_t np, const double *t_, const double *pl_, double *xt_, double *xf_,
                  const double *fx_)
{
    const int N = ntau;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    orc_ua_tables(N, tau, ltau);
    const cplx *PL = (const cplx *)pl_, *FX = (const cplx *)fx_;
    cplx *XT = (cplx *)xt_, *XF = (cplx *)xf_;
    #pragma omp parallel if (g_threads > 1)
    {
    orc_fft_plan *plan = orc_fft_new(N);
    cplx rf[2][ORC_MAX_NTAU];
    #pragma omp for schedule(static)
    for (int64_t m = 0; m < np; m++) {
        cplx *x1 = XT + (size_t)N * (0 + 2 * m), *x2 = XT + (size_t)N * (1 + 2 * m);
        cplx *f1 = XF + (size_t)N * (0 + 2 * m), *f2 = XF + (size_t)N * (1 + 2 * m);
        orc_fft_exec(plan, x1, 1, f1, 1, -1);                                     /* :217 */
        orc_fft_exec(plan, x2, 1, f2, 1, -1);                                     /* :218 */
        double t = t_[m];                                                         /* :220 */
        for (int n = 0; n < N; n++) {
            /* elt = exp(-(0,1)*ltau(n)*t/eps) / ntau     :224-225 */
            cplx elt = c_expi(((-1.0 * ltau[n]) * t) / eps);
            elt = c_rdiv(elt, (double)N);
            cplx pl = PL[n + (size_t)N * m];
            rf[0][n] = c_add(c_mul(elt, f1[n]), c_mul(pl, FX[n + (size_t)N * (0 + 2 * m)]));   /* :226 */
            rf[1][n] = c_add(c_mul(elt, f2[n]), c_mul(pl, FX[n + (size_t)N * (1 + 2 * m)]));   /* :227 */
        }
        orc_fft_exec(plan, rf[0], 1, x1, 1, +1);                                  /* :231 */
        orc_fft_exec(plan, rf[1], 1, x2, 1, +1);                                  /* :232 */
    }
    orc_fft_free(plan);
    }
}

/* ------------------------------------------------------------------------------------ */
/* ua_step2 (corrector)              fortran/ua_steps.F90:238-272                          */
/* ------------------------------------------------------------------------------------ */

void orc_ua_step2(int ntau, double eps, int64_t np, const double *t_, const double *pl_, const double *ql_,
                  double *xt_, const double *xf_, const double *fx_, const double *gx_)
{
    const int N = ntau;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    orc_ua_tables(N, tau, ltau);
    const cplx *PL = (const cplx *)pl_, *QL = (const cplx *)ql_, *FX = (const cplx *)fx_, *GX = (const cplx *)gx_;
    const cplx *XF = (const cplx *)xf_;
    cplx *XT = (cplx *)xt_;
    #pragma omp parallel if (g_threads > 1)
    {
    orc_fft_plan *plan = orc_fft_new(N);
    cplx rf[2][ORC_MAX_NTAU];
    #pragma omp for schedule(static)
    for (int64_t m = 0; m < np; m++) {
        double t = t_[m];                                                         /* :254 */
        for (int n = 0; n < N; n++) {
            cplx elt = c_expi(((-1.0 * ltau[n]) * t) / eps);                      /* :258 */
            elt = c_rdiv(elt, (double)N);                                         /* :259 */
            cplx pl = PL[n + (size_t)N * m], ql = QL[n + (size_t)N * m];
            for (int comp = 0; comp < 2; comp++) {
                size_t i = n + (size_t)N * (comp + 2 * m);
                /* elt*xf + pl*fx + ql*(gx-fx)/t     :260-263 */
                cplx a = c_add(c_mul(elt, XF[i]), c_mul(pl, FX[i]));
                cplx c = c_rdiv(c_mul(ql, c_sub(GX[i], FX[i])), t);
                rf[comp][n] = c_add(a, c);
            }
        }
        orc_fft_exec(plan, rf[0], 1, XT + (size_t)N * (0 + 2 * m), 1, +1);        /* :267 */
        orc_fft_exec(plan, rf[1], 1, XT + (size_t)N * (1 + 2 * m), 1, +1);        /* :268 */
    }
    orc_fft_free(plan);
    }
}

/* ------------------------------------------------------------------------------------ */
/* compute_v                         fortran/ua_steps.F90:274-307                          */
/* ------------------------------------------------------------------------------------ */

void orc_compute_v(int ntau, double eps, int64_t np, const double *t_, const double *yt_, double *yf_, double *v)
{
    const int N = ntau;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU];
    orc_ua_tables(N, tau, ltau);
    const cplx *YT = (const cplx *)yt_;
    cplx *YF = (cplx *)yf_;
    #pragma omp parallel if (g_threads > 1)
    {
    orc_fft_plan *plan = orc_fft_new(N);
    #pragma omp for schedule(static)
    for (int64_t m = 0; m < np; m++) {
        double t = t_[m];                                                         /* :288 */
        cplx *f1 = YF + (size_t)N * (0 + 2 * m), *f2 = YF + (size_t)N * (1 + 2 * m);
        orc_fft_exec(plan, YT + (size_t)N * (0 + 2 * m), 1, f1, 1, -1);           /* :290 */
        orc_fft_exec(plan, YT + (size_t)N * (1 + 2 * m), 1, f2, 1, -1);           /* :291 */
        cplx px = c_make(0.0, 0.0), py = c_make(0.0, 0.0);                        /* :293-294 */
        for (int n = 0; n < N; n++) {
            cplx elt = c_expi((ltau[n] * t) / eps);                               /* :297 */
            px = c_add(px, c_mul(c_rdiv(f1[n], (double)N), elt));                 /* :298 */
            py = c_add(py, c_mul(c_rdiv(f2[n], (double)N), elt));                 /* :299 */
        }
        double c = cos(t / eps), s = sin(t / eps);
        v[2 * m]     = c * px.re + s * py.re;                                     /* :302 real(cos*px+sin*py) */
        v[2 * m + 1] = c * py.re - s * px.re;                                     /* :303 */
    }
    orc_fft_free(plan);
    }
}

/* ------------------------------------------------------------------------------------ */
/* The driver loop                   fortran/bupdate.F90:89-128                              */
/* energy[0] after the initial Poisson, then 2 per step (test/bupdate.jl:65,90,106).        */
/* sumv (optional): 2 per step = the values bupdate prints (:125).                          */
/* faithful=1 also performs the third (dead) interpolation of every step (:121).            */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    orc_mesh m; int ntau; double eps, dt, w; int64_t np; int wrap, faithful;
    double *x, *v;                     /* caller-owned (2,np) */
    double *rho, *emesh, *ep, *b, *t, *pl, *ql, *et, *xt, *xf, *yt, *yf, *fx, *fy, *gx, *gy;
} orc_sim;

void orc_sim_destroy(orc_sim *s)
{
    if (!s) return;
    free(s->rho); free(s->emesh); free(s->ep); free(s->b); free(s->t); free(s->pl); free(s->ql); free(s->et);
    free(s->xt); free(s->xf); free(s->yt); free(s->yf); free(s->fx); free(s->fy); free(s->gx); free(s->gy);
    free(s);
}

/* allocations of bupdate.F90:71-87 */
orc_sim *orc_sim_create(const orc_mesh *m, int ntau, double eps, double dt, int64_t np, double w, double *x, double *v,
                        int wrap, int faithful)
{
    if (ntau > ORC_MAX_NTAU) return NULL;
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    if (!s) return NULL;
    s->m = *m; s->ntau = ntau; s->eps = eps; s->dt = dt; s->w = w; s->np = np; s->wrap = wrap; s->faithful = faithful;
    s->x = x; s->v = v;
    const size_t nrho = (size_t)(m->nx + 1) * (size_t)(m->ny + 1);
    const size_t big = (size_t)ntau * 2 * (size_t)np;
    s->rho = (double *)calloc(nrho, sizeof(double));
    s->emesh = (double *)calloc(2 * nrho, sizeof(double));
    s->ep = (double *)calloc(2 * (size_t)np + 1, sizeof(double));
    s->b = (double *)malloc(sizeof(double) * ((size_t)np + 1));
    s->t = (double *)malloc(sizeof(double) * ((size_t)np + 1));
    s->pl = (double *)malloc(sizeof(cplx) * ((size_t)ntau * (size_t)np + 1));
    s->ql = (double *)malloc(sizeof(cplx) * ((size_t)ntau * (size_t)np + 1));
    s->et = (double *)malloc(sizeof(double) * (big + 1));
    s->xt = (double *)malloc(sizeof(cplx) * (big + 1)); s->xf = (double *)malloc(sizeof(cplx) * (big + 1));
    s->yt = (double *)malloc(sizeof(cplx) * (big + 1)); s->yf = (double *)malloc(sizeof(cplx) * (big + 1));
    s->fx = (double *)malloc(sizeof(cplx) * (big + 1)); s->fy = (double *)malloc(sizeof(cplx) * (big + 1));
    s->gx = (double *)malloc(sizeof(cplx) * (big + 1)); s->gy = (double *)malloc(sizeof(cplx) * (big + 1));
    if (!s->rho || !s->emesh || !s->ep || !s->b || !s->t || !s->pl || !s->ql || !s->et || !s->xt || !s->xf || !s->yt ||
        !s->yf || !s->fx || !s->fy || !s->gx || !s->gy) { orc_sim_destroy(s); return NULL; }
    return s;
}

/* bupdate.F90:89-93 ; returns the electric energy of the initial field */
double orc_sim_init(orc_sim *s)
{
    orc_compute_rho_m6(&s->m, s->np, s->x, s->w, s->rho, s->wrap);                /* :89 */
    double nrj = orc_poisson(&s->m, s->rho, s->emesh);                             /* :91 */
    orc_interpol_eb_m6(&s->m, s->emesh, s->np, s->x, s->ep, s->wrap);             /* :93 */
    return nrj;
}

/* one pass of the loop body bupdate.F90:97-125 ; energy2 gets the two Poisson energies, sumv2 (optional) what :125 prints */
void orc_sim_step(orc_sim *s, double *energy2, double *sumv2)
{
    const orc_mesh *m = &s->m;
    const int N = s->ntau; const double eps = s->eps; const int64_t np = s->np; const int wrap = s->wrap;
    orc_preparation(N, eps, s->dt, np, s->x, s->v, s->ep, s->b, s->t, s->pl, s->ql, s->xt, s->yt);   /* :97 */
    orc_interpol_eb_m6_tau(m, s->emesh, N, np, s->xt, s->et, wrap);               /* :99 */
    orc_compute_f(N, eps, np, s->b, s->xt, s->yt, s->et, s->fx, s->fy, 1);        /* :101 */
    orc_ua_step1(N, eps, np, s->t, s->pl, s->xt, s->xf, s->fx);                   /* :103 */
    orc_ua_step1(N, eps, np, s->t, s->pl, s->yt, s->yf, s->fy);                   /* :104 */
    orc_compute_rho_m6_tau(m, N, eps, np, s->xt, s->t, s->w, s->rho, s->x, wrap); /* :106 */
    energy2[0] = orc_poisson(m, s->rho, s->emesh);                                /* :108 */
    orc_interpol_eb_m6_tau(m, s->emesh, N, np, s->xt, s->et, wrap);               /* :110 */
    orc_compute_f(N, eps, np, s->b, s->xt, s->yt, s->et, s->gx, s->gy, 1);        /* :112 */
    orc_ua_step2(N, eps, np, s->t, s->pl, s->ql, s->xt, s->xf, s->fx, s->gx);     /* :114 */
    orc_ua_step2(N, eps, np, s->t, s->pl, s->ql, s->yt, s->yf, s->fy, s->gy);     /* :115 */
    orc_compute_rho_m6_tau(m, N, eps, np, s->xt, s->t, s->w, s->rho, s->x, wrap); /* :117 */
    energy2[1] = orc_poisson(m, s->rho, s->emesh);                                /* :119 */
    if (s->faithful) orc_interpol_eb_m6_tau(m, s->emesh, N, np, s->xt, s->et, wrap);   /* :121 (result never read) */
    orc_compute_v(N, eps, np, s->t, s->yt, s->yf, s->v);                          /* :123 */
    if (sumv2) {
        double sx = 0.0, sy = 0.0;
        for (int64_t k = 0; k < np; k++) { sx += s->v[2 * k]; sy += s->v[2 * k + 1]; }   /* :125 */
        sumv2[0] = sx; sumv2[1] = sy;
    }
}

void orc_sim_get(const orc_sim *s, double *e_part_out, double *emesh_out)
{
    const size_t nrho = (size_t)(s->m.nx + 1) * (size_t)(s->m.ny + 1);
    if (e_part_out) memcpy(e_part_out, s->ep, sizeof(double) * 2 * (size_t)s->np);
    if (emesh_out) memcpy(emesh_out, s->emesh, sizeof(double) * 2 * nrho);
}

int orc_run_bupdate(const orc_mesh *m, int ntau, double eps, double dt, int nstep, int64_t np, double w,
                    double *x, double *v, double *e_part_out, double *emesh_out, double *energy, double *sumv,
                    int wrap, int faithful)
{
    orc_sim *s = orc_sim_create(m, ntau, eps, dt, np, w, x, v, wrap, faithful);
    if (!s) return -2;
    int ie = 0;
    energy[ie++] = orc_sim_init(s);
    for (int istep = 0; istep < nstep; istep++) {
        orc_sim_step(s, energy + ie, sumv ? sumv + 2 * istep : NULL);
        ie += 2;
    }
    orc_sim_get(s, e_part_out, emesh_out);
    orc_sim_destroy(s);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* init_particles_2d densities      fortran/particles.F90:68-103, src/plasma.jl:17-48       */
/* The uniform deviates are supplied by the caller (u must hold enough draws); returns the   */
/* number of deviates consumed, or -1 if u ran out.  Keeps the reference's draw order.       */
/* ------------------------------------------------------------------------------------ */

int64_t orc_plasma_from_uniforms(const orc_mesh *m, int64_t np, double alpha, double kx, const double *u, int64_t nu,
                                 double *x, double *v)
{
    const double dimx = m->xmax - m->xmin, dimy = m->ymax - m->ymin;
    int64_t iu = 0, k = 0;
    while (k < np) {
        if (iu + 3 > nu) return -1;
        double xi = u[iu++] * dimx;
        double yi = u[iu++] * dimy;
        double zi = (2.0 + alpha) * u[iu++];
        double temm = 1.0 + sin(yi) + alpha * cos(kx * xi);
        if (temm >= zi) { x[2 * k] = xi; x[2 * k + 1] = yi; k++; }
    }
    k = 0;
    while (k < np) {
        if (iu + 3 > nu) return -1;
        double xi = (u[iu++] - 0.5) * 10.0;
        double yi = (u[iu++] - 0.5) * 10.0;
        double zi = u[iu++];
        double temm = (exp(-((xi - 2.0) * (xi - 2.0) + yi * yi) / 2.0) + exp(-((xi + 2.0) * (xi + 2.0) + yi * yi) / 2.0)) / 2.0;
        if (temm >= zi) { v[2 * k] = xi; v[2 * k + 1] = yi; k++; }
    }
    return iu;
}

/* ------------------------------------------------------------------------------------ */
/* Counter-based synthetic loads of the BENCHMARK configs (*/
/* "counter-based RNG keyed by (seed, particle index)").  Not reference code: the reference  */
/* draws from the Fortran RANDOM_NUMBER stream (particles.F90:68-103) or from Sobol points   */
/* (src/landau.jl:19-43), neither of which can hand a shard its slice.  Stated here so that  */
/* the CPU arm of bench.py and the parity tests run on the SAME particles as the device      */
/* generator k_generate (uapic.jl_b200/csrc/uapic_kernels.cu): same hash, same draw order,   */
/* same densities (kind 0: particles.F90:68-103; kind 1: the intent of src/landau.jl:19-43). */
/* global particle id = first + k*stride.                                                    */
/* ------------------------------------------------------------------------------------ */
static uint64_t orc_splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double orc_uniform01(uint64_t seed, uint64_t particle, uint32_t stream, uint32_t draw)
{
    uint64_t h = orc_splitmix64(seed ^ orc_splitmix64(particle * 0xD1342543DE82EF95ull + stream));
    h = orc_splitmix64(h + draw);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

void orc_generate(const orc_mesh *m, int kind, uint64_t seed, int64_t first, int64_t stride, int64_t np, int64_t np_global,
                  double alpha, double kx, double *x, double *v)
{
    const double dimx = m->xmax - m->xmin, dimy = m->ymax - m->ymin;
    const double pi = 3.14159265358979323846;
    #pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t k = 0; k < np; ++k) {
        const uint64_t id = (uint64_t)(first + k * stride);
        double x1, x2, v1, v2;
        if (kind == 0) {
            uint32_t d = 0;
            for (;;) {
                const double xi = orc_uniform01(seed, id, 0, d) * dimx;
                const double yi = orc_uniform01(seed, id, 0, d + 1) * dimy;
                const double zi = (2.0 + alpha) * orc_uniform01(seed, id, 0, d + 2);
                d += 3;
                if (1.0 + sin(yi) + alpha * cos(kx * xi) >= zi) { x1 = xi; x2 = yi; break; }
            }
            d = 0;
            for (;;) {
                const double xi = (orc_uniform01(seed, id, 1, d) - 0.5) * 10.0;
                const double yi = (orc_uniform01(seed, id, 1, d + 1) - 0.5) * 10.0;
                const double zi = orc_uniform01(seed, id, 1, d + 2);
                d += 3;
                const double temm = (exp(-((xi - 2.0) * (xi - 2.0) + yi * yi) / 2.0) + exp(-((xi + 2.0) * (xi + 2.0) + yi * yi) / 2.0)) / 2.0;
                if (temm >= zi) { v1 = xi; v2 = yi; break; }
            }
            x1 += m->xmin; x2 += m->ymin;
        } else {
            const double r1 = orc_uniform01(seed, id, 2, 0), r2 = orc_uniform01(seed, id, 2, 1), r3 = orc_uniform01(seed, id, 2, 2);
            const double target = r2 * (2.0 * pi / kx);
            double x0 = target;
            for (int it = 0; it < 50; ++it) {
                const double pfun = x0 + alpha * sin(kx * x0) / kx;
                const double f = 1.0 + alpha * cos(kx * x0);
                const double xn = x0 - (pfun - target) / f;
                const int done = fabs(xn - x0) <= 1e-12;
                x0 = xn;
                if (done) break;
            }
            x1 = m->xmin + x0;
            x2 = m->ymin + r3 * dimy;
            const double vv = sqrt(-2.0 * log(((double)id + 0.5) / (double)np_global));
            v1 = vv * cos(2.0 * pi * r1); v2 = vv * sin(2.0 * pi * r1);
        }
        x[2 * k] = x1; x[2 * k + 1] = x2;
        v[2 * k] = v1; v[2 * k + 1] = v2;
    }
}

/* ==================================================================================== */
/* Sibling scheme: the 3D rotation-push PIC of fortran/uapic3d.f90 */
/* Statement-by-statement restatement, quirks included (see uapic_mrc3d.cu's header).    */
/* Arrays: x, v, e_part (3,np); rho (nx+1,ny+1,nz+1); e (3,nx+1,ny+1,nz+1), x fastest.    */
/* ==================================================================================== */

typedef struct { double xmin[3], xmax[3]; int32_t n[3]; } orc3_mesh;

#define ORC3_NODE(i, j, k) ((size_t)(i) + (size_t)(nx + 1) * ((size_t)(j) + (size_t)(ny + 1) * (size_t)(k)))

/* compute_rho_cic                                   fortran/compute_rho_cic.f90:11-79 */
void orc3_compute_rho_cic(const orc3_mesh *m, int64_t np, const double *x, double w, double *rho)
{
    const int nx = m->n[0], ny = m->n[1], nz = m->n[2];
    const double dx = (m->xmax[0] - m->xmin[0]) / (double)nx, dy = (m->xmax[1] - m->xmin[1]) / (double)ny, dz = (m->xmax[2] - m->xmin[2]) / (double)nz;
    const size_t nn = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    for (size_t q = 0; q < nn; q++) rho[q] = 0.0;                                   /* :28 */
    const double vol = w / (dx * dy * dz);                                          /* :30 */
    for (int64_t p = 0; p < np; p++) {
        const double xp = x[3 * p] / dx, yp = x[3 * p + 1] / dy, zp = x[3 * p + 2] / dz;   /* :34-36 */
        int ip = (int)floor(xp), jp = (int)floor(yp), kp = (int)floor(zp);
        const double dxp = xp - (double)ip, dyp = yp - (double)jp, dzp = zp - (double)kp;
        const double a1 = (1.0 - dxp) * (1.0 - dyp) * (1.0 - dzp), a2 = dxp * (1.0 - dyp) * (1.0 - dzp);
        const double a3 = (1.0 - dxp) * dyp * (1.0 - dzp), a4 = dxp * dyp * (1.0 - dzp);
        const double a5 = (1.0 - dxp) * (1.0 - dyp) * dzp, a6 = dxp * (1.0 - dyp) * dzp;
        const double a7 = (1.0 - dxp) * dyp * dzp, a8 = dxp * dyp * dzp;
        /* outside the box the reference is undefined; clamp the cell like the CUDA kernel does (memory safety only) */
        if (ip < 0) ip = 0;
        if (ip >= nx) ip = nx - 1;
        if (jp < 0) jp = 0;
        if (jp >= ny) jp = ny - 1;
        if (kp < 0) kp = 0;
        if (kp >= nz) kp = nz - 1;
        rho[ORC3_NODE(ip, jp, kp)] += a1 * vol;          rho[ORC3_NODE(ip + 1, jp, kp)] += a2 * vol;          /* :58-65 */
        rho[ORC3_NODE(ip, jp + 1, kp)] += a3 * vol;      rho[ORC3_NODE(ip + 1, jp + 1, kp)] += a4 * vol;
        rho[ORC3_NODE(ip, jp, kp + 1)] += a5 * vol;      rho[ORC3_NODE(ip + 1, jp, kp + 1)] += a6 * vol;
        rho[ORC3_NODE(ip, jp + 1, kp + 1)] += a7 * vol;  rho[ORC3_NODE(ip + 1, jp + 1, kp + 1)] += a8 * vol;
    }
    for (int k = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) rho[ORC3_NODE(nx, j, k)] = rho[ORC3_NODE(0, j, k)];   /* :69 */
    for (int k = 0; k <= nz; k++) for (int i = 0; i <= nx; i++) rho[ORC3_NODE(i, ny, k)] = rho[ORC3_NODE(i, 0, k)];   /* :70 */
    for (int j = 0; j <= ny; j++) for (int i = 0; i <= nx; i++) rho[ORC3_NODE(i, j, nz)] = rho[ORC3_NODE(i, j, 0)];   /* :71 */
}

/* interpolate_eb_cic                                fortran/interpolation_cic.f90:10-66 */
void orc3_interpolate_eb_cic(const orc3_mesh *m, const double *e, int64_t np, const double *x, double *ep)
{
    const int nx = m->n[0], ny = m->n[1], nz = m->n[2];
    const double dx = (m->xmax[0] - m->xmin[0]) / (double)nx, dy = (m->xmax[1] - m->xmin[1]) / (double)ny, dz = (m->xmax[2] - m->xmin[2]) / (double)nz;
    for (int64_t p = 0; p < np; p++) {
        const double xp = x[3 * p] / dx, yp = x[3 * p + 1] / dy, zp = x[3 * p + 2] / dz;
        int i = (int)floor(xp), j = (int)floor(yp), k = (int)floor(zp);
        const double dxp = xp - (double)i, dyp = yp - (double)j, dzp = zp - (double)k;
        const double a1 = (1.0 - dxp) * (1.0 - dyp) * (1.0 - dzp), a2 = dxp * (1.0 - dyp) * (1.0 - dzp);
        const double a3 = (1.0 - dxp) * dyp * (1.0 - dzp), a4 = dxp * dyp * (1.0 - dzp);
        const double a5 = (1.0 - dxp) * (1.0 - dyp) * dzp, a6 = dxp * (1.0 - dyp) * dzp;
        const double a7 = (1.0 - dxp) * dyp * dzp, a8 = dxp * dyp * dzp;
        if (i < 0) i = 0;
        if (i >= nx) i = nx - 1;
        if (j < 0) j = 0;
        if (j >= ny) j = ny - 1;
        if (k < 0) k = 0;
        if (k >= nz) k = nz - 1;
        for (int d = 0; d < 3; d++)                                                  /* :56-60 */
            ep[3 * p + d] = a1 * e[d + 3 * ORC3_NODE(i, j, k)] + a2 * e[d + 3 * ORC3_NODE(i + 1, j, k)]
                          + a3 * e[d + 3 * ORC3_NODE(i, j + 1, k)] + a4 * e[d + 3 * ORC3_NODE(i + 1, j + 1, k)]
                          + a5 * e[d + 3 * ORC3_NODE(i, j, k + 1)] + a6 * e[d + 3 * ORC3_NODE(i + 1, j, k + 1)]
                          + a7 * e[d + 3 * ORC3_NODE(i, j + 1, k + 1)] + a8 * e[d + 3 * ORC3_NODE(i + 1, j + 1, k + 1)];
    }
}

static void orc3_fft3(cplx *a, int nx, int ny, int nz, int sign, orc_fft_plan *px, orc_fft_plan *py, orc_fft_plan *pz)
{
    for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) { cplx *l = a + (size_t)nx * (j + (size_t)ny * k); orc_fft_exec(px, l, 1, l, 1, sign); }
    for (int k = 0; k < nz; k++) for (int i = 0; i < nx; i++) { cplx *l = a + i + (size_t)nx * ny * k; orc_fft_exec(py, l, nx, l, nx, sign); }
    for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) { cplx *l = a + i + (size_t)nx * j; orc_fft_exec(pz, l, nx * ny, l, nx * ny, sign); }
}

/* solve_poisson                                     fortran/poisson_3d.f90:47-191 */
void orc3_poisson(const orc3_mesh *m, const double *rho, double *e)
{
    const int nx = m->n[0], ny = m->n[1], nz = m->n[2];
    const double pi = 4.0 * atan(1.0);
    const double dimx = m->xmax[0] - m->xmin[0], dimy = m->xmax[1] - m->xmin[1], dimz = m->xmax[2] - m->xmin[2];
    const size_t nc = (size_t)nx * ny * nz;
    cplx *rhs = (cplx *)malloc(sizeof(cplx) * nc), *psi = (cplx *)malloc(sizeof(cplx) * nc);
    orc_fft_plan *px = orc_fft_new(nx), *py = orc_fft_new(ny), *pz = orc_fft_new(nz);
    for (int comp = 0; comp < 3; comp++) {
        for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++)
            rhs[i + (size_t)nx * (j + (size_t)ny * k)] = c_make(rho[ORC3_NODE(i, j, k)], 0.0);              /* :71, :98, :131 */
        orc3_fft3(rhs, nx, ny, nz, -1, px, py, pz);
        for (int k = 0; k < nz; k++) {
            const int ind_z = (k + 1 <= nz / 2) ? k : -nz + k;                                              /* :75-79 */
            const double kz = 2.0 * pi * (double)ind_z / dimz;
            for (int j = 0; j < ny; j++) {
                const int ind_y = (j + 1 <= ny / 2) ? j : -ny + j;
                const double ky = 2.0 * pi * (double)ind_y / dimy;
                for (int i = 0; i < nx; i++) {
                    const int ind_x = (i + 1 <= nx / 2) ? i : -nx + i;
                    const double kx = 2.0 * pi * (double)ind_x / dimx;
                    cplx *r = &rhs[i + (size_t)nx * (j + (size_t)ny * k)];
                    if (ind_x == 0 && ind_y == 0 && ind_z == 0) { *r = c_make(0.0, 0.0); continue; }
                    const double kk = comp == 0 ? kx : (comp == 1 ? ky : kz), k2 = kx * kx + ky * ky + kz * kz;
                    const cplx t = c_mul(c_make(-0.0, -kk), *r);                                             /* -cmplx(0,kk) * rhs */
                    *r = c_make(t.re / k2, t.im / k2);
                }
            }
        }
        for (size_t q = 0; q < nc; q++) psi[q] = rhs[q];
        orc3_fft3(psi, nx, ny, nz, +1, px, py, pz);
        for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++)
            e[comp + 3 * ORC3_NODE(i, j, k)] = psi[i + (size_t)nx * (j + (size_t)ny * k)].re;                /* :96, :129, :163 */
    }
    for (int d = 0; d < 3; d++) {                                                                             /* :184-186 */
        for (int k = 0; k <= nz; k++) for (int j = 0; j <= ny; j++) e[d + 3 * ORC3_NODE(nx, j, k)] = e[d + 3 * ORC3_NODE(0, j, k)];
    }
    for (int d = 0; d < 3; d++) for (int k = 0; k <= nz; k++) for (int i = 0; i <= nx; i++) e[d + 3 * ORC3_NODE(i, ny, k)] = e[d + 3 * ORC3_NODE(i, 0, k)];
    for (int d = 0; d < 3; d++) for (int j = 0; j <= ny; j++) for (int i = 0; i <= nx; i++) e[d + 3 * ORC3_NODE(i, j, nz)] = e[d + 3 * ORC3_NODE(i, j, 0)];
    const size_t nn = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    const double sc = (double)nx * (double)ny * (double)nz;
    for (size_t q = 0; q < 3 * nn; q++) e[q] = e[q] / sc;                                                     /* :188 */
    orc_fft_free(px); orc_fft_free(py); orc_fft_free(pz);
    free(rhs); free(psi);
}

static double orc3_modulo(double a, double p) { double r = fmod(a, p); if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p; return r; }

static void orc3_push(const orc3_mesh *m, int64_t np, double *x, const double *v, double delta_t)          /* uapic3d.f90:224-242 */
{
    for (int64_t p = 0; p < np; p++)
        for (int c = 0; c < 3; c++) {
            const double d = delta_t * v[3 * p + c] - m->xmin[c];
            x[3 * p + c] = m->xmin[c] + orc3_modulo(x[3 * p + c] + d, m->xmax[c] - m->xmin[c]);
        }
}

static void orc3_rotate(int kind, int64_t np, const double *x, double *v, const double *ep, double dt, double eps, double coef, double delta,
                        int index_quirk)
{
    for (int64_t p = 0; p < np; p++) {
        const double x1 = x[3 * p], x2 = x[3 * p + 1];
        const double xa = (kind == 2 && index_quirk) ? x[p] : x1;                   /* p%x(m,1), uapic3d.f90:179,182 */
        double Bm[3], Ee[3], vv[3], vxB[3], ExB[3];
        Bm[0] = (x2 - 9.0) * delta / sqrt(1.0 + ((xa - 9.0) * (xa - 9.0) + (x2 - 9.0) * (x2 - 9.0)) * delta * delta);
        Bm[1] = -(x1 - 9.0) * delta / sqrt(1.0 + ((xa - 9.0) * (xa - 9.0) + (x2 - 9.0) * (x2 - 9.0)) * delta * delta);
        Bm[2] = 1.0 / sqrt(1.0 + ((x1 - 9.0) * (x1 - 9.0) + (x2 - 9.0) * (x2 - 9.0)) * delta * delta);
        for (int c = 0; c < 3; c++) { Ee[c] = ep[3 * p + c]; vv[c] = v[3 * p + c]; }
        vxB[0] = vv[1] * Bm[2] - vv[2] * Bm[1]; vxB[1] = vv[2] * Bm[0] - vv[0] * Bm[2]; vxB[2] = vv[0] * Bm[1] - vv[1] * Bm[0];
        ExB[0] = Ee[1] * Bm[2] - Ee[2] * Bm[1]; ExB[1] = Ee[2] * Bm[0] - Ee[0] * Bm[2]; ExB[2] = Ee[0] * Bm[1] - Ee[1] * Bm[0];
        const double EB = Bm[0] * Ee[0] + Bm[1] * Ee[1] + Bm[2] * Ee[2], vB = Bm[0] * vv[0] + Bm[1] * vv[1] + Bm[2] * vv[2];
        for (int c = 0; c < 3; c++) {
            if (kind == 0)          /* :113-121 */
                v[3 * p + c] = cos(dt / eps) * vv[c] + sin(dt / eps) * vxB[c] + eps * sin(dt / eps) * Ee[c] + (dt - eps * sin(dt / eps)) * EB * Bm[c]
                             + (eps - eps * cos(dt / eps)) * ExB[c] + (1.0 - cos(dt / eps)) * vB * Bm[c];
            else if (kind == 1)     /* :152-159 */
                v[3 * p + c] = cos(dt) * vv[c] + sin(dt) * vxB[c] + coef * sin(dt) * Ee[c] + coef * (dt - sin(dt)) * EB * Bm[c]
                             + coef * (1.0 - cos(dt)) * ExB[c] + (1.0 - cos(dt)) * vB * Bm[c];
            else                    /* :187-194 */
                v[3 * p + c] = cos(dt) * vv[c] + sin(-dt) * vxB[c] - coef * sin(-dt) * Ee[c] - coef * (-dt - sin(-dt)) * EB * Bm[c]
                             - coef * (1.0 - cos(dt)) * ExB[c] + (1.0 - cos(dt)) * vB * Bm[c];
        }
    }
}

/* the program: uapic3d.f90:44-61 (parameters), :74-83 (initial fields), :91-206 (time loop).  max_outer > 0 caps the outer loop. */
int64_t orc3_run(const orc3_mesh *m, int64_t np, double *x, double *v, double *ep_out, double w, double eps, double delta, int nmrc, int nmrcm,
                 double tfinal, int max_outer, int index_quirk, double *e_out, double *rho_out)
{
    const int nx = m->n[0], ny = m->n[1], nz = m->n[2];
    const size_t nn = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    const double pi = 4.0 * atan(1.0);
    double *rho = (double *)malloc(8 * nn), *e = (double *)malloc(24 * nn), *ep = (double *)malloc(24 * (size_t)(np ? np : 1));
    int64_t done = 0;
#define ORC3_FIELDS() do { orc3_compute_rho_cic(m, np, x, w, rho); orc3_poisson(m, rho, e); orc3_interpolate_eb_cic(m, e, np, x, ep); } while (0)
    ORC3_FIELDS();                                                                   /* :76-83 */
    const long n0 = lround(tfinal / eps / (2.0 * pi) / (double)nmrc);                /* :48 */
    if (n0 == 0 || n0 == 1) {
        const double dt = eps * (2.0 * pi) / (double)nmrc;
        long nstep = lround(tfinal / dt);
        if (max_outer > 0 && nstep > max_outer) nstep = max_outer;
        for (long it = 0; it < nstep; it++) {
            orc3_push(m, np, x, v, 0.5 * dt);
            ORC3_FIELDS();
            orc3_rotate(0, np, x, v, ep, dt, eps, 1.0, delta, index_quirk);
            for (int64_t q = 0; q < 3 * np; q++) x[q] = x[q] + 0.5 * dt * v[q];      /* :123 (no wrap) */
            done++;
        }
    } else {
        const double alpha = 0.5 * (1.0 + 1.0 / (double)n0) * eps * (double)n0, beta = 0.5 * (1.0 - 1.0 / (double)n0) * eps * (double)n0;
        const double dt = (2.0 * pi) / (double)nmrcm;
        int outer = nmrc;
        if (max_outer > 0 && outer > max_outer) outer = max_outer;
        for (int istep = 0; istep < outer; istep++) {
            for (int n = 0; n < nmrcm; n++) {
                orc3_push(m, np, x, v, 0.5 * dt * alpha);
                ORC3_FIELDS();
                orc3_rotate(1, np, x, v, ep, dt, eps, alpha, delta, index_quirk);
                orc3_push(m, np, x, v, 0.5 * dt * alpha);
                done++;
            }
            for (int n = 0; n < nmrcm; n++) {
                orc3_push(m, np, x, v, 0.5 * dt * beta);
                ORC3_FIELDS();
                orc3_rotate(2, np, x, v, ep, dt, eps, beta, delta, index_quirk);
                orc3_push(m, np, x, v, 0.5 * dt * beta);
                done++;
            }
        }
    }
    if (ep_out) memcpy(ep_out, ep, 24 * (size_t)np);
    if (e_out) memcpy(e_out, e, 24 * nn);
    if (rho_out) memcpy(rho_out, rho, 8 * nn);
    free(rho); free(e); free(ep);
    return done;
}

/* init_particles_3d densities (particles.F90:152-190) from the counter-based stream of k3_generate (uapic_mrc3d.cu) */
void orc3_generate(const orc3_mesh *m, uint64_t seed, int64_t first, int64_t np, double *x, double *v)
{
    const double pi = 3.14159265358979323846;
    for (int64_t p = 0; p < np; p++) {
        const uint64_t id = (uint64_t)(first + p);
        x[3 * p + 2] = m->xmin[2] + (m->xmax[2] - m->xmin[2]) * orc_uniform01(seed, id, 3, 0);
        for (uint32_t d = 0;; d += 3) {
            const double xi = 9.0 * orc_uniform01(seed, id, 4, d), yi = 2.0 * pi * orc_uniform01(seed, id, 4, d + 1), zi = (1.0 + 0.02) * orc_uniform01(seed, id, 4, d + 2);
            if ((1.0 + 0.02 * cos(4.0 * yi)) * exp(-5.0 * (xi - 4.8) * (xi - 4.8)) >= zi) { x[3 * p] = cos(yi) * xi + 9.0; x[3 * p + 1] = sin(yi) * xi + 9.0; break; }
        }
        for (uint32_t d = 0;; d += 4) {
            const double xi = (orc_uniform01(seed, id, 5, d) - 0.5) * 8.0, yi = (orc_uniform01(seed, id, 5, d + 1) - 0.5) * 8.0, wi = (orc_uniform01(seed, id, 5, d + 2) - 0.5) * 8.0;
            if (exp(-2.0 * (xi * xi + yi * yi + wi * wi)) >= orc_uniform01(seed, id, 5, d + 3)) { v[3 * p] = xi; v[3 * p + 1] = yi; v[3 * p + 2] = wi; break; }
        }
    }
}

/* ==================================================================================== */
/* External-field two-scale program      fortran/efd.f90 (test/test_efd.jl is its twin)  */
/* ==================================================================================== */
/*
 * Every particle is integrated on its own in the prescribed field
 *   E(x,t) = (cos(x1/2) sin(x2) / 2, sin(x1/2) cos(x2)) (1 + sin(t)/2),   b(x) = 1 + sin(x1) sin(x2) / 2
 * with a third-order prepared initial datum (efd.f90:133-383) and `nstep` second-order IMEX steps in tau-Fourier
 * space (efd.f90:388-454); the physical state is read off at tau = tfinal b / eps (efd.f90:456-478).  As shipped the
 * program handles ONE particle and stops after printing intermediate sums (efd.f90:131,257); this is the text behind
 * that `stop`, run over all particles, which is what produced the constants on efd.f90:481:
 * with init_particles_2d's load (libgfortran stream, seed of particles.F90:57-64), ntau = 16, eps = 1e-3,
 * dt = pi/16, tfinal = pi/2 it gives sum(v) = (-857.95049281063, -593.40700170710), the printed reference
 * values, to 13 digits -- PINNED by tests/test_efd_oracle.py.
 * Forward transforms carry 1/ntau (fft.f90:44-72).  Quantities the program computes and never uses (pl, ql, gx, ave2:
 * efd.f90:143-149,275,283) are left out.
 */
#include <complex.h>
typedef double complex zc;

typedef struct {
    int ntau;
    orc_fft_plan *plan;
    double tau[ORC_MAX_NTAU], ltau[ORC_MAX_NTAU], ct[ORC_MAX_NTAU], st[ORC_MAX_NTAU];
} efd_ctx;

static void efd_fft(const efd_ctx *c, const zc *in, zc *out)        /* fft.f90:61-72 */
{
    orc_fft_exec(c->plan, (const cplx *)in, 1, (cplx *)out, 1, -1);
    for (int n = 0; n < c->ntau; n++) out[n] /= (double)c->ntau;
}
static void efd_ifft(const efd_ctx *c, const zc *in, zc *out)       /* fft.f90:74-81 */
{
    orc_fft_exec(c->plan, (const cplx *)in, 1, (cplx *)out, 1, +1);
}
/* tilde(n) = -i tilde(n) / ltau(n), n >= 2 ; tilde(1) = 0 ; back-transform      (efd.f90:183-188 and its repeats) */
static void efd_primitive(const efd_ctx *c, const zc *in, zc *out)
{
    zc t[ORC_MAX_NTAU];
    efd_fft(c, in, t);
    for (int n = 1; n < c->ntau; n++) t[n] = -I * t[n] / c->ltau[n];
    t[0] = 0.0;
    efd_ifft(c, t, out);
}
static inline double efd_b(double a, double b) { return 1.0 + 0.5 * sin(a) * sin(b); }

/* efd.f90:509-524 */
static void efd_compute_fy(const efd_ctx *c, double eps, double bx, double time, const zc (*xt)[ORC_MAX_NTAU],
                           const zc (*yt)[ORC_MAX_NTAU], zc (*fy)[ORC_MAX_NTAU])
{
    for (int n = 0; n < c->ntau; n++) {
        double a = creal(xt[0][n]), b = creal(xt[1][n]);
        double e1 = (0.5 * cos(a / 2.0) * sin(b)) * (1.0 + 0.5 * sin(time));
        double e2 = (cos(b) * sin(a / 2.0)) * (1.0 + 0.5 * sin(time));
        double interv = (efd_b(a, b) - bx) / bx / eps;
        double t1 = (c->ct[n] * e1 - c->st[n] * e2) / bx;
        double t2 = (c->ct[n] * e2 + c->st[n] * e1) / bx;
        fy[0][n] = t1 + interv * yt[1][n];
        fy[1][n] = t2 - interv * yt[0][n];
    }
}

static void efd_one(const efd_ctx *c, double eps, double dt, double tfinal, int nstep, const double *box, double *xp, double *vp)
{
    const int nt = c->ntau;
    const double *ct = c->ct, *st = c->st;
    zc xt[2][ORC_MAX_NTAU], yt[2][ORC_MAX_NTAU], h[2][ORC_MAX_NTAU], r[2][ORC_MAX_NTAU], fx[2][ORC_MAX_NTAU], fy[2][ORC_MAX_NTAU];
    zc temp[2][ORC_MAX_NTAU], tilde[2][ORC_MAX_NTAU], xf[2][ORC_MAX_NTAU], yf[2][ORC_MAX_NTAU], gx[2][ORC_MAX_NTAU], gy[2][ORC_MAX_NTAU];
    const double x1 = xp[0], x2 = xp[1], v1 = vp[0], v2 = vp[1];
    double time = 0.0;
    const double bx = efd_b(x1, x2);                                    /* efd.f90:138-140 */
    const double ds = dt * bx;
    double ave[2], e1, e2, interv;

    /* first-order datum (efd.f90:157-195) */
    for (int n = 0; n < nt; n++) {
        h[0][n] = eps * (st[n] * (v1 / bx) - ct[n] * (v2 / bx));
        h[1][n] = eps * (st[n] * (v2 / bx) + ct[n] * (v1 / bx));
    }
    for (int n = 0; n < nt; n++) { xt[0][n] = x1 + h[0][n] - h[0][0]; xt[1][n] = x2 + h[1][n] - h[1][0]; }
    e1 = (0.5 * cos(x1 / 2.0) * sin(x2)) * (1.0 + 0.5 * sin(time));
    e2 = (sin(x1 / 2.0) * cos(x2)) * (1.0 + 0.5 * sin(time));
    for (int n = 0; n < nt; n++) {
        interv = (efd_b(creal(xt[0][n]), creal(xt[1][n])) - bx) / bx;
        r[0][n] = interv * v2;
        r[1][n] = -interv * v1;
    }
    efd_fft(c, r[0], tilde[0]); efd_fft(c, r[1], tilde[1]);
    ave[0] = creal(tilde[0][0]) / eps; ave[1] = creal(tilde[1][0]) / eps;
    for (int d = 0; d < 2; d++) {
        for (int n = 1; n < nt; n++) tilde[d][n] = -I * tilde[d][n] / c->ltau[n];
        tilde[d][0] = 0.0;
        efd_ifft(c, tilde[d], r[d]);
    }
    for (int n = 0; n < nt; n++) {
        r[0][n] = eps * (st[n] * e1 + ct[n] * e2) / bx + r[0][n];
        r[1][n] = eps * (st[n] * e2 - ct[n] * e1) / bx + r[1][n];
    }
    for (int n = 0; n < nt; n++) { yt[0][n] = v1 + (r[0][n] - r[0][0]); yt[1][n] = v2 + (r[1][n] - r[1][0]); }

    /* second-order position (efd.f90:200-222) */
    for (int n = 0; n < nt; n++) {
        temp[0][n] = eps * (ct[n] * yt[0][n] + st[n] * yt[1][n]) / bx;
        temp[1][n] = eps * (ct[n] * yt[1][n] - st[n] * yt[0][n]) / bx;
    }
    efd_primitive(c, temp[0], h[0]); efd_primitive(c, temp[1], h[1]);
    for (int n = 0; n < nt; n++) {
        h[0][n] = h[0][n] - eps * eps / bx * (-ct[n] * ave[0] - st[n] * ave[1]);
        h[1][n] = h[1][n] - eps * eps / bx * (-ct[n] * ave[1] + st[n] * ave[0]);
    }
    for (int n = 0; n < nt; n++) { xt[0][n] = x1 + h[0][n] - h[0][0]; xt[1][n] = x2 + h[1][n] - h[1][0]; }

    /* second-order velocity (efd.f90:226-310): the time derivative of E enters here */
    e1 = (0.5 * cos(x1 / 2.0) * sin(x2)) * 0.5 * cos(time);
    e2 = (sin(x1 / 2.0) * cos(x2)) * 0.5 * cos(time);
    for (int n = 0; n < nt; n++) {
        interv = (efd_b(creal(xt[0][n]), creal(xt[1][n])) - bx) / bx;
        fx[0][n] = interv * ave[1];
        fx[1][n] = -interv * ave[0];
        fy[0][n] = eps / bx * (st[n] * ave[0] - ct[n] * ave[1]);
        fy[1][n] = eps / bx * (ct[n] * ave[0] + st[n] * ave[1]);
        double w = creal(cos(x1) * sin(x2) * fy[0][n] + sin(x1) * cos(x2) * fy[1][n]);  /* `interv` is real(8) */
        fy[0][n] = w / bx / 2.0 * v2 + fx[0][n];
        fy[1][n] = -w / bx / 2.0 * v1 + fx[1][n];
        fx[0][n] = eps / (bx * bx) * (-st[n] * e2 + ct[n] * e1);
        fx[1][n] = eps / (bx * bx) * (st[n] * e1 + ct[n] * e2);
    }
    for (int d = 0; d < 2; d++) {
        for (int n = 0; n < nt; n++) temp[d][n] = fy[d][n] + fx[d][n];
        efd_fft(c, temp[d], tilde[d]);
        for (int n = 1; n < nt; n++) {
            fx[d][n] = -I * tilde[d][n] / c->ltau[n];
            tilde[d][n] = -tilde[d][n] / (c->ltau[n] * c->ltau[n]);
        }
        fx[d][0] = 0.0; tilde[d][0] = 0.0;
        efd_ifft(c, tilde[d], temp[d]);
        for (int n = 0; n < nt; n++) r[d][n] = -eps * temp[d][n];
        efd_ifft(c, fx[d], fy[d]);
    }
    for (int n = 0; n < nt; n++) {
        double a = creal(xt[0][n]), b = creal(xt[1][n]);
        e1 = (0.5 * cos(a / 2.0) * sin(b)) * (1.0 + 0.5 * sin(time));
        e2 = (sin(a / 2.0) * cos(b)) * (1.0 + 0.5 * sin(time));
        interv = (efd_b(a, b) - bx) / bx;
        temp[0][n] = interv * yt[1][n] + eps / bx * (-st[n] * e2 + ct[n] * e1);
        temp[1][n] = -interv * yt[0][n] + eps / bx * (st[n] * e1 + ct[n] * e2);
    }
    zc ydot[2];
    for (int d = 0; d < 2; d++) {
        efd_fft(c, temp[d], tilde[d]);
        ydot[d] = tilde[d][0] / eps;                                     /* xf(1,:), efd.f90:299 */
        for (int n = 1; n < nt; n++) tilde[d][n] = -I * tilde[d][n] / c->ltau[n];
        tilde[d][0] = 0.0;
        efd_ifft(c, tilde[d], temp[d]);
        for (int n = 0; n < nt; n++) r[d][n] = r[d][n] + temp[d][n];
    }
    for (int n = 0; n < nt; n++) { yt[0][n] = v1 + r[0][n] - r[0][0]; yt[1][n] = v2 + r[1][n] - r[1][0]; }

    /* third-order position (efd.f90:315-383) */
    for (int n = 0; n < nt; n++) {
        temp[0][n] = (ct[n] * r[0][n] + st[n] * r[1][n]) / bx;
        temp[1][n] = (ct[n] * r[1][n] - st[n] * r[0][n]) / bx;
    }
    efd_fft(c, temp[0], tilde[0]); efd_fft(c, temp[1], tilde[1]);
    double w0 = creal(cos(x1) * sin(x2) * tilde[0][0] + sin(x1) * cos(x2) * tilde[1][0]);   /* real(8) `interv` again */
    zc acc[2];
    acc[0] = w0 / eps / bx * v2 / 2.0;
    acc[1] = -w0 / eps / bx * v1 / 2.0;
    for (int n = 0; n < nt; n++) temp[0][n] = (efd_b(creal(xt[0][n]), creal(xt[1][n])) - bx) / bx;
    efd_fft(c, temp[0], tilde[0]);
    acc[0] = acc[0] + tilde[0][0] / eps * ave[1];
    acc[1] = acc[1] - tilde[0][0] / eps * ave[0];
    for (int n = 0; n < nt; n++) {
        yf[0][n] = ydot[0] + fy[0][n];
        yf[1][n] = ydot[1] + fy[1][n];
        temp[0][n] = ct[n] * yf[0][n] + st[n] * yf[1][n];
        temp[1][n] = ct[n] * yf[1][n] - st[n] * yf[0][n];
    }
    efd_primitive(c, temp[0], temp[0]); efd_primitive(c, temp[1], temp[1]);
    for (int n = 0; n < nt; n++) {
        fy[0][n] = temp[0][n] * eps / bx - eps * eps / bx * (-ct[n] * acc[0] - st[n] * acc[1]);
        fy[1][n] = temp[1][n] * eps / bx - eps * eps / bx * (-ct[n] * acc[1] + st[n] * acc[0]);
    }
    efd_primitive(c, fy[0], temp[0]); efd_primitive(c, fy[1], temp[1]);
    for (int d = 0; d < 2; d++) for (int n = 0; n < nt; n++) h[d][n] = -eps * temp[d][n];
    for (int n = 0; n < nt; n++) {
        temp[0][n] = eps * (ct[n] * yt[0][n] + st[n] * yt[1][n]) / bx;
        temp[1][n] = eps * (ct[n] * yt[1][n] - st[n] * yt[0][n]) / bx;
    }
    efd_primitive(c, temp[0], temp[0]); efd_primitive(c, temp[1], temp[1]);
    for (int d = 0; d < 2; d++) for (int n = 0; n < nt; n++) h[d][n] = h[d][n] + temp[d][n];
    for (int n = 0; n < nt; n++) { xt[0][n] = x1 + h[0][n] - h[0][0]; xt[1][n] = x2 + h[1][n] - h[1][0]; }

    /* IMEX2 steps (efd.f90:388-454) */
    for (int istep = 0; istep < nstep; istep++) {
        efd_compute_fy(c, eps, bx, time, xt, yt, fy);
        for (int d = 0; d < 2; d++) {
            for (int n = 0; n < nt; n++) gy[d][n] = yt[d][n] + ds / 2.0 * fy[d][n];
            efd_fft(c, gy[d], fy[d]);
            for (int n = 0; n < nt; n++) fy[d][n] = fy[d][n] / (1.0 + I * ds / 2.0 * c->ltau[n] / eps);
            efd_ifft(c, fy[d], yf[d]);
        }
        for (int n = 0; n < nt; n++) {
            fx[0][n] = (ct[n] * yf[0][n] + st[n] * yf[1][n]) / bx;
            fx[1][n] = (ct[n] * yf[1][n] - st[n] * yf[0][n]) / bx;
        }
        for (int d = 0; d < 2; d++) {
            for (int n = 0; n < nt; n++) gx[d][n] = xt[d][n] + ds / 2.0 * fx[d][n];
            efd_fft(c, gx[d], fx[d]);
            for (int n = 0; n < nt; n++) fx[d][n] = fx[d][n] / (1.0 + I * ds / 2.0 * c->ltau[n] / eps);
            efd_ifft(c, fx[d], xf[d]);
        }
        time = time + dt / 2.0;
        efd_compute_fy(c, eps, bx, time, xf, yf, fy);
        for (int d = 0; d < 2; d++) {
            efd_fft(c, fy[d], gy[d]);
            efd_fft(c, yt[d], yf[d]);
            for (int n = 0; n < nt; n++)
                fy[d][n] = (yf[d][n] * (1.0 - I * ds / eps / 2.0 * c->ltau[n]) + ds * gy[d][n]) / (1.0 + I * ds / 2.0 * c->ltau[n] / eps);
            for (int n = 0; n < nt; n++) yf[d][n] = yt[d][n];
            efd_ifft(c, fy[d], yt[d]);
            for (int n = 0; n < nt; n++) yf[d][n] = (yt[d][n] + yf[d][n]) / 2.0;
        }
        for (int n = 0; n < nt; n++) {
            fx[0][n] = (ct[n] * yf[0][n] + st[n] * yf[1][n]) / bx;
            fx[1][n] = (ct[n] * yf[1][n] - st[n] * yf[0][n]) / bx;
        }
        for (int d = 0; d < 2; d++) {
            efd_fft(c, fx[d], gx[d]);
            efd_fft(c, xt[d], xf[d]);
            for (int n = 0; n < nt; n++)
                fx[d][n] = (xf[d][n] * (1.0 - I * ds / eps / 2.0 * c->ltau[n]) + ds * gx[d][n]) / (1.0 + I * ds / 2.0 * c->ltau[n] / eps);
            efd_ifft(c, fx[d], xt[d]);
        }
        time = time + dt / 2.0;
    }

    /* physical state at tau = tfinal b / eps (efd.f90:456-478), positions wrapped as apply_bc (efd.f90:526-544) */
    zc s[2];
    for (int d = 0; d < 2; d++) {
        efd_fft(c, xt[d], tilde[d]);
        s[d] = 0.0;
        for (int n = 0; n < nt; n++) s[d] = s[d] + tilde[d][n] * cexp(I * c->ltau[n] * tfinal * bx / eps);
    }
    double xx = creal(s[0]), yy = creal(s[1]);
    const double dimx = box[1] - box[0], dimy = box[3] - box[2];
    while (xx > box[1]) xx -= dimx;
    while (xx < box[0]) xx += dimx;
    while (yy > box[3]) yy -= dimy;
    while (yy < box[2]) yy += dimy;
    xp[0] = xx; xp[1] = yy;
    for (int d = 0; d < 2; d++) {
        efd_fft(c, yt[d], tilde[d]);
        s[d] = 0.0;
        for (int n = 0; n < nt; n++) s[d] = s[d] + tilde[d][n] * cexp(I * c->ltau[n] * tfinal * bx / eps);
    }
    vp[0] = creal(cos(tfinal * bx / eps) * s[0] + sin(tfinal * bx / eps) * s[1]);
    vp[1] = creal(cos(tfinal * bx / eps) * s[1] - sin(tfinal * bx / eps) * s[0]);
}

/* x, v: (2, np), overwritten with the state at tfinal.  box = xmin, xmax, ymin, ymax.  nstep = nint(tfinal/dt) (efd.f90:102) */
int orc_efd_run(int ntau, int64_t np, double eps, double dt, double tfinal, const double *box, double *x, double *v)
{
    if (ntau < 2 || ntau > ORC_MAX_NTAU || (ntau & 1)) return -1;
    efd_ctx c;
    c.ntau = ntau;
    c.plan = orc_fft_new(ntau);
    const double pi = 4.0 * atan(1.0);
    const double dtau = 2.0 * pi / ntau;                                /* efd.f90:107,111-116 */
    for (int n = 0; n < ntau; n++) {
        c.tau[n] = n * dtau;
        c.ltau[n] = (n < ntau / 2) ? (double)n : (double)(n - ntau);
        c.ct[n] = cos(c.tau[n]);
        c.st[n] = sin(c.tau[n]);
    }
    const int nstep = (int)lround(tfinal / dt);
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (int64_t m = 0; m < np; m++) efd_one(&c, eps, dt, tfinal, nstep, box, x + 2 * m, v + 2 * m);
    orc_fft_free(c.plan);
    return nstep;
}

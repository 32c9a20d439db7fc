// uapic_capi.cu -- the C ABI declared in include/uapic_b200.h: host-side glue only (allocation, H2D/D2H,
// launch ordering).  No compute happens on the host and there is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>

#include "../../include/uapic_b200.h"
#include "uapic_internal.h"

using namespace uapic;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                              \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? UAPIC_ENOMEM : UAPIC_ECUDA, "%s failed: %s (%s:%d)", \
                        #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
    } while (0)

}  // namespace

int uapic_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

namespace {

struct DeviceInfo { int ok; int sm_count; int major, minor; };

int device_info(int device, DeviceInfo *out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(UAPIC_ENODEVICE, "no CUDA device available (%s); libuapic_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(UAPIC_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    // three attribute queries, remembered per device: cudaGetDeviceProperties fills ~100 fields on every call and every
    // stage-API entry point comes through here
    constexpr int kCached = 64;
    static std::mutex mu;
    static DeviceInfo cache[kCached] = {};
    {
        std::lock_guard<std::mutex> lock(mu);
        if (device < kCached && cache[device].ok) { *out = cache[device]; return UAPIC_OK; }
    }
    int major = 0, minor = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (major != 10)
        return fail(UAPIC_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
    out->ok = 1; out->sm_count = sms; out->major = major; out->minor = minor;
    if (device < kCached) {
        std::lock_guard<std::mutex> lock(mu);
        cache[device] = *out;
    }
    return UAPIC_OK;
}

int check_mesh(const uapic_mesh_t *mesh) {
    if (!mesh) return fail(UAPIC_EINVAL, "mesh is null");
    if (mesh->nx < 2 || mesh->ny < 2) return fail(UAPIC_EINVAL, "mesh needs nx, ny >= 2 (got %d x %d)", mesh->nx, mesh->ny);
    if (!(mesh->xmax > mesh->xmin) || !(mesh->ymax > mesh->ymin)) return fail(UAPIC_EINVAL, "mesh extent must be positive");
    return UAPIC_OK;
}

MeshDev make_mesh(const uapic_mesh_t *mesh) {
    MeshDev m;
    m.xmin = mesh->xmin; m.ymin = mesh->ymin;
    m.dimx = mesh->xmax - mesh->xmin; m.dimy = mesh->ymax - mesh->ymin;
    m.dx = (mesh->xmax - mesh->xmin) / (double)mesh->nx;      // meshfields.F90:71
    m.dy = (mesh->ymax - mesh->ymin) / (double)mesh->ny;      // meshfields.F90:72
    m.nx = mesh->nx; m.ny = mesh->ny; m.ld = mesh->nx + 1;
    return m;
}

int check_ntau(int ntau) {
    if (!ntau_supported(ntau) && !generic_ntau_supported(ntau))
        return fail(UAPIC_EINVAL, "ntau must be even and in [2,%d] (got %d)", kGenericMaxNtau, ntau);
    return UAPIC_OK;
}

// RAII device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        bytes = n;
        if (n == 0) { p = nullptr; return cudaSuccess; }
        return cudaMalloc(&p, n);
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    void swap(DevBuf &o) { std::swap(p, o.p); std::swap(bytes, o.bytes); }
};

// scratch context for the synchronous stage API (device 0 of the current context, default stream)
struct StageCtx {
    LaunchCtx lc;
    int init() {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) { cudaGetLastError(); dev = 0; }
        DeviceInfo di{};
        int rc = device_info(dev, &di);
        if (rc) return rc;
        lc.stream = 0; lc.sm_count = di.sm_count; lc.launches = nullptr;
        return UAPIC_OK;
    }
};

#define STAGE_BEGIN()            \
    StageCtx sc;                 \
    {                            \
        int _rc = sc.init();     \
        if (_rc) return _rc;     \
    }

int up(DevBuf &d, const void *h, size_t bytes) {
    CU(d.alloc(bytes));
    if (bytes && h) CU(cudaMemcpy(d.p, h, bytes, cudaMemcpyHostToDevice));
    return UAPIC_OK;
}
int down(void *h, const DevBuf &d, size_t bytes) {
    if (bytes && h) CU(cudaMemcpy(h, d.p, bytes, cudaMemcpyDeviceToHost));
    return UAPIC_OK;
}
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

double fixed_point_scale(double total_mass) {
    // partial sums never exceed the total deposited mass; keep one guard bit below 2^62
    int e = 0;
    std::frexp(total_mass > 0 ? total_mass : 1.0, &e);   // total_mass < 2^e
    int S = 61 - e;
    if (S > 60) S = 60;
    if (S < 8) S = 8;
    return std::ldexp(1.0, S);
}

struct RawRho {
    DevBuf buf;
    RhoAcc acc{};
    int init(const MeshDev &m, int deposit_mode, double total_mass, cudaStream_t stream) {
        const size_t n = (size_t)m.ld * (m.ny + 1);
        CU(buf.alloc(n * 8));
        CU(cudaMemsetAsync(buf.p, 0, n * 8, stream));
        if (deposit_mode == UAPIC_DEPOSIT_FIXED_POINT) {
            acc.f64 = nullptr; acc.i64 = buf.as<unsigned long long>(); acc.scale = fixed_point_scale(total_mass);
        } else if (deposit_mode == UAPIC_DEPOSIT_FP64_ATOMIC) {
            acc.f64 = buf.as<double>(); acc.i64 = nullptr; acc.scale = 1.0;
        } else {
            return fail(UAPIC_EINVAL, "unknown deposit_mode %d", deposit_mode);
        }
        return UAPIC_OK;
    }
};

}  // namespace

// ================================================================================================
extern "C" {

const char *uapic_last_error(void) { return g_err.c_str(); }
int uapic_version(void) { return 100; }
int uapic_compiled_arch(void) { return 100; }

int uapic_fixed_point_scale(double total_mass, double *scale) {
    if (!scale) return fail(UAPIC_EINVAL, "scale is null");
    *scale = fixed_point_scale(total_mass);
    return UAPIC_OK;
}

int uapic_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count) *count = n;
    if (n <= 0) return fail(UAPIC_ENODEVICE, "no CUDA device available; libuapic_b200 has no CPU fallback");
    return UAPIC_OK;
}

int uapic_probe_fp64_peak(int device, int launches, double *dfma_per_s, double *ms_per_launch) {
    if (launches < 1 || !dfma_per_s) return fail(UAPIC_EINVAL, "uapic_probe_fp64_peak: bad argument");
    DeviceInfo di{};
    TRY(device_info(device, &di));
    CU(cudaSetDevice(device));
    LaunchCtx lc{0, di.sm_count, nullptr};
    CU(probe_fp64_peak(lc, launches, dfma_per_s, ms_per_launch));
    return UAPIC_OK;
}

// ---- stage API ---------------------------------------------------------------------------------------------

static int compute_rho_stage(const uapic_mesh_t *mesh, int64_t nbpart, double *x, double w, double *rho, int wrap,
                             int deposit_mode, double *rho_total, int scheme) {
    TRY(check_mesh(mesh));
    if (!x || !rho || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_compute_rho: null pointer or negative nbpart");
    STAGE_BEGIN();
    const MeshDev m = make_mesh(mesh);
    const size_t nrho = (size_t)m.ld * (m.ny + 1);
    DevBuf dx, drho, dtot;
    RawRho raw;
    TRY(up(dx, x, sizeof(double) * 2 * (size_t)nbpart));
    TRY(raw.init(m, deposit_mode, m.dimx * m.dimy > w * (double)nbpart ? m.dimx * m.dimy : w * (double)nbpart, 0));
    CU(drho.alloc(nrho * 8));
    CU(dtot.alloc(8));
    CU(launch_deposit(sc.lc, m, nbpart, dx.as<double>(), w, raw.acc, wrap, scheme));
    CU(launch_rho_epilogue(sc.lc, m, raw.acc, drho.as<double>(), dtot.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(rho, drho, nrho * 8));
    if (rho_total) TRY(down(rho_total, dtot, 8));
    if (wrap == UAPIC_WRAP_JULIA) TRY(down(x, dx, sizeof(double) * 2 * (size_t)nbpart));
    return UAPIC_OK;
}

int uapic_compute_rho_m6(const uapic_mesh_t *mesh, int64_t nbpart, double *x, double w, double *rho, int wrap,
                         int deposit_mode, double *rho_total) {
    return compute_rho_stage(mesh, nbpart, x, w, rho, wrap, deposit_mode, rho_total, UAPIC_SCHEME_M6);
}

int uapic_compute_rho_cic(const uapic_mesh_t *mesh, int64_t nbpart, double *x, double w, double *rho, int wrap,
                          int deposit_mode, double *rho_total) {
    return compute_rho_stage(mesh, nbpart, x, w, rho, wrap, deposit_mode, rho_total, UAPIC_SCHEME_CIC);
}

static int interpol_eb_stage(const uapic_mesh_t *mesh, const double *e, int64_t nbpart, double *x, double *ep, int wrap, int scheme) {
    TRY(check_mesh(mesh));
    if (!e || !x || !ep || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_interpol_eb: null pointer or negative nbpart");
    STAGE_BEGIN();
    const MeshDev m = make_mesh(mesh);
    const size_t nrho = (size_t)m.ld * (m.ny + 1);
    DevBuf de, dx, dep;
    TRY(up(de, e, nrho * 16));
    TRY(up(dx, x, sizeof(double) * 2 * (size_t)nbpart));
    CU(dep.alloc(sizeof(double) * 2 * (size_t)nbpart));
    CU(launch_gather(sc.lc, m, de.as<double>(), nbpart, dx.as<double>(), dep.as<double>(), wrap, scheme));
    CU(cudaDeviceSynchronize());
    TRY(down(ep, dep, sizeof(double) * 2 * (size_t)nbpart));
    if (wrap == UAPIC_WRAP_JULIA) TRY(down(x, dx, sizeof(double) * 2 * (size_t)nbpart));
    return UAPIC_OK;
}

int uapic_interpol_eb_m6(const uapic_mesh_t *mesh, const double *e, int64_t nbpart, double *x, double *ep, int wrap) {
    return interpol_eb_stage(mesh, e, nbpart, x, ep, wrap, UAPIC_SCHEME_M6);
}

int uapic_interpol_eb_cic(const uapic_mesh_t *mesh, const double *e, int64_t nbpart, double *x, double *ep, int wrap) {
    return interpol_eb_stage(mesh, e, nbpart, x, ep, wrap, UAPIC_SCHEME_CIC);
}

int uapic_poisson(const uapic_mesh_t *mesh, const double *rho, double *e, double *energy) {
    TRY(check_mesh(mesh));
    if (!rho || !e) return fail(UAPIC_EINVAL, "uapic_poisson: null pointer");
    if (!poisson_size_supported(mesh->nx) || !poisson_size_supported(mesh->ny))
        return fail(UAPIC_EUNSUPPORTED, "Poisson mesh %d x %d: sizes must be powers of two <= 1024 or any n <= 512", mesh->nx, mesh->ny);
    STAGE_BEGIN();
    const MeshDev m = make_mesh(mesh);
    const size_t nrho = (size_t)m.ld * (m.ny + 1);
    const size_t nk = (size_t)(m.nx / 2 + 1) * m.ny;
    DevBuf drho, de, drk, dek, dnrj;
    TRY(up(drho, rho, nrho * 8));
    CU(de.alloc(nrho * 16));
    CU(drk.alloc(nk * 16));
    CU(dek.alloc(nk * 32));
    CU(dnrj.alloc(8));
    PoissonWork pw{drk.as<double2>(), dek.as<double2>()};
    CU(launch_poisson(sc.lc, m, pw, drho.as<double>(), de.as<double>(), dnrj.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(e, de, nrho * 16));
    if (energy) TRY(down(energy, dnrj, 8));
    return UAPIC_OK;
}

int uapic_preparation(int ntau, double eps, double dt, int64_t nbpart, const double *x, const double *v, const double *e,
                      double *b, double *t, double *pl, double *ql, double *xt, double *yt) {
    TRY(check_ntau(ntau));
    if (!x || !v || !e || !b || !t || !pl || !ql || !xt || !yt || nbpart < 0)
        return fail(UAPIC_EINVAL, "uapic_preparation: null pointer or negative nbpart");
    STAGE_BEGIN();
    const size_t np = (size_t)nbpart, big = 16 * (size_t)ntau * 2 * np;
    DevBuf dx, dv, de, db, dtt, dpl, dql, dxt, dyt;
    TRY(up(dx, x, 16 * np)); TRY(up(dv, v, 16 * np)); TRY(up(de, e, 16 * np));
    CU(db.alloc(8 * np)); CU(dtt.alloc(8 * np));
    CU(dpl.alloc(16 * (size_t)ntau * np)); CU(dql.alloc(16 * (size_t)ntau * np));
    CU(dxt.alloc(big)); CU(dyt.alloc(big));
    CU(launch_preparation(sc.lc, ntau, eps, dt, nbpart, dx.as<double>(), dv.as<double>(), de.as<double>(), db.as<double>(),
                          dtt.as<double>(), dpl.as<double>(), dql.as<double>(), dxt.as<double>(), dyt.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(b, db, 8 * np)); TRY(down(t, dtt, 8 * np));
    TRY(down(pl, dpl, 16 * (size_t)ntau * np)); TRY(down(ql, dql, 16 * (size_t)ntau * np));
    TRY(down(xt, dxt, big)); TRY(down(yt, dyt, big));
    return UAPIC_OK;
}

int uapic_interpol_eb_m6_tau(const uapic_mesh_t *mesh, const double *e, int ntau, int64_t nbpart, const double *xt,
                             double *et, int wrap) {
    TRY(check_mesh(mesh));
    if (ntau < 1 || !e || !xt || !et || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_interpol_eb_m6_tau: bad argument");
    STAGE_BEGIN();
    const MeshDev m = make_mesh(mesh);
    const size_t nrho = (size_t)m.ld * (m.ny + 1), ns = (size_t)ntau * 2 * (size_t)nbpart;
    DevBuf de, dxt, det;
    TRY(up(de, e, nrho * 16));
    TRY(up(dxt, xt, ns * 16));
    CU(det.alloc(ns * 8));
    CU(launch_gather_tau(sc.lc, m, de.as<double>(), ntau, nbpart, dxt.as<double>(), det.as<double>(), wrap));
    CU(cudaDeviceSynchronize());
    TRY(down(et, det, ns * 8));
    return UAPIC_OK;
}

int uapic_compute_f(int ntau, double eps, int64_t nbpart, const double *b, const double *xt, const double *yt,
                    const double *et, double *fx, double *fy, int normalise) {
    TRY(check_ntau(ntau));
    if (!b || !xt || !yt || !et || !fx || !fy || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_compute_f: bad argument");
    STAGE_BEGIN();
    const size_t np = (size_t)nbpart, ns = (size_t)ntau * 2 * np;
    DevBuf db, dxt, dyt, det, dfx, dfy;
    TRY(up(db, b, 8 * np)); TRY(up(dxt, xt, 16 * ns)); TRY(up(dyt, yt, 16 * ns)); TRY(up(det, et, 8 * ns));
    CU(dfx.alloc(16 * ns)); CU(dfy.alloc(16 * ns));
    CU(launch_compute_f(sc.lc, ntau, eps, nbpart, db.as<double>(), dxt.as<double>(), dyt.as<double>(), det.as<double>(),
                        dfx.as<double>(), dfy.as<double>(), normalise));
    CU(cudaDeviceSynchronize());
    TRY(down(fx, dfx, 16 * ns)); TRY(down(fy, dfy, 16 * ns));
    return UAPIC_OK;
}

int uapic_fft_tau(int ntau, int64_t nvec, const double *in, double *out, int sign, int normalise) {
    TRY(check_ntau(ntau));
    if (!in || !out || nvec < 0 || (sign != 1 && sign != -1)) return fail(UAPIC_EINVAL, "uapic_fft_tau: bad argument");
    STAGE_BEGIN();
    const size_t ns = (size_t)ntau * (size_t)nvec;
    DevBuf din, dout;
    TRY(up(din, in, 16 * ns));
    CU(dout.alloc(16 * ns));
    CU(launch_fft_tau(sc.lc, ntau, nvec, din.as<double>(), dout.as<double>(), sign, normalise));
    CU(cudaDeviceSynchronize());
    TRY(down(out, dout, 16 * ns));
    return UAPIC_OK;
}

static int step_pointwise(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql,
                          const double *xf, const double *fx, const double *gx, double *xt) {
    if (ntau < 2 || (ntau & 1)) return fail(UAPIC_EINVAL, "ntau must be even (got %d)", ntau);
    if (!t || !pl || !xf || !fx || !xt || nbpart < 0) return fail(UAPIC_EINVAL, "ua_step: bad argument");
    STAGE_BEGIN();
    const size_t np = (size_t)nbpart, ns = (size_t)ntau * 2 * np;
    DevBuf dt_, dpl, dql, dxf, dfx, dgx, dout;
    TRY(up(dt_, t, 8 * np)); TRY(up(dpl, pl, 16 * (size_t)ntau * np));
    if (gx) { TRY(up(dql, ql, 16 * (size_t)ntau * np)); TRY(up(dgx, gx, 16 * ns)); }
    TRY(up(dxf, xf, 16 * ns)); TRY(up(dfx, fx, 16 * ns));
    CU(dout.alloc(16 * ns));
    CU(launch_step_pointwise(sc.lc, ntau, eps, nbpart, dt_.as<double>(), dpl.as<double>(), gx ? dql.as<double>() : nullptr,
                             dxf.as<double>(), dfx.as<double>(), gx ? dgx.as<double>() : nullptr, dout.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(xt, dout, 16 * ns));
    return UAPIC_OK;
}

int uapic_ua_step_predict(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *xf,
                          const double *fx, double *xt) {
    return step_pointwise(ntau, eps, nbpart, t, pl, nullptr, xf, fx, nullptr, xt);
}

int uapic_ua_step_correct(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql,
                          const double *xf, const double *fx, const double *gx, double *xt) {
    if (!ql || !gx) return fail(UAPIC_EINVAL, "uapic_ua_step_correct: null pointer");
    return step_pointwise(ntau, eps, nbpart, t, pl, ql, xf, fx, gx, xt);
}

static int step_fortran(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql,
                        double *xt, double *xf, const double *fx, const double *gx, int corrector) {
    TRY(check_ntau(ntau));
    if (!t || !pl || !xt || !xf || !fx || nbpart < 0) return fail(UAPIC_EINVAL, "ua_step1/2: bad argument");
    STAGE_BEGIN();
    const size_t np = (size_t)nbpart, ns = (size_t)ntau * 2 * np;
    DevBuf dt_, dpl, dql, dxt, dxf, dfx, dgx;
    TRY(up(dt_, t, 8 * np)); TRY(up(dpl, pl, 16 * (size_t)ntau * np));
    TRY(up(dfx, fx, 16 * ns));
    if (corrector) {
        TRY(up(dql, ql, 16 * (size_t)ntau * np)); TRY(up(dgx, gx, 16 * ns)); TRY(up(dxf, xf, 16 * ns));
        CU(dxt.alloc(16 * ns));
    } else {
        TRY(up(dxt, xt, 16 * ns));
        CU(dxf.alloc(16 * ns));
    }
    CU(launch_step_fortran(sc.lc, ntau, eps, nbpart, dt_.as<double>(), dpl.as<double>(), corrector ? dql.as<double>() : nullptr,
                           dxt.as<double>(), dxf.as<double>(), dfx.as<double>(), corrector ? dgx.as<double>() : nullptr, corrector));
    CU(cudaDeviceSynchronize());
    TRY(down(xt, dxt, 16 * ns));
    if (!corrector) TRY(down(xf, dxf, 16 * ns));
    return UAPIC_OK;
}

int uapic_ua_step1(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, double *xt, double *xf,
                   const double *fx) {
    return step_fortran(ntau, eps, nbpart, t, pl, nullptr, xt, xf, fx, nullptr, 0);
}

int uapic_ua_step2(int ntau, double eps, int64_t nbpart, const double *t, const double *pl, const double *ql, double *xt,
                   const double *xf, const double *fx, const double *gx) {
    if (!ql || !gx) return fail(UAPIC_EINVAL, "uapic_ua_step2: null pointer");
    return step_fortran(ntau, eps, nbpart, t, pl, ql, xt, const_cast<double *>(xf), fx, gx, 1);
}

int uapic_compute_rho_m6_tau(const uapic_mesh_t *mesh, int ntau, double eps, int64_t nbpart, const double *xt,
                             const double *t, double w, double *rho, double *x, int wrap, int deposit_mode,
                             double *rho_total) {
    TRY(check_mesh(mesh));
    TRY(check_ntau(ntau));
    if (!xt || !t || !rho || !x || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_compute_rho_m6_tau: bad argument");
    STAGE_BEGIN();
    const MeshDev m = make_mesh(mesh);
    const size_t np = (size_t)nbpart, ns = (size_t)ntau * 2 * np, nrho = (size_t)m.ld * (m.ny + 1);
    DevBuf dxt, dt_, dx, drho, dtot;
    RawRho raw;
    TRY(up(dxt, xt, 16 * ns)); TRY(up(dt_, t, 8 * np));
    CU(dx.alloc(16 * np)); CU(drho.alloc(8 * nrho)); CU(dtot.alloc(8));
    TRY(raw.init(m, deposit_mode, m.dimx * m.dimy > w * (double)nbpart ? m.dimx * m.dimy : w * (double)nbpart, 0));
    CU(launch_deposit_tau(sc.lc, m, ntau, eps, nbpart, dxt.as<double>(), dt_.as<double>(), w, raw.acc, dx.as<double>(), wrap));
    CU(launch_rho_epilogue(sc.lc, m, raw.acc, drho.as<double>(), dtot.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(rho, drho, 8 * nrho)); TRY(down(x, dx, 16 * np));
    if (rho_total) TRY(down(rho_total, dtot, 8));
    return UAPIC_OK;
}

int uapic_compute_v(int ntau, double eps, int64_t nbpart, const double *t, const double *yt, int yt_is_fourier, double *v) {
    TRY(check_ntau(ntau));
    if (!t || !yt || !v || nbpart < 0) return fail(UAPIC_EINVAL, "uapic_compute_v: bad argument");
    STAGE_BEGIN();
    const size_t np = (size_t)nbpart, ns = (size_t)ntau * 2 * np;
    DevBuf dt_, dyt, dv;
    TRY(up(dt_, t, 8 * np)); TRY(up(dyt, yt, 16 * ns));
    CU(dv.alloc(16 * np));
    CU(launch_compute_v(sc.lc, ntau, eps, nbpart, dt_.as<double>(), dyt.as<double>(), yt_is_fourier, dv.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(v, dv, 16 * np));
    return UAPIC_OK;
}

static int efd_check(const uapic_efd_config_t *cfg, int64_t nbpart, const double *x, const double *v, double *xo, double *vo, int *nstep) {
    if (!cfg || nbpart < 0 || (nbpart > 0 && (!x || !v || !xo || !vo))) return fail(UAPIC_EINVAL, "uapic_efd_run: bad argument");
    if (!efd_ntau_supported(cfg->ntau)) return fail(UAPIC_EINVAL, "ntau must be even, 2..%d (got %d)", kGenericMaxNtau, cfg->ntau);
    if (!(cfg->eps > 0) || !(cfg->dt > 0) || !(cfg->tfinal >= 0) || !(cfg->xmax > cfg->xmin) || !(cfg->ymax > cfg->ymin) || cfg->nstep < 0)
        return fail(UAPIC_EINVAL, "uapic_efd_run: eps, dt must be positive, tfinal and nstep non-negative, the box non-empty");
    *nstep = cfg->nstep > 0 ? cfg->nstep : (int)std::lround(cfg->tfinal / cfg->dt);          // efd.f90:102
    return UAPIC_OK;
}

int uapic_efd_run_device(const uapic_efd_config_t *cfg, int64_t nbpart, const double *x, const double *v, double *x_out,
                         double *v_out, void *stream) {
    int nstep = 0;
    TRY(efd_check(cfg, nbpart, x, v, x_out, v_out, &nstep));
    STAGE_BEGIN();
    sc.lc.stream = reinterpret_cast<cudaStream_t>(stream);
    const double box[4] = {cfg->xmin, cfg->xmax, cfg->ymin, cfg->ymax};
    CU(launch_efd(sc.lc, cfg->ntau, cfg->eps, cfg->dt, cfg->tfinal, nstep, box, nbpart, x, v, x_out, v_out));
    return UAPIC_OK;
}

int uapic_efd_run(const uapic_efd_config_t *cfg, int64_t nbpart, const double *x, const double *v, double *x_out, double *v_out) {
    int nstep = 0;
    TRY(efd_check(cfg, nbpart, x, v, x_out, v_out, &nstep));
    STAGE_BEGIN();
    const size_t bytes = 16 * (size_t)nbpart;
    DevBuf dx, dv;
    TRY(up(dx, x, bytes)); TRY(up(dv, v, bytes));
    const double box[4] = {cfg->xmin, cfg->xmax, cfg->ymin, cfg->ymax};
    CU(launch_efd(sc.lc, cfg->ntau, cfg->eps, cfg->dt, cfg->tfinal, nstep, box, nbpart, dx.as<double>(), dv.as<double>(),
                  dx.as<double>(), dv.as<double>()));
    CU(cudaDeviceSynchronize());
    TRY(down(x_out, dx, bytes)); TRY(down(v_out, dv, bytes));
    return UAPIC_OK;
}

}  // extern "C"

// ================================================================================================
// session
// ================================================================================================

#ifndef UAPIC_HOST_CHUNKS
#define UAPIC_HOST_CHUNKS 64     // most particle chunks of uapic_session_step_host (copies of a chunk overlap the kernels of its neighbours)
#endif
#ifndef UAPIC_FUSE_DEFAULT
#define UAPIC_FUSE_DEFAULT 0            // phase B of step n inside phase A of step n+1 (uapic_session_set_fusion); $UAPIC_FUSE_BA overrides
#endif
#ifndef UAPIC_FUSE_SORT_INTERVAL
#define UAPIC_FUSE_SORT_INTERVAL 4      // fused mode: the particles can only be reordered where the store is dead, every this many steps
#endif
#ifndef UAPIC_RAW_COPIES
#define UAPIC_RAW_COPIES 8      // CTA-private copies of the two raw deposit meshes of the one-pass kernels (measured: -2 % on phase A)
#endif
// ---- NCCL, bound at run time (dlopen): the library links nothing but the CUDA runtime -------------------------------
// Prototypes restated from nccl.h (NCCL 2.x ABI; ncclUniqueId is 128 opaque bytes passed by value; ncclSum = 0, ncclInt64 = 4,
// ncclFloat64 = 8).  Lookup order: $UAPIC_NCCL_LIB, then "libnccl.so.2" (which resolves to an NCCL the process has already
// loaded -- e.g. the one bundled with torch -- before the system copy).
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
    void *h = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *env = getenv("UAPIC_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        if (!n || !*n) continue;
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
        api.err = dlerror();
    }
    if (!api.h) return &api;
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(api.h, n); if (!p) { ok = false; api.err = std::string("missing symbol ") + n; } return p; };
    api.GetVersion = reinterpret_cast<int (*)(int *)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<int (*)(NcclId *)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void **, int, NcclId, int)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(void *)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<const char *(*)(int)>(sym("ncclGetErrorString"));
    if (!ok) { dlclose(api.h); api.h = nullptr; }
    return &api;
}
int nccl_need(NcclApi **out) {
    NcclApi *a = nccl_api();
    if (!a->h) return fail(UAPIC_EUNSUPPORTED, "NCCL is not loadable (%s); set UAPIC_NCCL_LIB or LD_LIBRARY_PATH to a libnccl.so.2", a->err.c_str());
    *out = a;
    return UAPIC_OK;
}
}  // namespace

// the same run-time NCCL binding for the other translation units (uapic_mrc3d.cu)
int uapic_internal_nccl_comm_init(void **comm, const void *id128, int nranks, int rank) {
    NcclApi *a = nullptr;
    TRY(nccl_need(&a));
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    const int rc = a->CommInitRank(comm, nranks, id, rank);
    if (rc) return fail(UAPIC_ECUDA, "ncclCommInitRank(rank %d of %d) failed: %s", rank, nranks, a->GetErrorString(rc));
    return UAPIC_OK;
}
int uapic_internal_nccl_allreduce(void *comm, void *buf, size_t count, int is_i64, cudaStream_t st) {
    NcclApi *a = nccl_api();
    const int rc = a->AllReduce(buf, buf, count, is_i64 ? 4 : 8, 0, comm, st);
    if (rc) return fail(UAPIC_ECUDA, "ncclAllReduce failed: %s", a->GetErrorString(rc));
    return UAPIC_OK;
}
void uapic_internal_nccl_comm_destroy(void *comm) {
    NcclApi *a = nccl_api();
    if (a->h && comm) a->CommDestroy(comm);
}

struct uapic_session {
    uapic_config_t cfg;
    MeshDev m;
    LaunchCtx lc;
    int64_t launches = 0;
    int64_t np_global = 0;
    DevBuf x, v, ep, store, tb, raw, rho, emesh, ehalo, rk, ek, energy, sumv;
    DevBuf rec, emesh_p, ehalo_p;     // one-pass modes: per-particle record, predictor field
    DevBuf gb, gt, gpl, gql, gxt, gyt, gxf, gyf, gfx, gfy, ggx, ggy, get;   // generic-ntau sessions: the reference's own arrays (uapic_generic.cu)
    bool generic = false;
    DevBuf rho_p, rk2, ek2, solve_scratch;   // second work set + scratch of the batched one-launch field solve (k_field_solve)
    bool split_solve = false;          // $UAPIC_SPLIT_SOLVE=1: the six separate kernels per solve (A/B measurement, stage-API kernels)
    RhoAcc acc{};
    RhoAcc acc_c{};                    // one-pass modes: corrector deposit mesh (second half of raw)
    bool onepass = false;
    // spatial reordering (uapic_sort.cu): alternate buffers, slot -> original index, scratch; allocated at the first sort
    DevBuf x2, v2, ep2, perm, perm2, binid, hist;
    bool permuted = false;             // device arrays are in sorted order, perm is valid
    bool fuse = false;                 // run phase B of step n inside phase A of step n+1 (k_onepass_a<..., FUSEB>; lean layout)
    bool pending_b = false;            // fused mode: phase B of the last step has not run yet (v on the device is one step old)
    bool sort_bufs_ready = false;      // the whole set of alternate buffers exists (all-or-nothing)
    int sort_interval = 0, sort_shift = 3;
    int64_t steps_done = 0;
    // uapic_session_step_host: copy streams and per-chunk events (created at first use)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    int64_t n_energy = 0, cap_energy = 0;
    uapic_allreduce_fn reduce = nullptr;
    void *reduce_ctx = nullptr;
    // peer-memory exchange (uapic_session_init_peers): every rank's folded deposits sit in an IPC-exported buffer that the other
    // ranks map over NVLink; the sum over the ranks happens inside k_field_solve -- no collective call at all
    DevBuf xchg, peer_err;             // [256 B of flags | parity 0: 2 meshes | parity 1: 2 meshes]
    void *peer_base[kMaxPeers] = {};   // rank r's xchg as mapped here (own entry = xchg.p)
    int npeers = 0, peer_rank = 0;
    unsigned long long xchg_seq = 0;
    void *nccl_comm = nullptr;         // ncclComm_t: the library sums the raw rho meshes itself (uapic_session_init_nccl)
    bool nccl_owned = false;
    int nccl_rank = 0, nccl_nranks = 1;
    bool have_particles = false, fields_ready = false;
    int64_t bytes = 0;
    // optional per-kernel timing
    bool timing = false;
    std::vector<cudaEvent_t> ev;     // 4 per step: A0 A1 B0 B1
    double ms_a = 0, ms_b = 0, ms_f = 0;   // phase A, phase B, and what lies between them: fold + all-reduce + field solve
    int64_t timed_steps = 0;
    double ms_field_last = 0;
    ~uapic_session() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        for (cudaEvent_t e : chunk_ev) cudaEventDestroy(e);
        if (up_stream) cudaStreamDestroy(up_stream);
        if (down_stream) cudaStreamDestroy(down_stream);
        if (nccl_comm && nccl_owned) { NcclApi *a = nccl_api(); if (a->h) a->CommDestroy(nccl_comm); }
        for (int r = 0; r < npeers; ++r) if (r != peer_rank && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
    }
};

namespace {

int session_alloc(uapic_session *s, DevBuf &b, size_t n) {
    cudaError_t e = b.alloc(n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(UAPIC_ENOMEM, "device allocation of %.3f GB failed (%s); session already holds %.3f GB", (double)n / 1e9,
                    cudaGetErrorString(e), (double)s->bytes / 1e9);
    }
    s->bytes += (int64_t)n;
    return UAPIC_OK;
}

int session_bind(uapic_session *s) {
    CU(cudaSetDevice(s->cfg.device));
    return UAPIC_OK;
}

int session_clear_raw(uapic_session *s) {
    CU(cudaMemsetAsync(s->raw.p, 0, s->raw.bytes, s->lc.stream));
    return UAPIC_OK;
}

// summed raw deposits -> neutralised rho -> E (+ halo copy), energy appended      compute_rho_m6.F90:191-200 + poisson_2d.f90:85-111
// (the six separate kernels; kept for $UAPIC_SPLIT_SOLVE=1 and as the cross-check of k_field_solve)
int session_solve_split(uapic_session *s, const RhoAcc &acc, DevBuf &emesh, DevBuf &ehalo) {
    CU(launch_rho_epilogue(s->lc, s->m, acc, s->rho.as<double>(), nullptr));
    if (s->n_energy >= s->cap_energy) return fail(UAPIC_ESTATE, "energy history full (%lld entries): session_reserve_energy was not called", (long long)s->cap_energy);
    PoissonWork pw{s->rk.as<double2>(), s->ek.as<double2>()};
    CU(launch_poisson(s->lc, s->m, pw, s->rho.as<double>(), emesh.as<double>(), s->energy.as<double>() + s->n_energy));
    if (s->onepass) CU(launch_extend_emesh_tiled(s->lc, s->m, emesh.as<double>(), ehalo.as<double2>()));
    else CU(launch_extend_emesh(s->lc, s->m, emesh.as<double>(), ehalo.as<double2>()));
    s->n_energy++;
    return UAPIC_OK;
}

// the field barrier of a step: sum the raw deposit mesh(es) over the ranks (the only exchange of the scheme), then ONE
// cooperative launch takes all `nmesh` meshes from raw deposits to E, halo copy and energy (k_field_solve).
// nmesh = 2 (one-pass kernels): predictor mesh -> emesh_p / ehalo_p, corrector mesh -> emesh / ehalo, energies in this order.
int session_field_barrier(uapic_session *s, int nmesh) {
    if (s->n_energy + nmesh > s->cap_energy) return fail(UAPIC_ESTATE, "energy history full (%lld entries): session_reserve_energy was not called", (long long)s->cap_energy);
    const size_t nrho = (size_t)s->m.ld * (s->m.ny + 1);
    const bool copies = s->onepass && nmesh == 2;
    const bool exchange = s->nccl_comm || s->reduce;
    int fold_in_kernel = 1;
    if (s->npeers > 0 && s->split_solve)
        return fail(UAPIC_EUNSUPPORTED, "UAPIC_SPLIT_SOLVE=1 (the six-kernel field solve) cannot sum over peer memory: use NCCL or the callback with it");
    if (s->npeers > 0) {
        // fold my copies into my exchange buffer, publish, and let the solve kernel add the ranks' buffers over NVLink
        const unsigned long long seq = ++s->xchg_seq;
        const size_t half = 2 * nrho * 8;                              // bytes of one parity (room for two meshes)
        const size_t off = 256 + (size_t)(seq & 1) * half;
        CU(launch_fold_publish(s->lc, s->acc, (int64_t)nrho * nmesh, copies ? UAPIC_RAW_COPIES : 1, s->xchg.as<char>() + off,
                               s->xchg.as<unsigned long long>(), seq));
        SolveBatch B{};
        B.nb = nmesh;
        B.partial = s->solve_scratch.as<double>();
        B.halo_tiled = s->onepass ? 1 : 0;
        B.fold_copies = 1;
        B.fold_stride = 2 * nrho;
        B.npeers = s->npeers;
        for (int r = 0; r < s->npeers; ++r) {
            B.peer_data[r] = reinterpret_cast<const unsigned long long *>(static_cast<const char *>(s->peer_base[r]) + off);
            B.peer_flag[r] = reinterpret_cast<const unsigned long long *>(s->peer_base[r]);
        }
        B.peer_seq = seq;
        B.peer_mesh_stride = nrho;
        B.peer_error = s->peer_err.as<int>();
        if (nmesh == 2) {
            B.acc[0] = s->acc;   B.rho[0] = s->rho_p.as<double>(); B.emesh[0] = s->emesh_p.as<double2>(); B.ehalo[0] = s->ehalo_p.as<double2>();
            B.rk[0] = s->rk2.as<double2>(); B.ek[0] = s->ek2.as<double2>(); B.energy[0] = s->energy.as<double>() + s->n_energy;
            B.acc[1] = s->acc_c; B.rho[1] = s->rho.as<double>();   B.emesh[1] = s->emesh.as<double2>();   B.ehalo[1] = s->ehalo.as<double2>();
            B.rk[1] = s->rk.as<double2>();  B.ek[1] = s->ek.as<double2>();  B.energy[1] = s->energy.as<double>() + s->n_energy + 1;
        } else {
            B.acc[0] = s->acc;   B.rho[0] = s->rho.as<double>();   B.emesh[0] = s->emesh.as<double2>();   B.ehalo[0] = s->ehalo.as<double2>();
            B.rk[0] = s->rk.as<double2>();  B.ek[0] = s->ek.as<double2>();  B.energy[0] = s->energy.as<double>() + s->n_energy;
        }
        CU(launch_field_solve(s->lc, s->m, B));
        s->n_energy += nmesh;
        return UAPIC_OK;
    }
    if (copies) {
        if (exchange || s->split_solve) CU(launch_fold_raw(s->lc, s->acc, 2 * (int64_t)nrho, UAPIC_RAW_COPIES));
        else fold_in_kernel = UAPIC_RAW_COPIES;        // single GPU: the solve kernel folds the copies while it scales them
    }
    if (s->nccl_comm) {
        // in-library collective on the session's stream: no host callback, capturable in a CUDA graph
        NcclApi *a = nccl_api();
        const int rc = a->AllReduce(s->raw.p, s->raw.p, nrho * nmesh, s->acc.i64 ? 4 /* ncclInt64 */ : 8 /* ncclFloat64 */, 0 /* ncclSum */,
                                    s->nccl_comm, s->lc.stream);
        if (rc) return fail(UAPIC_ECUDA, "ncclAllReduce failed: %s", a->GetErrorString(rc));
    } else if (s->reduce) {
        int rc = s->reduce(s->reduce_ctx, s->raw.p, (int64_t)nrho * nmesh, s->acc.i64 ? 1 : 0, (void *)s->lc.stream);
        if (rc) return fail(UAPIC_ECUDA, "allreduce callback failed with code %d", rc);
    }
    if (s->split_solve) {
        if (nmesh == 2) {
            TRY(session_solve_split(s, s->acc, s->emesh_p, s->ehalo_p));
            return session_solve_split(s, s->acc_c, s->emesh, s->ehalo);
        }
        return session_solve_split(s, s->acc, s->emesh, s->ehalo);
    }
    SolveBatch B{};
    B.nb = nmesh;
    B.partial = s->solve_scratch.as<double>();
    B.halo_tiled = s->onepass ? 1 : 0;
    B.fold_copies = fold_in_kernel;
    B.fold_stride = 2 * nrho;
    if (nmesh == 2) {
        B.acc[0] = s->acc;   B.rho[0] = s->rho_p.as<double>(); B.emesh[0] = s->emesh_p.as<double2>(); B.ehalo[0] = s->ehalo_p.as<double2>();
        B.rk[0] = s->rk2.as<double2>(); B.ek[0] = s->ek2.as<double2>(); B.energy[0] = s->energy.as<double>() + s->n_energy;
        B.acc[1] = s->acc_c; B.rho[1] = s->rho.as<double>();   B.emesh[1] = s->emesh.as<double2>();   B.ehalo[1] = s->ehalo.as<double2>();
        B.rk[1] = s->rk.as<double2>();  B.ek[1] = s->ek.as<double2>();  B.energy[1] = s->energy.as<double>() + s->n_energy + 1;
    } else {
        B.acc[0] = s->acc;   B.rho[0] = s->rho.as<double>();   B.emesh[0] = s->emesh.as<double2>();   B.ehalo[0] = s->ehalo.as<double2>();
        B.rk[0] = s->rk.as<double2>();  B.ek[0] = s->ek.as<double2>();  B.energy[0] = s->energy.as<double>() + s->n_energy;
    }
    CU(launch_field_solve(s->lc, s->m, B));
    s->n_energy += nmesh;
    return UAPIC_OK;
}

int session_field_solve(uapic_session *s) { return session_field_barrier(s, 1); }

OnepassParams session_onepass_params(uapic_session *s) {
    OnepassParams p;
    p.m = s->m; p.eps = s->cfg.eps; p.dt = s->cfg.dt; p.weight = s->cfg.weight; p.np = s->cfg.nbpart;
    p.wrap = s->cfg.wrap; p.ntau = s->cfg.ntau; p.full = s->cfg.storage_mode == UAPIC_STORE_ONEPASS; p.scheme = s->cfg.scheme;
    p.x = s->x.as<double2>(); p.v = s->v.as<double2>(); p.ep = s->ep.as<double2>();
    p.ehalo = s->ehalo.as<double2>();
    p.store = s->store.as<char>(); p.rec = s->rec.as<double>();
    p.rho_p = s->acc; p.rho_c = s->acc_c; p.rho_copies = UAPIC_RAW_COPIES;
    p.out_perm = nullptr; p.x_out = nullptr; p.v_out = nullptr; p.ehalo_b = nullptr; p.fuse_b = 0;
    return p;
}

// fused mode leaves phase B of the last step pending (it runs inside the next step's first kernel); everything that reads or
// replaces v, or reorders the particles, runs it on its own first
int session_flush_b(uapic_session *s) {
    if (!s->pending_b) return UAPIC_OK;
    OnepassParams op = session_onepass_params(s);
    op.ehalo = s->ehalo_p.as<double2>();
    CU(launch_onepass_b(s->lc, op));
    s->pending_b = false;
    return UAPIC_OK;
}

// reorder x, v, ep by coarse mesh bin; callers never see the order (downloads undo it)
int session_alloc_sort_buffers(uapic_session *s) {
    const size_t np = (size_t)s->cfg.nbpart;
    if (s->sort_bufs_ready || np == 0) return UAPIC_OK;
    int rc = UAPIC_OK;
    if (!rc) rc = session_alloc(s, s->x2, 16 * np);
    if (!rc) rc = session_alloc(s, s->v2, 16 * np);
    if (!rc) rc = session_alloc(s, s->ep2, 16 * np);
    if (!rc) rc = session_alloc(s, s->perm, 4 * np);
    if (!rc) rc = session_alloc(s, s->perm2, 4 * np);
    if (!rc) rc = session_alloc(s, s->binid, 2 * np);
    if (!rc) rc = session_alloc(s, s->hist, 4 * 4096);
    if (rc) {
        // all or nothing: a partial set would let a later step write through null buffers
        for (DevBuf *b : {&s->x2, &s->v2, &s->ep2, &s->perm, &s->perm2, &s->binid, &s->hist}) {
            if (b->p) { s->bytes -= (int64_t)b->bytes; cudaFree(b->p); b->p = nullptr; b->bytes = 0; }
        }
        return rc;
    }
    s->sort_bufs_ready = true;
    return UAPIC_OK;
}

// the reordering is an optimisation: if its buffers do not fit, run without it rather than fail the step
int session_want_sort(uapic_session *s, bool *on) {
    *on = false;
    if (s->sort_interval <= 0 || s->cfg.nbpart == 0) return UAPIC_OK;
    const int rc = session_alloc_sort_buffers(s);
    if (rc == UAPIC_ENOMEM) { s->sort_interval = 0; return UAPIC_OK; }
    if (rc) return rc;
    *on = true;
    return UAPIC_OK;
}

// room for `need` more entries of the energy history (checked BEFORE anything of a step is launched); grows by doubling
int session_reserve_energy(uapic_session *s, int64_t need) {
    if (s->n_energy + need <= s->cap_energy) return UAPIC_OK;
    int64_t cap = s->cap_energy;
    while (cap < s->n_energy + need) cap *= 2;
    DevBuf nb;
    cudaError_t e = nb.alloc(8 * (size_t)cap);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(UAPIC_ENOMEM, "cannot grow the energy history to %lld entries", (long long)cap); }
    CU(cudaMemcpyAsync(nb.p, s->energy.p, 8 * (size_t)s->n_energy, cudaMemcpyDeviceToDevice, s->lc.stream));
    CU(cudaStreamSynchronize(s->lc.stream));
    s->bytes += (int64_t)nb.bytes - (int64_t)s->energy.bytes;
    s->energy.swap(nb);
    s->cap_energy = cap;
    return UAPIC_OK;
}

int session_sort(uapic_session *s) {
    const size_t np = (size_t)s->cfg.nbpart;
    if (np == 0) return UAPIC_OK;
    TRY(session_alloc_sort_buffers(s));
    CU(launch_sort_particles(s->lc, s->m, s->sort_shift, s->cfg.nbpart, s->x.as<double2>(), s->v.as<double2>(), s->ep.as<double2>(),
                             s->permuted ? s->perm.as<uint32_t>() : nullptr, s->x2.as<double2>(), s->v2.as<double2>(),
                             s->ep2.as<double2>(), s->perm2.as<uint32_t>(), s->binid.as<uint16_t>(), s->hist.as<unsigned>()));
    s->x.swap(s->x2); s->v.swap(s->v2); s->ep.swap(s->ep2); s->perm.swap(s->perm2);
    s->permuted = true;
    return UAPIC_OK;
}

// device array in slot order -> host array in the caller's particle order
int session_download_pairs(uapic_session *s, DevBuf &src, DevBuf &tmp, double *host) {
    const size_t n = 16 * (size_t)s->cfg.nbpart;
    if (s->permuted) {
        CU(launch_unpermute(s->lc, s->cfg.nbpart, s->perm.as<uint32_t>(), src.as<double2>(), tmp.as<double2>()));
        CU(cudaMemcpyAsync(host, tmp.p, n, cudaMemcpyDeviceToHost, s->lc.stream));
    } else {
        CU(cudaMemcpyAsync(host, src.p, n, cudaMemcpyDeviceToHost, s->lc.stream));
    }
    return UAPIC_OK;
}

PhaseParams session_params(uapic_session *s) {
    PhaseParams p;
    p.m = s->m; p.eps = s->cfg.eps; p.dt = s->cfg.dt; p.weight = s->cfg.weight; p.np = s->cfg.nbpart;
    p.wrap = s->cfg.wrap; p.ntau = s->cfg.ntau;
    p.x = s->x.as<double2>(); p.v = s->v.as<double2>(); p.ep = s->ep.as<double2>();
    p.emesh = s->emesh.as<double2>(); p.ehalo = s->ehalo.as<double2>(); p.store = s->store.as<double2>(); p.etstore = s->store.as<double>(); p.hybrid = s->cfg.storage_mode == UAPIC_STORE_HYBRID; p.tb = s->tb.as<double2>();
    p.rho = s->acc;
    return p;
}

}  // namespace

extern "C" {

int uapic_session_create(const uapic_config_t *cfg, uapic_session_t **out) {
    if (!cfg || !out) return fail(UAPIC_EINVAL, "uapic_session_create: null pointer");
    *out = nullptr;
    TRY(check_mesh(&cfg->mesh));
    TRY(check_ntau(cfg->ntau));
    if (cfg->nbpart < 0) return fail(UAPIC_EINVAL, "nbpart must be >= 0");
    if (!(cfg->eps > 0) || !(cfg->dt > 0) || !(cfg->weight > 0)) return fail(UAPIC_EINVAL, "eps, dt and weight must be positive");
    if (cfg->scheme != UAPIC_SCHEME_M6 && cfg->scheme != UAPIC_SCHEME_CIC) return fail(UAPIC_EINVAL, "unknown scheme %d", cfg->scheme);
    if (cfg->scheme == UAPIC_SCHEME_CIC && cfg->storage_mode != UAPIC_STORE_ONEPASS_LEAN)
        return fail(UAPIC_EUNSUPPORTED, "UAPIC_SCHEME_CIC is implemented for storage_mode UAPIC_STORE_ONEPASS_LEAN only (ntau = 8, 16, 32)");
    const bool onepass = cfg->storage_mode == UAPIC_STORE_ONEPASS || cfg->storage_mode == UAPIC_STORE_ONEPASS_LEAN;
    if (cfg->storage_mode != UAPIC_STORE_FULL && cfg->storage_mode != UAPIC_STORE_HYBRID && !onepass) return fail(UAPIC_EINVAL, "unknown storage_mode %d", cfg->storage_mode);
    if (onepass && !onepass_ntau_supported(cfg->ntau)) return fail(UAPIC_EUNSUPPORTED, "the one-pass storage modes need ntau = 8, 16 or 32 (got %d)", cfg->ntau);
    const bool generic = !ntau_supported(cfg->ntau);
    if (generic && (cfg->storage_mode != UAPIC_STORE_FULL || cfg->scheme != UAPIC_SCHEME_M6))
        return fail(UAPIC_EUNSUPPORTED, "ntau = %d runs on the general one-warp-per-particle kernels: UAPIC_STORE_FULL and UAPIC_SCHEME_M6 only", cfg->ntau);
    if (cfg->wrap != UAPIC_WRAP_FORTRAN && cfg->wrap != UAPIC_WRAP_JULIA) return fail(UAPIC_EINVAL, "unknown wrap %d", cfg->wrap);
    if (!poisson_size_supported(cfg->mesh.nx) || !poisson_size_supported(cfg->mesh.ny) || cfg->mesh.nx < 4 || cfg->mesh.ny < 4)
        return fail(UAPIC_EUNSUPPORTED, "session mesh %d x %d unsupported (need 4 <= n, powers of two <= 1024 or any n <= 512)", cfg->mesh.nx, cfg->mesh.ny);
    DeviceInfo di{};
    TRY(device_info(cfg->device, &di));
    CU(cudaSetDevice(cfg->device));

    uapic_session *s = new (std::nothrow) uapic_session();
    if (!s) return fail(UAPIC_ENOMEM, "host allocation failed");
    s->cfg = *cfg;
    s->m = make_mesh(&cfg->mesh);
    s->lc.stream = (cudaStream_t)cfg->stream;
    s->lc.sm_count = di.sm_count;
    s->lc.launches = &s->launches;
    const double total_mass = cfg->total_mass > 0 ? cfg->total_mass : s->m.dimx * s->m.dimy;
    s->np_global = (int64_t)llround(total_mass / cfg->weight);

    const size_t np = (size_t)cfg->nbpart, N = (size_t)cfg->ntau;
    const size_t nrho = (size_t)s->m.ld * (s->m.ny + 1), nk = (size_t)(s->m.nx / 2 + 1) * s->m.ny;
    int rc = UAPIC_OK;
    s->cap_energy = 1 << 12;
    if (!rc) rc = session_alloc(s, s->x, 16 * (np ? np : 1));
    if (!rc) rc = session_alloc(s, s->v, 16 * (np ? np : 1));
    if (!rc) rc = session_alloc(s, s->ep, 16 * (np ? np : 1));
    // store-full: 128 B per particle-tau; hybrid: 16 B (E at the tau samples); one-pass: 72 B / 48 B (lean)
    s->onepass = onepass;
    // the one-pass kernels are bound by the L1 data pipe of the M6 gathers: keep the particles ordered by 8 x 8-cell bins
    s->sort_interval = onepass ? 1 : 0;
    s->sort_shift = 3;
    while (sort_bins(s->m, s->sort_shift) > 4096) s->sort_shift++;
    s->generic = generic;
    if (!rc && !onepass && !generic) rc = session_alloc(s, s->tb, 16 * (np ? np : 1));
    const size_t per_tau = onepass ? (cfg->storage_mode == UAPIC_STORE_ONEPASS ? 72 : 48) : (cfg->storage_mode == UAPIC_STORE_HYBRID ? 16 : 128);
    if (!rc && !generic) rc = session_alloc(s, s->store, per_tau * N * (np ? np : 1));
    if (generic) {
        // the reference's own working set (bupdate.F90:79-87): b, t, pl, ql, xt, xf, yt, yf, fx, fy, gx, gy, et
        const size_t n1 = np ? np : 1, big = 32 * N * n1;
        if (!rc) rc = session_alloc(s, s->gb, 8 * n1);
        if (!rc) rc = session_alloc(s, s->gt, 8 * n1);
        if (!rc) rc = session_alloc(s, s->gpl, 16 * N * n1);
        if (!rc) rc = session_alloc(s, s->gql, 16 * N * n1);
        for (DevBuf *b : {&s->gxt, &s->gyt, &s->gxf, &s->gyf, &s->gfx, &s->gfy, &s->ggx, &s->ggy})
            if (!rc) rc = session_alloc(s, *b, big);
        if (!rc) rc = session_alloc(s, s->get, 16 * N * n1);
    }
    if (!rc) rc = session_alloc(s, s->raw, 8 * nrho * (onepass ? 2 * UAPIC_RAW_COPIES : 1));
    if (onepass) {
        if (!rc) rc = session_alloc(s, s->rec, 64 * (np ? np : 1));
        if (!rc) rc = session_alloc(s, s->emesh_p, 16 * nrho);
        if (!rc) rc = session_alloc(s, s->ehalo_p, 16 * ehalo_tiled_nodes(s->m));
    }
    if (!rc) rc = session_alloc(s, s->rho, 8 * nrho);
    if (!rc) rc = session_alloc(s, s->emesh, 16 * nrho);
    if (!rc) rc = session_alloc(s, s->ehalo, 16 * (onepass ? ehalo_tiled_nodes(s->m) : ehalo_nodes(s->m)));
    if (!rc) rc = session_alloc(s, s->rk, 16 * nk);
    if (!rc) rc = session_alloc(s, s->ek, 32 * nk);
    if (!rc) rc = session_alloc(s, s->solve_scratch, field_solve_scratch_bytes());
    if (onepass) {
        if (!rc) rc = session_alloc(s, s->rho_p, 8 * nrho);
        if (!rc) rc = session_alloc(s, s->rk2, 16 * nk);
        if (!rc) rc = session_alloc(s, s->ek2, 32 * nk);
    }
    { const char *e = getenv("UAPIC_SPLIT_SOLVE"); s->split_solve = e && *e == '1'; }
    { const char *e = getenv("UAPIC_FUSE_BA"); s->fuse = (e ? *e == '1' : UAPIC_FUSE_DEFAULT) && cfg->storage_mode == UAPIC_STORE_ONEPASS_LEAN;
      const char *k = getenv("UAPIC_FUSE_SORT"); if (s->fuse) s->sort_interval = k ? atoi(k) : UAPIC_FUSE_SORT_INTERVAL; }
    if (!rc) rc = session_alloc(s, s->energy, 8 * (size_t)s->cap_energy);
    if (!rc) rc = session_alloc(s, s->sumv, 16 + sum_v_scratch_bytes());
    if (rc) { delete s; return rc; }
    if (cfg->deposit_mode == UAPIC_DEPOSIT_FIXED_POINT) {
        s->acc.f64 = nullptr; s->acc.i64 = s->raw.as<unsigned long long>(); s->acc.scale = fixed_point_scale(total_mass);
        s->acc_c = s->acc; s->acc_c.i64 += nrho;
    } else if (cfg->deposit_mode == UAPIC_DEPOSIT_FP64_ATOMIC) {
        s->acc.f64 = s->raw.as<double>(); s->acc.i64 = nullptr; s->acc.scale = 1.0;
        s->acc_c = s->acc; s->acc_c.f64 += nrho;
    } else {
        delete s;
        return fail(UAPIC_EINVAL, "unknown deposit_mode %d", cfg->deposit_mode);
    }
    cudaError_t e = cudaMemsetAsync(s->emesh.p, 0, 16 * nrho, s->lc.stream);
    if (e != cudaSuccess) { delete s; return fail(UAPIC_ECUDA, "memset failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return UAPIC_OK;
}

int uapic_session_destroy(uapic_session_t *s) {
    if (!s) return UAPIC_OK;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->lc.stream);
    delete s;
    return UAPIC_OK;
}

int uapic_session_set_allreduce(uapic_session_t *s, uapic_allreduce_fn fn, void *ctx) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    s->reduce = fn; s->reduce_ctx = ctx;
    return UAPIC_OK;
}

int uapic_nccl_unique_id(void *id128) {
    if (!id128) return fail(UAPIC_EINVAL, "uapic_nccl_unique_id: null pointer");
    NcclApi *a = nullptr;
    TRY(nccl_need(&a));
    NcclId id;
    const int rc = a->GetUniqueId(&id);
    if (rc) return fail(UAPIC_ECUDA, "ncclGetUniqueId failed: %s", a->GetErrorString(rc));
    memcpy(id128, id.internal, sizeof(id.internal));
    return UAPIC_OK;
}

int uapic_nccl_version(int *version) {
    if (!version) return fail(UAPIC_EINVAL, "uapic_nccl_version: null pointer");
    NcclApi *a = nullptr;
    TRY(nccl_need(&a));
    const int rc = a->GetVersion(version);
    if (rc) return fail(UAPIC_ECUDA, "ncclGetVersion failed: %s", a->GetErrorString(rc));
    return UAPIC_OK;
}

static void session_drop_nccl(uapic_session *s) {
    if (s->nccl_comm && s->nccl_owned) { NcclApi *a = nccl_api(); if (a->h) a->CommDestroy(s->nccl_comm); }
    s->nccl_comm = nullptr; s->nccl_owned = false; s->nccl_rank = 0; s->nccl_nranks = 1;
}

int uapic_session_init_nccl(uapic_session_t *s, const void *id128, int nranks, int rank) {
    if (!s || !id128) return fail(UAPIC_EINVAL, "uapic_session_init_nccl: null pointer");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(UAPIC_EINVAL, "uapic_session_init_nccl: rank %d of %d", rank, nranks);
    NcclApi *a = nullptr;
    TRY(nccl_need(&a));
    TRY(session_bind(s));
    session_drop_nccl(s);
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void *comm = nullptr;
    const int rc = a->CommInitRank(&comm, nranks, id, rank);
    if (rc) return fail(UAPIC_ECUDA, "ncclCommInitRank(rank %d of %d) failed: %s", rank, nranks, a->GetErrorString(rc));
    s->nccl_comm = comm; s->nccl_owned = true; s->nccl_rank = rank; s->nccl_nranks = nranks;
    return UAPIC_OK;
}

int uapic_session_set_nccl_comm(uapic_session_t *s, void *comm) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (comm) { NcclApi *a = nullptr; TRY(nccl_need(&a)); }
    session_drop_nccl(s);
    s->nccl_comm = comm; s->nccl_owned = false;
    return UAPIC_OK;
}

int uapic_session_peer_handle(uapic_session_t *s, void *handle64) {
    if (!s || !handle64) return fail(UAPIC_EINVAL, "uapic_session_peer_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == UAPIC_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    TRY(session_bind(s));
    if (!s->xchg.p) {
        const size_t nrho = (size_t)s->m.ld * (s->m.ny + 1);
        TRY(session_alloc(s, s->xchg, 256 + 2 * (2 * nrho * 8)));
        TRY(session_alloc(s, s->peer_err, 256));
        CU(cudaMemset(s->xchg.p, 0, s->xchg.bytes));
        CU(cudaMemset(s->peer_err.p, 0, 256));
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->xchg.p));
    memcpy(handle64, &h, sizeof(h));
    return UAPIC_OK;
}

int uapic_session_init_peers(uapic_session_t *s, const void *handles, int nranks, int rank) {
    if (!s || !handles) return fail(UAPIC_EINVAL, "uapic_session_init_peers: null pointer");
    if (nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks) return fail(UAPIC_EINVAL, "uapic_session_init_peers: rank %d of %d (at most %d ranks)", rank, nranks, kMaxPeers);
    if (!s->xchg.p) return fail(UAPIC_ESTATE, "call uapic_session_peer_handle first (every rank), then exchange the handles");
    if (s->npeers) return fail(UAPIC_ESTATE, "peers are already attached");
    TRY(session_bind(s));
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { s->peer_base[r] = s->xchg.p; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (q != rank && s->peer_base[q]) { cudaIpcCloseMemHandle(s->peer_base[q]); s->peer_base[q] = nullptr; }
            return fail(UAPIC_ECUDA, "cudaIpcOpenMemHandle for rank %d failed: %s (peer access over NVLink is needed)", r, cudaGetErrorString(e));
        }
        s->peer_base[r] = p;
    }
    s->npeers = nranks; s->peer_rank = rank;
    return UAPIC_OK;
}

int uapic_session_close_peers(uapic_session_t *s) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    CU(cudaStreamSynchronize(s->lc.stream));
    int err = 0;
    if (s->peer_err.p) CU(cudaMemcpy(&err, s->peer_err.p, sizeof(int), cudaMemcpyDeviceToHost));
    for (int r = 0; r < s->npeers; ++r) if (r != s->peer_rank && s->peer_base[r]) { cudaIpcCloseMemHandle(s->peer_base[r]); s->peer_base[r] = nullptr; }
    s->npeers = 0;
    if (err) return fail(UAPIC_ECUDA, "a peer rank did not publish its deposits within 4 s during this session: results after that point are invalid");
    return UAPIC_OK;
}

int uapic_session_set_fusion(uapic_session_t *s, int enable) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (enable && s->cfg.storage_mode != UAPIC_STORE_ONEPASS_LEAN)
        return fail(UAPIC_EUNSUPPORTED, "phase fusion exists for storage_mode UAPIC_STORE_ONEPASS_LEAN only");
    TRY(session_bind(s));
    if (!enable) TRY(session_flush_b(s));
    if (enable && !s->fuse && s->sort_interval == 1) s->sort_interval = UAPIC_FUSE_SORT_INTERVAL;   // see the header
    if (!enable && s->fuse && s->sort_interval == UAPIC_FUSE_SORT_INTERVAL) s->sort_interval = 1;
    s->fuse = enable != 0;
    return UAPIC_OK;
}

int uapic_session_upload_particles(uapic_session_t *s, const double *x, const double *v) {
    if (!s || !x || !v) return fail(UAPIC_EINVAL, "uapic_session_upload_particles: null pointer");
    TRY(session_bind(s));
    s->pending_b = false;          // v is replaced: the pending compute_v of the old particles has no reader
    const size_t n = 16 * (size_t)s->cfg.nbpart;
    CU(cudaMemcpyAsync(s->x.p, x, n, cudaMemcpyHostToDevice, s->lc.stream));
    CU(cudaMemcpyAsync(s->v.p, v, n, cudaMemcpyHostToDevice, s->lc.stream));
    CU(cudaStreamSynchronize(s->lc.stream));
    if (s->permuted) {
        // x and v are back in the caller's order: bring particles.e along, then forget the permutation
        CU(launch_unpermute(s->lc, s->cfg.nbpart, s->perm.as<uint32_t>(), s->ep.as<double2>(), s->ep2.as<double2>()));
        s->ep.swap(s->ep2);
        s->permuted = false;
    }
    s->have_particles = true;
    return UAPIC_OK;
}

int uapic_session_upload_particle_e(uapic_session_t *s, const double *ep) {
    if (!s || !ep) return fail(UAPIC_EINVAL, "uapic_session_upload_particle_e: null pointer");
    TRY(session_bind(s));
    TRY(session_flush_b(s));
    if (s->permuted) {
        // the device arrays are in sorted order: undo it for x and v so that everything is in the caller's order again
        CU(launch_unpermute(s->lc, s->cfg.nbpart, s->perm.as<uint32_t>(), s->x.as<double2>(), s->x2.as<double2>()));
        CU(launch_unpermute(s->lc, s->cfg.nbpart, s->perm.as<uint32_t>(), s->v.as<double2>(), s->v2.as<double2>()));
        s->x.swap(s->x2); s->v.swap(s->v2);
        s->permuted = false;
    }
    CU(cudaMemcpyAsync(s->ep.p, ep, 16 * (size_t)s->cfg.nbpart, cudaMemcpyHostToDevice, s->lc.stream));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_set_sort(uapic_session_t *s, int interval, int bin_cells_log2) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (interval < 0 || bin_cells_log2 < 0 || bin_cells_log2 > 12) return fail(UAPIC_EINVAL, "uapic_session_set_sort: bad argument");
    if (sort_bins(s->m, bin_cells_log2) > 4096) return fail(UAPIC_EUNSUPPORTED, "more than 4096 sort bins: use larger bins");
    if (s->cfg.nbpart >= ((int64_t)1 << 32)) return fail(UAPIC_EUNSUPPORTED, "sorting needs nbpart < 2^32 per session");
    s->sort_interval = interval; s->sort_shift = bin_cells_log2;
    return UAPIC_OK;
}

namespace {
int drain_timing(uapic_session *s) {
    if (s->ev.empty()) return UAPIC_OK;
    CU(cudaStreamSynchronize(s->lc.stream));
    for (size_t i = 0; i + 3 < s->ev.size(); i += 4) {
        float a = 0, b = 0, f = 0;
        CU(cudaEventElapsedTime(&a, s->ev[i], s->ev[i + 1]));
        CU(cudaEventElapsedTime(&b, s->ev[i + 2], s->ev[i + 3]));
        CU(cudaEventElapsedTime(&f, s->ev[i + 1], s->ev[i + 2]));
        s->ms_a += a; s->ms_b += b; s->ms_f += f; s->timed_steps++;
    }
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    s->ev.clear();
    return UAPIC_OK;
}
}  // namespace

int uapic_session_enable_timing(uapic_session_t *s, int enable) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    TRY(drain_timing(s));
    s->timing = enable != 0;
    s->ms_a = s->ms_b = s->ms_f = 0; s->timed_steps = 0;
    return UAPIC_OK;
}

int uapic_session_phase_times(uapic_session_t *s, double *ms_phase_a, double *ms_phase_b, int64_t *steps) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    TRY(drain_timing(s));
    if (ms_phase_a) *ms_phase_a = s->ms_a;
    if (ms_phase_b) *ms_phase_b = s->ms_b;
    if (steps) *steps = s->timed_steps;
    s->ms_field_last = s->ms_f;
    s->ms_a = s->ms_b = s->ms_f = 0; s->timed_steps = 0;
    return UAPIC_OK;
}

int uapic_session_field_barrier_time(uapic_session_t *s, double *ms_field_barrier) {
    if (!s || !ms_field_barrier) return fail(UAPIC_EINVAL, "null pointer");
    *ms_field_barrier = s->ms_field_last;
    return UAPIC_OK;
}

int uapic_session_generate_particles_strided(uapic_session_t *s, int kind, uint64_t seed, int64_t first_global_index,
                                             int64_t index_stride, double alpha, double kx) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (kind != 0 && kind != 1) return fail(UAPIC_EINVAL, "unknown load kind %d", kind);
    if (index_stride < 1 || first_global_index < 0) return fail(UAPIC_EINVAL, "bad particle index range");
    TRY(session_bind(s));
    s->pending_b = false;
    CU(launch_generate(s->lc, s->m, kind, seed, first_global_index, index_stride, s->cfg.nbpart, s->np_global, alpha, kx,
                       s->x.as<double>(), s->v.as<double>()));
    s->permuted = false;
    s->have_particles = true;
    return UAPIC_OK;
}

int uapic_session_generate_particles(uapic_session_t *s, int kind, uint64_t seed, int64_t first_global_index, double alpha,
                                     double kx) {
    return uapic_session_generate_particles_strided(s, kind, seed, first_global_index, 1, alpha, kx);
}

int uapic_session_init_fields(uapic_session_t *s) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (!s->have_particles) return fail(UAPIC_ESTATE, "upload or generate particles before uapic_session_init_fields");
    TRY(session_bind(s));
    s->n_energy = 0;
    TRY(session_reserve_energy(s, 1));
    TRY(session_clear_raw(s));
    CU(launch_deposit(s->lc, s->m, s->cfg.nbpart, s->x.as<double>(), s->cfg.weight, s->acc, s->cfg.wrap, s->cfg.scheme));   // bupdate.F90:89
    TRY(session_field_solve(s));                                                                              // :91
    CU(launch_gather(s->lc, s->m, s->emesh.as<double>(), s->cfg.nbpart, s->x.as<double>(), s->ep.as<double>(), s->cfg.wrap, s->cfg.scheme));  // :93
    s->fields_ready = true;
    return UAPIC_OK;
}

int uapic_session_step(uapic_session_t *s, int nsteps) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    if (!s->fields_ready) return fail(UAPIC_ESTATE, "call uapic_session_init_fields before uapic_session_step");
    if (nsteps < 0) return fail(UAPIC_EINVAL, "nsteps must be >= 0");
    TRY(session_bind(s));
    TRY(session_reserve_energy(s, 2 * (int64_t)nsteps));
    for (int it = 0; it < nsteps; ++it) {
        if (s->sort_interval > 0 && s->steps_done % s->sort_interval == 0) {
            bool sort_on = false;
            TRY(session_want_sort(s, &sort_on));
            if (sort_on) {
                TRY(session_flush_b(s));      // the store is in the old order: its phase B has to run before the particles move
                TRY(session_sort(s));
            }
        }
        s->steps_done++;
        const PhaseParams p = session_params(s);
        OnepassParams op = s->onepass ? session_onepass_params(s) : OnepassParams{};
        cudaEvent_t e4[4] = {nullptr, nullptr, nullptr, nullptr};
        if (s->timing) {
            if (s->ev.size() >= 4096) TRY(drain_timing(s));
            for (int q = 0; q < 4; ++q) { CU(cudaEventCreate(&e4[q])); s->ev.push_back(e4[q]); }
        }
        TRY(session_clear_raw(s));
        if (s->onepass) {
            // one field barrier per step: both deposits come out of the first kernel (uapic_onepass.cu)
            if (s->timing) CU(cudaEventRecord(e4[0], s->lc.stream));
            op.ehalo = s->ehalo.as<double2>();
            op.ehalo_b = s->ehalo_p.as<double2>();
            op.fuse_b = s->pending_b ? 1 : 0;                              // phase B of the previous step rides along
            CU(launch_onepass_a(s->lc, op));                               // bupdate.F90:97-106, :112-117 (x part)
            s->pending_b = false;
            if (s->timing) CU(cudaEventRecord(e4[1], s->lc.stream));
            TRY(session_field_barrier(s, 2));      // :108 predictor field and :119 field of the next step, one exchange + one launch
            if (s->timing) CU(cudaEventRecord(e4[2], s->lc.stream));
            if (s->fuse) {
                s->pending_b = true;                                       // :110-115 (y part), :123 run inside the next kernel
            } else {
                op.ehalo = s->ehalo_p.as<double2>();
                op.fuse_b = 0;
                CU(launch_onepass_b(s->lc, op));                           // :110-115 (y part), :123
            }
            if (s->timing) CU(cudaEventRecord(e4[3], s->lc.stream));
            continue;
        }
        if (s->generic) {
            // any even ntau: the literal call sequence of bupdate.F90:97-123 on the reference-shaped arrays (uapic_generic.cu)
            const int N = s->cfg.ntau;
            const double eps = s->cfg.eps;
            const int64_t np = s->cfg.nbpart;
            double *b = s->gb.as<double>(), *t = s->gt.as<double>(), *pl = s->gpl.as<double>(), *ql = s->gql.as<double>();
            double *xt = s->gxt.as<double>(), *yt = s->gyt.as<double>(), *xf = s->gxf.as<double>(), *yf = s->gyf.as<double>();
            double *fx = s->gfx.as<double>(), *fy = s->gfy.as<double>(), *gx = s->ggx.as<double>(), *gy = s->ggy.as<double>(), *et = s->get.as<double>();
            if (s->timing) CU(cudaEventRecord(e4[0], s->lc.stream));
            CU(launch_preparation(s->lc, N, eps, s->cfg.dt, np, s->x.as<double>(), s->v.as<double>(), s->ep.as<double>(), b, t, pl, ql, xt, yt));   // :97
            CU(launch_gather_tau(s->lc, s->m, s->emesh.as<double>(), N, np, xt, et, s->cfg.wrap));                                   // :99
            CU(launch_compute_f(s->lc, N, eps, np, b, xt, yt, et, fx, fy, 1));                                                      // :101
            CU(launch_step_fortran(s->lc, N, eps, np, t, pl, nullptr, xt, xf, fx, nullptr, 0));                                     // :103
            CU(launch_step_fortran(s->lc, N, eps, np, t, pl, nullptr, yt, yf, fy, nullptr, 0));                                     // :104
            CU(launch_deposit_tau(s->lc, s->m, N, eps, np, xt, t, s->cfg.weight, s->acc, s->x.as<double>(), s->cfg.wrap));          // :106
            if (s->timing) CU(cudaEventRecord(e4[1], s->lc.stream));
            TRY(session_field_solve(s));                                                                                            // :108
            TRY(session_clear_raw(s));
            if (s->timing) CU(cudaEventRecord(e4[2], s->lc.stream));
            CU(launch_gather_tau(s->lc, s->m, s->emesh.as<double>(), N, np, xt, et, s->cfg.wrap));                                   // :110
            CU(launch_compute_f(s->lc, N, eps, np, b, xt, yt, et, gx, gy, 1));                                                      // :112
            CU(launch_step_fortran(s->lc, N, eps, np, t, pl, ql, xt, xf, fx, gx, 1));                                               // :114
            CU(launch_step_fortran(s->lc, N, eps, np, t, pl, ql, yt, yf, fy, gy, 1));                                               // :115
            CU(launch_deposit_tau(s->lc, s->m, N, eps, np, xt, t, s->cfg.weight, s->acc, s->x.as<double>(), s->cfg.wrap));          // :117
            CU(launch_compute_v(s->lc, N, eps, np, t, yt, 0, s->v.as<double>()));                                                   // :123 (needs no field)
            if (s->timing) CU(cudaEventRecord(e4[3], s->lc.stream));
            TRY(session_field_solve(s));                                                                                            // :119
            continue;
        }
        if (s->timing) CU(cudaEventRecord(e4[0], s->lc.stream));
        CU(launch_phase_a(s->lc, p));          // bupdate.F90:97-106
        if (s->timing) CU(cudaEventRecord(e4[1], s->lc.stream));
        TRY(session_field_solve(s));           // :108
        TRY(session_clear_raw(s));
        if (s->timing) CU(cudaEventRecord(e4[2], s->lc.stream));
        CU(launch_phase_b(s->lc, p));          // :110-117, :123
        if (s->timing) CU(cudaEventRecord(e4[3], s->lc.stream));
        TRY(session_field_solve(s));           // :119
    }
    return UAPIC_OK;
}

// One UA step with the particles living in HOST memory (the end-to-end pattern): the particle range is cut into chunks;
// chunk c's upload overlaps phase A of chunk c-1, and chunk c's download overlaps phase B of chunk c+1.  Same result as
// upload_particles + upload_particle_e + step(1) + download_particles.
int uapic_session_step_host(uapic_session_t *s, const double *x_in, const double *v_in, const double *e_in, double *x_out,
                            double *v_out) {
    if (!s || !x_in || !v_in || !x_out || !v_out) return fail(UAPIC_EINVAL, "uapic_session_step_host: null pointer");
    if (!s->fields_ready) return fail(UAPIC_ESTATE, "call uapic_session_init_fields before uapic_session_step_host");
    TRY(session_bind(s));
    s->pending_b = false;          // x, v (and e) come from the caller: nothing on the device is carried over
    const int64_t np = s->cfg.nbpart;
    if (!s->onepass || np < (1 << 16)) {
        // two-barrier kernels / tiny problems: nothing to overlap
        TRY(uapic_session_upload_particles(s, x_in, v_in));
        if (e_in) TRY(uapic_session_upload_particle_e(s, e_in));
        TRY(uapic_session_step(s, 1));
        return uapic_session_download_particles(s, x_out, v_out);
    }
    constexpr int kMaxChunks = UAPIC_HOST_CHUNKS;
    // chunks of at least 2^18 particles (smaller ones cost more in launch tails than their overlap hides), at least 2
    const int kChunks = (int)std::min<int64_t>(kMaxChunks, std::max<int64_t>(2, np >> 18));
    if (!s->up_stream) {
        CU(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&s->down_stream, cudaStreamNonBlocking));
        for (int q = 0; q < 3 * kMaxChunks + 2; ++q) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->chunk_ev.push_back(e);
        }
    }
    if (s->permuted && !e_in) {
        // particles.e stays on the device in sorted order: bring it back to the caller's order first
        CU(launch_unpermute(s->lc, np, s->perm.as<uint32_t>(), s->ep.as<double2>(), s->ep2.as<double2>()));
        s->ep.swap(s->ep2);
    }
    s->permuted = false;
    TRY(session_reserve_energy(s, 2));
    bool sort = false;
    TRY(session_want_sort(s, &sort));
    cudaStream_t cs = s->lc.stream;
    cudaEvent_t ev_start = s->chunk_ev[3 * kMaxChunks], ev_done = s->chunk_ev[3 * kMaxChunks + 1];
    CU(cudaEventRecord(ev_start, cs));                       // earlier work on the session stream (previous step) is finished
    CU(cudaStreamWaitEvent(s->up_stream, ev_start, 0));
    CU(cudaStreamWaitEvent(s->down_stream, ev_start, 0));
    TRY(session_clear_raw(s));
    const int64_t per = ((np + kChunks - 1) / kChunks + 255) / 256 * 256;
    OnepassParams op = session_onepass_params(s);
    const size_t stride = onepass_store_bytes_per_particle(op.ntau, op.full);
    // ---- uploads || reorder + phase A, chunk by chunk ----
    for (int c = 0; c < kChunks; ++c) {
        const int64_t lo = (int64_t)c * per, n = std::min(per, np - lo);
        if (n <= 0) break;
        CU(cudaMemcpyAsync(s->x.as<double2>() + lo, x_in + 2 * lo, 16 * (size_t)n, cudaMemcpyHostToDevice, s->up_stream));
        CU(cudaMemcpyAsync(s->v.as<double2>() + lo, v_in + 2 * lo, 16 * (size_t)n, cudaMemcpyHostToDevice, s->up_stream));
        if (e_in) CU(cudaMemcpyAsync(s->ep.as<double2>() + lo, e_in + 2 * lo, 16 * (size_t)n, cudaMemcpyHostToDevice, s->up_stream));
        CU(cudaEventRecord(s->chunk_ev[c], s->up_stream));
        CU(cudaStreamWaitEvent(cs, s->chunk_ev[c], 0));
        OnepassParams pc = op;
        pc.np = n;
        if (sort) {
            CU(launch_sort_particles(s->lc, s->m, s->sort_shift, n, s->x.as<double2>() + lo, s->v.as<double2>() + lo,
                                     s->ep.as<double2>() + lo, nullptr, s->x2.as<double2>() + lo, s->v2.as<double2>() + lo,
                                     s->ep2.as<double2>() + lo, s->perm2.as<uint32_t>() + lo, s->binid.as<uint16_t>() + lo,
                                     s->hist.as<unsigned>(), (uint32_t)lo));
            pc.x = s->x2.as<double2>() + lo; pc.v = s->v2.as<double2>() + lo; pc.ep = s->ep2.as<double2>() + lo;
            // the new x goes straight back to the caller's order (perm2 holds global indices, all inside this chunk)
            pc.out_perm = s->perm2.as<uint32_t>() + lo; pc.x_out = s->x.as<double2>(); pc.v_out = s->v.as<double2>();
        } else {
            pc.x = s->x.as<double2>() + lo; pc.v = s->v.as<double2>() + lo; pc.ep = s->ep.as<double2>() + lo;
        }
        pc.store = op.store + (size_t)lo * stride;
        pc.rec = op.rec + 8 * lo;
        pc.ehalo = s->ehalo.as<double2>();
        CU(launch_onepass_a(s->lc, pc));
        // the new x is final after phase A (corrector deposit position, compute_rho_m6.F90:86-87) and already in the caller's
        // order: send it down now, while the device-to-host direction of the link is idle -- only v has to wait for phase B
        CU(cudaEventRecord(s->chunk_ev[2 * kMaxChunks + c], cs));
        CU(cudaStreamWaitEvent(s->down_stream, s->chunk_ev[2 * kMaxChunks + c], 0));
        CU(cudaMemcpyAsync(x_out + 2 * lo, s->x.as<double2>() + lo, 16 * (size_t)n, cudaMemcpyDeviceToHost, s->down_stream));
    }
    // ---- the one field barrier ----
    TRY(session_field_barrier(s, 2));
    // ---- phase B (+ undo the reordering) || downloads, chunk by chunk ----
    for (int c = 0; c < kChunks; ++c) {
        const int64_t lo = (int64_t)c * per, n = std::min(per, np - lo);
        if (n <= 0) break;
        OnepassParams pc = op;
        pc.np = n;
        pc.x = (sort ? s->x2 : s->x).as<double2>() + lo; pc.v = (sort ? s->v2 : s->v).as<double2>() + lo;
        pc.ep = (sort ? s->ep2 : s->ep).as<double2>() + lo;
        pc.store = op.store + (size_t)lo * stride;
        pc.rec = op.rec + 8 * lo;
        pc.ehalo = s->ehalo_p.as<double2>();
        if (sort) { pc.out_perm = s->perm2.as<uint32_t>() + lo; pc.x_out = s->x.as<double2>(); pc.v_out = s->v.as<double2>(); }
        CU(launch_onepass_b(s->lc, pc));
        CU(cudaEventRecord(s->chunk_ev[kMaxChunks + c], cs));
        CU(cudaStreamWaitEvent(s->down_stream, s->chunk_ev[kMaxChunks + c], 0));
        CU(cudaMemcpyAsync(v_out + 2 * lo, s->v.as<double2>() + lo, 16 * (size_t)n, cudaMemcpyDeviceToHost, s->down_stream));
    }
    CU(cudaEventRecord(ev_done, s->down_stream));
    CU(cudaStreamWaitEvent(cs, ev_done, 0));
    CU(cudaStreamSynchronize(cs));
    s->steps_done++;
    return UAPIC_OK;
}

int uapic_session_synchronize(uapic_session_t *s) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_download_particles(uapic_session_t *s, double *x, double *v) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    TRY(session_flush_b(s));
    if (x) TRY(session_download_pairs(s, s->x, s->x2, x));
    if (v) TRY(session_download_pairs(s, s->v, s->v2, v));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_download_particle_e(uapic_session_t *s, double *ep) {
    if (!s || !ep) return fail(UAPIC_EINVAL, "null pointer");
    TRY(session_bind(s));
    TRY(session_download_pairs(s, s->ep, s->ep2, ep));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_download_fields(uapic_session_t *s, double *e, double *rho) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    const size_t nrho = (size_t)s->m.ld * (s->m.ny + 1);
    if (e) CU(cudaMemcpyAsync(e, s->emesh.p, 16 * nrho, cudaMemcpyDeviceToHost, s->lc.stream));
    if (rho) CU(cudaMemcpyAsync(rho, s->rho.p, 8 * nrho, cudaMemcpyDeviceToHost, s->lc.stream));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_energy_history(uapic_session_t *s, double *out, int64_t capacity, int64_t *count) {
    if (!s) return fail(UAPIC_EINVAL, "session is null");
    TRY(session_bind(s));
    if (count) *count = s->n_energy;
    if (out) {
        const int64_t n = s->n_energy < capacity ? s->n_energy : capacity;
        if (n > 0) CU(cudaMemcpyAsync(out, s->energy.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, s->lc.stream));
        CU(cudaStreamSynchronize(s->lc.stream));
    }
    return UAPIC_OK;
}

int uapic_session_sum_v(uapic_session_t *s, double *sumv2) {
    if (!s || !sumv2) return fail(UAPIC_EINVAL, "null pointer");
    TRY(session_bind(s));
    TRY(session_flush_b(s));
    CU(launch_sum_v(s->lc, s->cfg.nbpart, s->v.as<double>(), s->sumv.as<double>() + 2, s->sumv.as<double>()));
    CU(cudaMemcpyAsync(sumv2, s->sumv.p, 16, cudaMemcpyDeviceToHost, s->lc.stream));
    CU(cudaStreamSynchronize(s->lc.stream));
    return UAPIC_OK;
}

int uapic_session_launch_count(uapic_session_t *s, int64_t *count) {
    if (!s || !count) return fail(UAPIC_EINVAL, "null pointer");
    *count = s->launches;
    return UAPIC_OK;
}

int uapic_session_device_bytes(uapic_session_t *s, int64_t *bytes) {
    if (!s || !bytes) return fail(UAPIC_EINVAL, "null pointer");
    *bytes = s->bytes;
    return UAPIC_OK;
}

}  // extern "C"

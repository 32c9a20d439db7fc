// uapic_mrc3d.cu -- the sibling scheme of the reference: the 3D rotation-push PIC of
// fortran/uapic3d.f90 with CIC deposition / interpolation (compute_rho_cic.f90, interpolation_cic.f90) and the 3D periodic
// spectral Poisson solve (poisson_3d.f90), sm_100a.  Self-contained: kernels + C ABI (uapic3d_* in include/uapic_b200.h).
//
// Reference behaviour kept on purpose ("results identical to the reference's on the same inputs"):
//  * positions are divided by dx without subtracting xmin (compute_rho_cic.f90:34-36, interpolation_cic.f90:30-32);
//  * charge deposited on the ghost planes i = nx+1, j = ny+1, k = nz+1 is OVERWRITTEN by the periodic copy of plane 1
//    (compute_rho_cic.f90:69-71), not folded back;
//  * E = -i k rho_hat / k^2 per component, real part, ghost planes, then 1/(nx ny nz) (poisson_3d.f90:47-191);
//  * the second half-push of the N0mrc <= 1 branch does not wrap (uapic3d.f90:123);
//  * the beta loop of the MRC branch reads p%x(m,1) -- the m-th element of the flattened (3, nbpart) array -- where
//    p%x(1,m) was meant (uapic3d.f90:179,182): `index_quirk = 1` (default) reproduces it, 0 uses x(1,m).
// Everything here is HBM/latency-bound mesh and particle streaming; nothing is shaped like a contraction.
#include <cmath>
#include <cstdlib>
#include <new>

#include "../../include/uapic_b200.h"
#include "uapic_internal.h"
#include "uapic_mesh.cuh"

namespace uapic {

namespace {

constexpr int k3Block = 256;

struct Mesh3 {
    double xmin[3], dim[3], d[3];
    int n[3];
    DEVINL int ld1() const { return n[0] + 1; }
    DEVINL int ld2() const { return n[1] + 1; }
    __host__ __device__ size_t nodes() const { return (size_t)(n[0] + 1) * (n[1] + 1) * (n[2] + 1); }
    __host__ __device__ size_t cells() const { return (size_t)n[0] * n[1] * n[2]; }
};

inline int grid3(int sm, int64_t items) {
    int64_t need = (items + k3Block - 1) / k3Block;
    if (need < 1) need = 1;
    const int64_t cap = (int64_t)sm * 8;
    return (int)(need < cap ? need : cap);
}

// push_particles (uapic3d.f90:224-242): x = xmin + modulo(x + dt v - xmin, dim)
__global__ void __launch_bounds__(k3Block) k3_push(Mesh3 m, int64_t np, double *x, const double *v, double delta_t, int wrap) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < 3 * np; q += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(q % 3);
        if (wrap) {
            const double dd = __dsub_rn(__dmul_rn(delta_t, v[q]), m.xmin[c]);
            x[q] = __dadd_rn(m.xmin[c], modulo_exact(__dadd_rn(x[q], dd), m.dim[c]));
        } else {
            x[q] = __dadd_rn(x[q], __dmul_rn(delta_t, v[q]));          // p%x = p%x + 0.5 dt p%v   (uapic3d.f90:123)
        }
    }
}

struct Cic3 { int i, j, k; double a[8]; };
DEVINL Cic3 cic3(const Mesh3 &m, const double *x, int64_t p) {     // compute_rho_cic.f90:32-53 = interpolation_cic.f90:30-51
    const double xp = __ddiv_rn(x[3 * p], m.d[0]), yp = __ddiv_rn(x[3 * p + 1], m.d[1]), zp = __ddiv_rn(x[3 * p + 2], m.d[2]);
    Cic3 c;
    c.i = (int)floor(xp); c.j = (int)floor(yp); c.k = (int)floor(zp);
    const double dx = __dsub_rn(xp, (double)c.i), dy = __dsub_rn(yp, (double)c.j), dz = __dsub_rn(zp, (double)c.k);
    const double ox = __dsub_rn(1.0, dx), oy = __dsub_rn(1.0, dy), oz = __dsub_rn(1.0, dz);
    c.a[0] = __dmul_rn(__dmul_rn(ox, oy), oz); c.a[1] = __dmul_rn(__dmul_rn(dx, oy), oz);
    c.a[2] = __dmul_rn(__dmul_rn(ox, dy), oz); c.a[3] = __dmul_rn(__dmul_rn(dx, dy), oz);
    c.a[4] = __dmul_rn(__dmul_rn(ox, oy), dz); c.a[5] = __dmul_rn(__dmul_rn(dx, oy), dz);
    c.a[6] = __dmul_rn(__dmul_rn(ox, dy), dz); c.a[7] = __dmul_rn(__dmul_rn(dx, dy), dz);
    // memory safety only (the reference has undefined behaviour for a position outside the box)
    c.i = min(max(c.i, 0), m.n[0] - 1); c.j = min(max(c.j, 0), m.n[1] - 1); c.k = min(max(c.k, 0), m.n[2] - 1);
    return c;
}
DEVINL size_t node3(const Mesh3 &m, int i, int j, int k) { return (size_t)i + (size_t)m.ld1() * ((size_t)j + (size_t)m.ld2() * k); }

__global__ void __launch_bounds__(k3Block) k3_deposit(Mesh3 m, int64_t np, const double *x, double vol, RhoAcc acc) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
        const Cic3 c = cic3(m, x, p);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const size_t idx = node3(m, c.i + (q & 1), c.j + ((q >> 1) & 1), c.k + (q >> 2));
            const double val = __dmul_rn(c.a[q], vol);
            if (acc.i64) atomicAdd(acc.i64 + idx, (unsigned long long)__double2ll_rn(__dmul_rn(val, acc.scale)));
            else atomicAdd(acc.f64 + idx, val);
        }
    }
}

// periodic copies in the reference's order x, y, z (compute_rho_cic.f90:69-71): the final value of node (i,j,k) is the raw
// value of (i mod nx, j mod ny, k mod nz) with "mod" applied to the ghost index only
__global__ void __launch_bounds__(k3Block) k3_rho_finish(Mesh3 m, RhoAcc acc, double *rho) {
    const size_t n = m.nodes();
    const double inv = acc.i64 ? 1.0 / acc.scale : 1.0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(q % m.ld1()), j = (int)((q / m.ld1()) % m.ld2()), k = (int)(q / ((size_t)m.ld1() * m.ld2()));
        if (i == m.n[0]) i = 0;
        if (j == m.n[1]) j = 0;
        if (k == m.n[2]) k = 0;
        const size_t s = node3(m, i, j, k);
        rho[q] = acc.i64 ? (double)(long long)acc.i64[s] * inv : acc.f64[s];
    }
}

// ---- 3D FFT passes on a complex nx*ny*nz array (x fastest) ---------------------------------------------------------------
// dir 0/1/2 = along x/y/z; one CTA per line.  mode 0: in place.  mode 1 (dir 0, forward): load Re = rho(i,j,k), Im = 0.
// mode 2 (dir 0, backward): load A multiplied by -i k_comp / k^2 (poisson_3d.f90:86-92), write to B.
__global__ void k3_fft_pass(Mesh3 m, int dir, int sign, int mode, int comp, const double *__restrict__ rho, const double2 *__restrict__ A,
                            double2 *__restrict__ B) {
    extern __shared__ double2 smem_raw[];
    const int nx = m.n[0], ny = m.n[1], nz = m.n[2];
    const int len = m.n[dir];
    cd *a = reinterpret_cast<cd *>(smem_raw), *tmp = a + len, *tw = tmp + len;
    const int line = blockIdx.x;
    size_t base, stride;
    int l0, l1;          // the two fixed indices of the line
    if (dir == 0) { l0 = line % ny; l1 = line / ny; base = (size_t)nx * (l0 + (size_t)ny * l1); stride = 1; }
    else if (dir == 1) { l0 = line % nx; l1 = line / nx; base = l0 + (size_t)nx * ny * l1; stride = nx; }
    else { l0 = line % nx; l1 = line / nx; base = l0 + (size_t)nx * l1; stride = (size_t)nx * ny; }
    line_twiddles(tw, len);
    const double pi = 3.14159265358979323846;
    for (int t = threadIdx.x; t < len; t += blockDim.x) {
        if (mode == 1) {
            a[t] = mk(rho[node3(m, t, l0, l1)], 0.0);
        } else if (mode == 2) {
            const double2 r = A[base + t];
            const int ix = t, iy = l0, iz = l1;
            const int kx_i = ix < nx / 2 ? ix : ix - nx, ky_i = iy < ny / 2 ? iy : iy - ny, kz_i = iz < nz / 2 ? iz : iz - nz;   // :64-85 (1-based i <= n/2)
            if (kx_i == 0 && ky_i == 0 && kz_i == 0) { a[t] = mk(0.0, 0.0); continue; }
            const double kx = 2.0 * pi * (double)kx_i / m.dim[0], ky = 2.0 * pi * (double)ky_i / m.dim[1], kz = 2.0 * pi * (double)kz_i / m.dim[2];
            const double kk = comp == 0 ? kx : (comp == 1 ? ky : kz), k2 = kx * kx + ky * ky + kz * kz;
            // -(0, kk) * (r.x, r.y) / k2
            a[t] = mk((kk * r.y) / k2, (-kk * r.x) / k2);
        } else {
            const double2 r = B[base + t * stride];
            a[t] = mk(r.x, r.y);
        }
    }
    line_fft(a, tmp, tw, len, sign);
    for (int t = threadIdx.x; t < len; t += blockDim.x) B[base + t * stride] = make_double2(a[t].re, a[t].im);
}

// fields%e(comp, :, :, :) = real(psi), periodic planes, / (nx ny nz)      poisson_3d.f90:96, :184-189
__global__ void __launch_bounds__(k3Block) k3_store_e(Mesh3 m, int comp, const double2 *__restrict__ B, double *__restrict__ e) {
    const size_t n = m.nodes();
    const double sc = (double)m.cells();
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(q % m.ld1()), j = (int)((q / m.ld1()) % m.ld2()), k = (int)(q / ((size_t)m.ld1() * m.ld2()));
        if (i == m.n[0]) i = 0;
        if (j == m.n[1]) j = 0;
        if (k == m.n[2]) k = 0;
        e[comp + 3 * q] = B[i + (size_t)m.n[0] * (j + (size_t)m.n[1] * k)].x / sc;
    }
}

// interpolate_eb_cic                                   interpolation_cic.f90:26-62
__global__ void __launch_bounds__(k3Block) k3_gather(Mesh3 m, int64_t np, const double *x, const double *__restrict__ e, double *ep) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
        const Cic3 c = cic3(m, x, p);
        size_t idx[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) idx[q] = 3 * node3(m, c.i + (q & 1), c.j + ((q >> 1) & 1), c.k + (q >> 2));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double s = __dmul_rn(c.a[0], e[idx[0] + d]);
#pragma unroll
            for (int q = 1; q < 8; ++q) s = __dadd_rn(s, __dmul_rn(c.a[q], e[idx[q] + d]));
            ep[3 * p + d] = s;
        }
    }
}

// the velocity rotations of uapic3d.f90:103-121 (kind 0), :141-161 (kind 1, alpha), :175-196 (kind 2, beta)
__global__ void __launch_bounds__(k3Block) k3_rotate(int kind, int64_t np, const double *x, double *v, const double *ep, double dt, double eps,
                                                     double coef, double delta, int index_quirk) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
        const double x1 = x[3 * p], x2 = x[3 * p + 1];
        const double xa = (kind == 2 && index_quirk) ? x[p] : x1;        // p%x(m,1): element m of the flattened array (uapic3d.f90:179,182)
        const double d1 = x1 - 9.0, d2 = x2 - 9.0, da = xa - 9.0;
        const double r12 = 1.0 + (da * da + d2 * d2) * delta * delta;     // Bm(1), Bm(2) (with the quirk in the beta loop)
        const double r3 = 1.0 + (d1 * d1 + d2 * d2) * delta * delta;      // Bm(3)
        const double B1 = d2 * delta / sqrt(r12), B2 = -d1 * delta / sqrt(r12), B3 = 1.0 / sqrt(r3);
        const double v1 = v[3 * p], v2 = v[3 * p + 1], v3 = v[3 * p + 2];
        const double E1 = ep[3 * p], E2 = ep[3 * p + 1], E3 = ep[3 * p + 2];
        const double vxB1 = v2 * B3 - v3 * B2, vxB2 = v3 * B1 - v1 * B3, vxB3 = v1 * B2 - v2 * B1;     // cross(v, B)
        const double ExB1 = E2 * B3 - E3 * B2, ExB2 = E3 * B1 - E1 * B3, ExB3 = E1 * B2 - E2 * B1;     // cross(E, B)
        const double EB = B1 * E1 + B2 * E2 + B3 * E3, vB = B1 * v1 + B2 * v2 + B3 * v3;
        double c0, c1, c2, c3, c4, c5;       // v' = c0 v + c1 vxB + c2 E + c3 (E.B) B + c4 ExB + c5 (v.B) B
        if (kind == 0) {
            const double a = dt / eps, s = sin(a), c = cos(a);
            c0 = c; c1 = s; c2 = eps * s; c3 = dt - eps * s; c4 = eps - eps * c; c5 = 1.0 - c;
        } else if (kind == 1) {
            const double s = sin(dt), c = cos(dt);
            c0 = c; c1 = s; c2 = coef * s; c3 = coef * (dt - s); c4 = coef * (1.0 - c); c5 = 1.0 - c;
        } else {
            const double s = sin(-dt), c = cos(dt);
            c0 = c; c1 = s; c2 = -coef * s; c3 = -coef * (-dt - s); c4 = -coef * (1.0 - c); c5 = 1.0 - c;
        }
        v[3 * p]     = c0 * v1 + c1 * vxB1 + c2 * E1 + c3 * EB * B1 + c4 * ExB1 + c5 * vB * B1;
        v[3 * p + 1] = c0 * v2 + c1 * vxB2 + c2 * E2 + c3 * EB * B2 + c4 * ExB2 + c5 * vB * B2;
        v[3 * p + 2] = c0 * v3 + c1 * vxB3 + c2 * E3 + c3 * EB * B3 + c4 * ExB3 + c5 * vB * B3;
    }
}

DEVINL uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
DEVINL double u01(uint64_t seed, uint64_t particle, uint32_t stream, uint32_t draw) {
    uint64_t h = mix64(seed ^ mix64(particle * 0xD1342543DE82EF95ull + stream));
    h = mix64(h + draw);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
// init_particles_3d densities (particles.F90:152-190) from a counter-based stream keyed by the particle index
__global__ void __launch_bounds__(k3Block) k3_generate(Mesh3 m, uint64_t seed, int64_t first, int64_t np, double *x, double *v) {
    const double pi = 3.14159265358979323846;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t id = (uint64_t)(first + p);
        x[3 * p + 2] = m.xmin[2] + m.dim[2] * u01(seed, id, 3, 0);
        for (uint32_t d = 0;; d += 3) {
            const double xi = 9.0 * u01(seed, id, 4, d), yi = 2.0 * pi * u01(seed, id, 4, d + 1), zi = (1.0 + 0.02) * u01(seed, id, 4, d + 2);
            if ((1.0 + 0.02 * cos(4.0 * yi)) * exp(-5.0 * (xi - 4.8) * (xi - 4.8)) >= zi) { x[3 * p] = cos(yi) * xi + 9.0; x[3 * p + 1] = sin(yi) * xi + 9.0; break; }
        }
        for (uint32_t d = 0;; d += 4) {
            const double xi = (u01(seed, id, 5, d) - 0.5) * 8.0, yi = (u01(seed, id, 5, d + 1) - 0.5) * 8.0, wi = (u01(seed, id, 5, d + 2) - 0.5) * 8.0;
            if (exp(-2.0 * (xi * xi + yi * yi + wi * wi)) >= u01(seed, id, 5, d + 3)) { v[3 * p] = xi; v[3 * p + 1] = yi; v[3 * p + 2] = wi; break; }
        }
    }
}

}  // namespace

}  // namespace uapic

// =====================================================================================================================
// C ABI
// =====================================================================================================================
using namespace uapic;

struct uapic3d_session {
    uapic3d_config_t cfg;
    Mesh3 m;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    double *x = nullptr, *v = nullptr, *ep = nullptr, *rho = nullptr, *e = nullptr;
    void *raw = nullptr;
    double2 *A = nullptr, *B = nullptr;
    RhoAcc acc{};
    bool have_particles = false, fields_ready = false;
    bool own_stream = false;
    void *nccl_comm = nullptr;         // particles sharded over ranks: the raw rho nodes are summed before the periodic copies
    // one captured sub-step per kind (uapic3d_substep): the loop body is ~22 launches of microsecond kernels, i.e. launch bound
    cudaGraphExec_t gexec[3] = {nullptr, nullptr, nullptr};
    double gdt[3] = {0, 0, 0}, gcoef[3] = {0, 0, 0};
    int64_t glaunches[3] = {0, 0, 0};
    ~uapic3d_session() {
        for (cudaGraphExec_t g : gexec) if (g) cudaGraphExecDestroy(g);
        if (nccl_comm) uapic_internal_nccl_comm_destroy(nccl_comm);
        if (own_stream && stream) cudaStreamDestroy(stream);
        for (void *p : {(void *)x, (void *)v, (void *)ep, (void *)rho, (void *)e, raw, (void *)A, (void *)B}) if (p) cudaFree(p);
    }
};

namespace {

#define CU3(expr)                                                                                                          \
    do {                                                                                                                   \
        cudaError_t _e = (expr);                                                                                           \
        if (_e != cudaSuccess)                                                                                             \
            return uapic_fail(_e == cudaErrorMemoryAllocation ? UAPIC_ENOMEM : UAPIC_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                                                 \
    } while (0)
#define TRY3(x) do { int _r = (x); if (_r) return _r; } while (0)

int make_mesh3(const uapic3d_mesh_t *mm, Mesh3 *out) {
    if (!mm) return uapic_fail(UAPIC_EINVAL, "3D mesh is null");
    for (int c = 0; c < 3; ++c) {
        if (mm->n[c] < 2 || !(mm->xmax[c] > mm->xmin[c])) return uapic_fail(UAPIC_EINVAL, "3D mesh: need n >= 2 and xmax > xmin in every direction");
        if (!poisson_size_supported(mm->n[c])) return uapic_fail(UAPIC_EUNSUPPORTED, "3D mesh size %d: powers of two <= 1024 or any n <= 512", mm->n[c]);
        out->xmin[c] = mm->xmin[c]; out->dim[c] = mm->xmax[c] - mm->xmin[c]; out->n[c] = mm->n[c];
        out->d[c] = (mm->xmax[c] - mm->xmin[c]) / (double)mm->n[c];                // meshfields.F90:103-105
    }
    return UAPIC_OK;
}

int device_sm_count(int device, int *sm) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return uapic_fail(UAPIC_ENODEVICE, "no CUDA device available; libuapic_b200 has no CPU fallback"); }
    if (device < 0 || device >= n) return uapic_fail(UAPIC_EINVAL, "device %d out of range", device);
    cudaDeviceProp prop;
    CU3(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return uapic_fail(UAPIC_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    *sm = prop.multiProcessorCount;
    CU3(cudaSetDevice(device));
    return UAPIC_OK;
}

// rho (ghosted, device) -> e (3, nx+1, ny+1, nz+1)     poisson_3d.f90:47-191 (one forward transform instead of three identical ones)
int solve3(const Mesh3 &m, int sm_count, cudaStream_t st, const double *rho, double2 *A, double2 *B, double *e, int64_t *launches) {
    const int nx = m.n[0], ny = m.n[1], nz = m.n[2];
    const int lines[3] = {ny * nz, nx * nz, nx * ny};
    auto threads = [](int len) { return len >= 512 ? 256 : (len >= 128 ? 128 : 64); };
    auto smem = [](int len) { return sizeof(double2) * 3 * (size_t)len; };
    for (int dir = 0; dir < 3; ++dir)
        if (smem(m.n[dir]) > 48 * 1024 - 256) CU3(cudaFuncSetAttribute(k3_fft_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(m.n[dir])));
    k3_fft_pass<<<lines[0], threads(nx), smem(nx), st>>>(m, 0, -1, 1, 0, rho, nullptr, A);
    k3_fft_pass<<<lines[1], threads(ny), smem(ny), st>>>(m, 1, -1, 0, 0, nullptr, nullptr, A);
    k3_fft_pass<<<lines[2], threads(nz), smem(nz), st>>>(m, 2, -1, 0, 0, nullptr, nullptr, A);
    for (int comp = 0; comp < 3; ++comp) {
        k3_fft_pass<<<lines[0], threads(nx), smem(nx), st>>>(m, 0, +1, 2, comp, nullptr, A, B);
        k3_fft_pass<<<lines[1], threads(ny), smem(ny), st>>>(m, 1, +1, 0, 0, nullptr, nullptr, B);
        k3_fft_pass<<<lines[2], threads(nz), smem(nz), st>>>(m, 2, +1, 0, 0, nullptr, nullptr, B);
        k3_store_e<<<grid3(sm_count, (int64_t)m.nodes()), k3Block, 0, st>>>(m, comp, B, e);
    }
    if (launches) *launches += 15;
    CU3(cudaGetLastError());
    return UAPIC_OK;
}

int field_update3(uapic3d_session *s) {        // compute_rho_cic -> solve_poisson -> interpolate_eb_cic
    const Mesh3 &m = s->m;
    const int64_t np = s->cfg.nbpart;
    CU3(cudaMemsetAsync(s->raw, 0, 8 * m.nodes(), s->stream));
    const double vol = s->cfg.weight / (m.d[0] * m.d[1] * m.d[2]);                   // compute_rho_cic.f90:30
    if (np > 0) k3_deposit<<<grid3(s->sm_count, np), k3Block, 0, s->stream>>>(m, np, s->x, vol, s->acc);
    if (s->nccl_comm) TRY3(uapic_internal_nccl_allreduce(s->nccl_comm, s->raw, m.nodes(), s->acc.i64 != nullptr, s->stream));
    k3_rho_finish<<<grid3(s->sm_count, (int64_t)m.nodes()), k3Block, 0, s->stream>>>(m, s->acc, s->rho);
    s->launches += 2;
    TRY3(solve3(m, s->sm_count, s->stream, s->rho, s->A, s->B, s->e, &s->launches));
    if (np > 0) k3_gather<<<grid3(s->sm_count, np), k3Block, 0, s->stream>>>(m, np, s->x, s->e, s->ep);
    s->launches += 1;
    CU3(cudaGetLastError());
    return UAPIC_OK;
}

// device scratch of the synchronous stage functions: freed on every return path
struct Scratch3 {
    void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    ~Scratch3() { for (int i = 0; i < n; ++i) if (p[i]) cudaFree(p[i]); }
    template <class T> cudaError_t alloc(T **out, size_t bytes) {
        void *q = nullptr;
        const cudaError_t e = cudaMalloc(&q, bytes ? bytes : 8);
        if (e == cudaSuccess) p[n++] = q;
        *out = reinterpret_cast<T *>(q);
        return e;
    }
};

double fixed_scale3(double total_mass_over_cell) {
    int e = 0;
    std::frexp(total_mass_over_cell > 0 ? total_mass_over_cell : 1.0, &e);
    int S = 61 - e;
    if (S > 60) S = 60;
    if (S < 8) S = 8;
    return std::ldexp(1.0, S);
}

}  // namespace

extern "C" {

int uapic3d_create(const uapic3d_config_t *cfg, uapic3d_session_t **out) {
    if (!cfg || !out) return uapic_fail(UAPIC_EINVAL, "uapic3d_create: null pointer");
    *out = nullptr;
    Mesh3 m{};
    TRY3(make_mesh3(&cfg->mesh, &m));
    if (cfg->nbpart < 0 || !(cfg->weight > 0) || !(cfg->ep > 0)) return uapic_fail(UAPIC_EINVAL, "uapic3d_create: nbpart >= 0, weight > 0, ep > 0");
    if (cfg->deposit_mode != UAPIC_DEPOSIT_FP64_ATOMIC && cfg->deposit_mode != UAPIC_DEPOSIT_FIXED_POINT) return uapic_fail(UAPIC_EINVAL, "unknown deposit_mode");
    int sm = 0;
    TRY3(device_sm_count(cfg->device, &sm));
    uapic3d_session *s = new (std::nothrow) uapic3d_session();
    if (!s) return uapic_fail(UAPIC_ENOMEM, "host allocation failed");
    s->cfg = *cfg; s->m = m; s->sm_count = sm; s->stream = (cudaStream_t)cfg->stream;
    if (!s->stream) {          // stream capture needs a real stream, not the legacy default one
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return uapic_fail(UAPIC_ECUDA, "cudaStreamCreate failed"); }
        s->own_stream = true;
    }
    const size_t np = (size_t)(cfg->nbpart ? cfg->nbpart : 1);
    cudaError_t e = cudaSuccess;
    auto al = [&](void **p, size_t n) { if (e == cudaSuccess) e = cudaMalloc(p, n); };
    al((void **)&s->x, 24 * np); al((void **)&s->v, 24 * np); al((void **)&s->ep, 24 * np);
    al((void **)&s->rho, 8 * m.nodes()); al((void **)&s->e, 24 * m.nodes()); al(&s->raw, 8 * m.nodes());
    al((void **)&s->A, 16 * m.cells()); al((void **)&s->B, 16 * m.cells());
    if (e != cudaSuccess) { delete s; cudaGetLastError(); return uapic_fail(UAPIC_ENOMEM, "device allocation failed: %s", cudaGetErrorString(e)); }
    if (cfg->deposit_mode == UAPIC_DEPOSIT_FIXED_POINT) {
        s->acc.i64 = (unsigned long long *)s->raw; s->acc.f64 = nullptr;
        s->acc.scale = fixed_scale3(cfg->weight * (double)(cfg->nbpart_global > 0 ? cfg->nbpart_global : cfg->nbpart) / (m.d[0] * m.d[1] * m.d[2]));
    } else {
        s->acc.f64 = (double *)s->raw; s->acc.i64 = nullptr; s->acc.scale = 1.0;
    }
    *out = s;
    return UAPIC_OK;
}

int uapic3d_destroy(uapic3d_session_t *s) {
    if (!s) return UAPIC_OK;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->stream);
    delete s;
    return UAPIC_OK;
}

int uapic3d_init_nccl(uapic3d_session_t *s, const void *id128, int nranks, int rank) {
    if (!s || !id128) return uapic_fail(UAPIC_EINVAL, "uapic3d_init_nccl: null pointer");
    if (nranks < 1 || rank < 0 || rank >= nranks) return uapic_fail(UAPIC_EINVAL, "uapic3d_init_nccl: rank %d of %d", rank, nranks);
    CU3(cudaSetDevice(s->cfg.device));
    if (s->nccl_comm) { uapic_internal_nccl_comm_destroy(s->nccl_comm); s->nccl_comm = nullptr; }
    for (cudaGraphExec_t &g : s->gexec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }       // captured sub-steps predate the collective
    return uapic_internal_nccl_comm_init(&s->nccl_comm, id128, nranks, rank);
}

int uapic3d_upload_particles(uapic3d_session_t *s, const double *x, const double *v) {
    if (!s || !x || !v) return uapic_fail(UAPIC_EINVAL, "uapic3d_upload_particles: null pointer");
    CU3(cudaSetDevice(s->cfg.device));
    CU3(cudaMemcpyAsync(s->x, x, 24 * (size_t)s->cfg.nbpart, cudaMemcpyHostToDevice, s->stream));
    CU3(cudaMemcpyAsync(s->v, v, 24 * (size_t)s->cfg.nbpart, cudaMemcpyHostToDevice, s->stream));
    CU3(cudaStreamSynchronize(s->stream));
    s->have_particles = true;
    return UAPIC_OK;
}

int uapic3d_generate_particles(uapic3d_session_t *s, uint64_t seed, int64_t first_global_index) {
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    CU3(cudaSetDevice(s->cfg.device));
    if (s->cfg.nbpart > 0) k3_generate<<<grid3(s->sm_count, s->cfg.nbpart), k3Block, 0, s->stream>>>(s->m, seed, first_global_index, s->cfg.nbpart, s->x, s->v);
    s->launches += 1;
    CU3(cudaGetLastError());
    s->have_particles = true;
    return UAPIC_OK;
}

int uapic3d_init_fields(uapic3d_session_t *s) {          // uapic3d.f90:76-83
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    if (!s->have_particles) return uapic_fail(UAPIC_ESTATE, "upload or generate particles before uapic3d_init_fields");
    CU3(cudaSetDevice(s->cfg.device));
    TRY3(field_update3(s));
    s->fields_ready = true;
    return UAPIC_OK;
}

// one sub-step: push(0.5 dt c) -> deposit -> Poisson -> interpolate -> rotate -> second half push
//   kind 0: uapic3d.f90:93-127 (c = 1, second half push unwrapped); kind 1: :133-165 (c = alpha); kind 2: :167-200 (c = beta)
int uapic3d_substep(uapic3d_session_t *s, int kind, double dt, double coef, int count) {
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    if (!s->fields_ready) return uapic_fail(UAPIC_ESTATE, "call uapic3d_init_fields first");
    if (kind < 0 || kind > 2 || count < 0) return uapic_fail(UAPIC_EINVAL, "uapic3d_substep: bad argument");
    CU3(cudaSetDevice(s->cfg.device));
    const int64_t np = s->cfg.nbpart;
    const double half = kind == 0 ? 0.5 * dt : 0.5 * dt * coef;
    auto body = [&]() -> int {
        if (np > 0) k3_push<<<grid3(s->sm_count, 3 * np), k3Block, 0, s->stream>>>(s->m, np, s->x, s->v, half, 1);
        TRY3(field_update3(s));
        if (np > 0) {
            k3_rotate<<<grid3(s->sm_count, np), k3Block, 0, s->stream>>>(kind, np, s->x, s->v, s->ep, dt, s->cfg.ep, coef, s->cfg.delta, s->cfg.index_quirk);
            k3_push<<<grid3(s->sm_count, np ? 3 * np : 1), k3Block, 0, s->stream>>>(s->m, np, s->x, s->v, half, kind == 0 ? 0 : 1);
        }
        s->launches += 3;
        CU3(cudaGetLastError());
        return UAPIC_OK;
    };
    const char *nog = getenv("UAPIC3D_NO_GRAPH");
    if (count >= 4 && !(nog && *nog == '1')) {
        // capture the sub-step once per (kind, dt, coef) and replay it: identical launches every time
        if (!s->gexec[kind] || s->gdt[kind] != dt || s->gcoef[kind] != coef) {
            if (s->gexec[kind]) { cudaGraphExecDestroy(s->gexec[kind]); s->gexec[kind] = nullptr; }
            const int64_t l0 = s->launches;
            cudaGraph_t g = nullptr;
            CU3(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = body();
            cudaError_t e = cudaStreamEndCapture(s->stream, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return uapic_fail(UAPIC_ECUDA, "stream capture of the 3D sub-step failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&s->gexec[kind], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return uapic_fail(UAPIC_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            s->gdt[kind] = dt; s->gcoef[kind] = coef;
            s->glaunches[kind] = s->launches - l0;
            s->launches = l0;
        }
        for (int it = 0; it < count; ++it) CU3(cudaGraphLaunch(s->gexec[kind], s->stream));
        s->launches += s->glaunches[kind] * count;
        return UAPIC_OK;
    }
    for (int it = 0; it < count; ++it) TRY3(body());
    return UAPIC_OK;
}

// the whole time loop of uapic3d.f90:44-61, :91-206 for the given Nmrc, Nmrcm, tfinal
int uapic3d_run(uapic3d_session_t *s, int nmrc, int nmrcm, double tfinal, int max_outer, int64_t *substeps) {
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    if (nmrc < 1 || nmrcm < 1 || !(tfinal > 0)) return uapic_fail(UAPIC_EINVAL, "uapic3d_run: bad argument");
    const double pi = 4.0 * std::atan(1.0), ep = s->cfg.ep;
    const long n0 = std::lround(tfinal / ep / (2.0 * pi) / (double)nmrc);             // nint, :48
    int64_t done = 0;
    if (n0 == 0 || n0 == 1) {
        const double dt = ep * (2.0 * pi) / (double)nmrc;                             // :52-53
        long nstep = std::lround(tfinal / dt);
        if (max_outer > 0 && nstep > max_outer) nstep = max_outer;
        TRY3(uapic3d_substep(s, 0, dt, 1.0, (int)nstep));
        done = nstep;
    } else {
        const double alpha = 0.5 * (1.0 + 1.0 / (double)n0) * ep * (double)n0;        // :55-57
        const double beta = 0.5 * (1.0 - 1.0 / (double)n0) * ep * (double)n0;
        const double dt = (2.0 * pi) / (double)nmrcm;
        int outer = nmrc;
        if (max_outer > 0 && outer > max_outer) outer = max_outer;
        for (int istep = 0; istep < outer; ++istep) {
            TRY3(uapic3d_substep(s, 1, dt, alpha, nmrcm));
            TRY3(uapic3d_substep(s, 2, dt, beta, nmrcm));
            done += 2 * (int64_t)nmrcm;
        }
    }
    if (substeps) *substeps = done;
    return UAPIC_OK;
}

int uapic3d_download_particles(uapic3d_session_t *s, double *x, double *v, double *ep) {
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    CU3(cudaSetDevice(s->cfg.device));
    const size_t n = 24 * (size_t)s->cfg.nbpart;
    if (x) CU3(cudaMemcpyAsync(x, s->x, n, cudaMemcpyDeviceToHost, s->stream));
    if (v) CU3(cudaMemcpyAsync(v, s->v, n, cudaMemcpyDeviceToHost, s->stream));
    if (ep) CU3(cudaMemcpyAsync(ep, s->ep, n, cudaMemcpyDeviceToHost, s->stream));
    CU3(cudaStreamSynchronize(s->stream));
    return UAPIC_OK;
}

int uapic3d_download_fields(uapic3d_session_t *s, double *e, double *rho) {
    if (!s) return uapic_fail(UAPIC_EINVAL, "session is null");
    CU3(cudaSetDevice(s->cfg.device));
    if (e) CU3(cudaMemcpyAsync(e, s->e, 24 * s->m.nodes(), cudaMemcpyDeviceToHost, s->stream));
    if (rho) CU3(cudaMemcpyAsync(rho, s->rho, 8 * s->m.nodes(), cudaMemcpyDeviceToHost, s->stream));
    CU3(cudaStreamSynchronize(s->stream));
    return UAPIC_OK;
}

int uapic3d_launch_count(uapic3d_session_t *s, int64_t *count) {
    if (!s || !count) return uapic_fail(UAPIC_EINVAL, "null pointer");
    *count = s->launches;
    return UAPIC_OK;
}

// ---- stage API of the 3D path (host buffers, synchronous) -------------------------------------------------------------
int uapic3d_compute_rho_cic(const uapic3d_mesh_t *mesh, int64_t nbpart, const double *x, double w, double *rho) {
    Mesh3 m{};
    TRY3(make_mesh3(mesh, &m));
    if (!x || !rho || nbpart < 0) return uapic_fail(UAPIC_EINVAL, "uapic3d_compute_rho_cic: bad argument");
    int sm = 0;
    TRY3(device_sm_count(0, &sm));
    double *dx = nullptr, *draw = nullptr, *drho = nullptr;
    Scratch3 sc;
    CU3(sc.alloc(&dx, 24 * (size_t)nbpart)); CU3(sc.alloc(&draw, 8 * m.nodes())); CU3(sc.alloc(&drho, 8 * m.nodes()));
    CU3(cudaMemcpy(dx, x, 24 * (size_t)nbpart, cudaMemcpyHostToDevice));
    CU3(cudaMemset(draw, 0, 8 * m.nodes()));
    RhoAcc acc{draw, nullptr, 1.0};
    if (nbpart > 0) k3_deposit<<<grid3(sm, nbpart), k3Block>>>(m, nbpart, dx, w / (m.d[0] * m.d[1] * m.d[2]), acc);
    k3_rho_finish<<<grid3(sm, (int64_t)m.nodes()), k3Block>>>(m, acc, drho);
    CU3(cudaGetLastError());
    CU3(cudaMemcpy(rho, drho, 8 * m.nodes(), cudaMemcpyDeviceToHost));
    return UAPIC_OK;
}

int uapic3d_poisson(const uapic3d_mesh_t *mesh, const double *rho, double *e) {
    Mesh3 m{};
    TRY3(make_mesh3(mesh, &m));
    if (!rho || !e) return uapic_fail(UAPIC_EINVAL, "uapic3d_poisson: null pointer");
    int sm = 0;
    TRY3(device_sm_count(0, &sm));
    double *drho = nullptr, *de = nullptr; double2 *A = nullptr, *B = nullptr;
    Scratch3 sc;
    CU3(sc.alloc(&drho, 8 * m.nodes())); CU3(sc.alloc(&de, 24 * m.nodes())); CU3(sc.alloc(&A, 16 * m.cells())); CU3(sc.alloc(&B, 16 * m.cells()));
    CU3(cudaMemcpy(drho, rho, 8 * m.nodes(), cudaMemcpyHostToDevice));
    TRY3(solve3(m, sm, 0, drho, A, B, de, nullptr));
    CU3(cudaMemcpy(e, de, 24 * m.nodes(), cudaMemcpyDeviceToHost));
    return UAPIC_OK;
}

int uapic3d_interpolate_eb_cic(const uapic3d_mesh_t *mesh, const double *e, int64_t nbpart, const double *x, double *ep) {
    Mesh3 m{};
    TRY3(make_mesh3(mesh, &m));
    if (!e || !x || !ep || nbpart < 0) return uapic_fail(UAPIC_EINVAL, "uapic3d_interpolate_eb_cic: bad argument");
    int sm = 0;
    TRY3(device_sm_count(0, &sm));
    double *dx = nullptr, *de = nullptr, *dep = nullptr;
    Scratch3 sc;
    CU3(sc.alloc(&dx, 24 * (size_t)nbpart)); CU3(sc.alloc(&de, 24 * m.nodes())); CU3(sc.alloc(&dep, 24 * (size_t)nbpart));
    CU3(cudaMemcpy(dx, x, 24 * (size_t)nbpart, cudaMemcpyHostToDevice)); CU3(cudaMemcpy(de, e, 24 * m.nodes(), cudaMemcpyHostToDevice));
    if (nbpart > 0) k3_gather<<<grid3(sm, nbpart), k3Block>>>(m, nbpart, dx, de, dep);
    CU3(cudaGetLastError());
    CU3(cudaMemcpy(ep, dep, 24 * (size_t)nbpart, cudaMemcpyDeviceToHost));
    return UAPIC_OK;
}

}  // extern "C"

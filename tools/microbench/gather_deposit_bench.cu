// gather_deposit_bench.cu -- MEASURED alternatives to the M6 gather and deposit of the one-pass kernels (north-star items 3, 4;
// VERDICT r1 rows N2, N3): the north star names "gather from shared-memory-staged E tiles, TMA-fed" and "tile-sorted
// shared-memory accumulation followed by a single flush".  Round 1 rejected both on a model; this program times them.
//
// Workload = the gather / deposit of BASELINE config 3 in isolation: Landau-like particles (x uniform, |v| Rayleigh, b(x) =
// 1 + 0.5 sin x sin y), sorted by 8 x 8-cell bin like the session does, 32 tau samples per particle on the first-order orbit of
// ua_steps.F90:78-82 (lane = sample, one particle per warp iteration, exactly the layout of k_onepass_b).
//
//   gather  G0  LDG.128 taps from the 2 x 4-tiled halo copy in global memory / L1            (the product: gather_tiled)
//           G1  LDG.256 pair loads                                                            (gather_tiled_pairs)
//           G2  per (bin, slice) work item the CTA stages the bin's E tile (+ orbit and stencil margin) in shared memory with
//               cp.async.bulk (TMA bulk copy, mbarrier completion), lanes gather with LDS.128; a sample whose stencil leaves
//               the tile falls back to G0
//   deposit D0  RED.F64 to 8 CTA-private global copies                                        (the product: deposit_split)
//           D1  shared-memory tile per (bin, slice), atomicAdd(double) on shared memory, one RED flush per touched node
//           D2  the same with int64 fixed point (atomicAdd(unsigned long long) on shared memory)
// Prints one JSON line.  Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a (tools/microbench/Makefile).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

#include "../../uapic.jl_b200/csrc/uapic_fast.cuh"

using namespace uapic;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kN = 32;                 // tau samples per particle
constexpr int kBinShift = 3;           // 8 x 8-cell bins
constexpr int kSlice = 512;            // particles per (bin, slice) work item
// shared E tile of a bin: cells [8bx - MX, 8bx + 8 + MX) need nodes 8bx - MX - 2 ... 8bx + 7 + MX + 3
constexpr int kMX = 3, kMY = 6;        // orbit margin in cells (config 3: dx = 2 dy, so twice the margin in y)
constexpr int kTW = 8 + 2 * kMX + 5, kTH = 8 + 2 * kMY + 5;   // tile nodes: 19 x 25
constexpr int kPitch = 26;             // double2 per tile row: 26 = 2 (mod 8) -> slot (i + 2j) % 8, conflict-free over any 2 x 4 box
constexpr int kBlock = 256;

struct Params {
    MeshDev m;
    MeshFast f;
    double eps;
    int64_t np;
    const double2 *x, *v;          // sorted by bin
    const int *bin_start;          // nbins + 1
    int nbx, nby;
    const double2 *ehalo_tiled;    // 2 x 4 tiled halo
    const double2 *ehalo_lin;      // linear halo (nx + 6) x (ny + 6), node (i, j) at [(i + 2) + (nx + 6) (j + 2)]
    double2 *out;                  // np * 32
    double *rho;                   // 8 copies of (nx+1)(ny+1)
    unsigned long long *rho_i;
    const int2 *work;              // (bin, first particle) per work item
    int nwork;
};

DEVINL void sample_pos(const Params &P, double2 xx, double2 vv, int n, double &px, double &py) {
    const double b = 1.0 + 0.5 * sin(xx.x) * sin(xx.y), rb = 1.0 / b;
    double s, c;
    sincospi(2.0 * (double)n / (double)kN, &s, &c);
    const double vxb = vv.x * rb, vyb = vv.y * rb;
    px = xx.x + P.eps * (s * vxb - c * vyb) + P.eps * vyb;      // ua_steps.F90:78-82
    py = xx.y + P.eps * (s * vyb + c * vxb) - P.eps * vxb;
}

template <int V> __global__ void __launch_bounds__(kBlock, 2) k_gather_global(Params P) {
    const int lane = threadIdx.x & 31, warps = kBlock / 32;
    const int64_t per = (P.np + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, P.np);
    for (int64_t p = lo + (threadIdx.x >> 5); p < hi; p += warps) {
        const double2 xx = P.x[p], vv = P.v[p];
        double px, py, xw, yw, e1, e2;
        sample_pos(P, xx, vv, lane, px, py);
        const Cell c = cell_fast(P.m, P.f, px, py, kWrapFortran, xw, yw);
        if (V == 0) gather_tiled(P.m, P.ehalo_tiled, c, e1, e2); else gather_tiled_pairs(P.m, P.ehalo_tiled, c, e1, e2);
        P.out[p * kN + lane] = make_double2(e1, e2);
    }
}

// ---- G2: TMA-staged shared-memory tile ----------------------------------------------------------------------------
DEVINL unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
DEVINL void mbar_init(unsigned long long *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
DEVINL void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEVINL void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DEVINL void mbar_wait(unsigned long long *bar, unsigned phase) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

DEVINL void gather_smem(const double2 *tile, int ti, int tj, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    const double2 *row = tile + tj * kPitch + ti;          // node (i-2, j-2) in tile coordinates
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double2 ev = row[a];
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
        row += kPitch;
    }
    e1 = s1; e2 = s2;
}

__global__ void __launch_bounds__(kBlock, 2) k_gather_tma(Params P, unsigned long long *fallbacks) {
    __shared__ __align__(128) double2 tile[kTH * kPitch];
    __shared__ __align__(8) unsigned long long bar;
    const int lane = threadIdx.x & 31, warps = kBlock / 32, ldh = P.m.nx + 6;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    unsigned phase = 0;
    unsigned long long nfall = 0;
    for (int w = blockIdx.x; w < P.nwork; w += gridDim.x) {
        const int bin = P.work[w].x, first = P.work[w].y;
        const int bx = bin % P.nbx, by = bin / P.nbx;
        // tile origin in halo coordinates (halo index = node + 2), clamped into the halo copy (orbits near the boundary that
        // wrap around fall back to the global path)
        int i0 = 8 * bx - kMX - 2 + 2, j0 = 8 * by - kMY - 2 + 2;
        i0 = min(max(i0, 0), P.m.nx + 6 - kTW); j0 = min(max(j0, 0), P.m.ny + 6 - kTH);
        __syncthreads();                      // everybody is done with the previous tile
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, kTH * kTW * 16);
            for (int r = 0; r < kTH; ++r) bulk_g2s(tile + r * kPitch, P.ehalo_lin + (size_t)(j0 + r) * ldh + i0, kTW * 16, &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        const int last = min(first + kSlice, P.bin_start[bin + 1]);
        for (int p = first + (threadIdx.x >> 5); p < last; p += warps) {
            const double2 xx = P.x[p], vv = P.v[p];
            double px, py, xw, yw, e1, e2;
            sample_pos(P, xx, vv, lane, px, py);
            const Cell c = cell_fast(P.m, P.f, px, py, kWrapFortran, xw, yw);
            const int ti = c.i - i0, tj = c.j - j0;      // halo index of node (i-2) is c.i
            if (ti >= 0 && ti + 6 <= kTW && tj >= 0 && tj + 6 <= kTH) gather_smem(tile, ti, tj, c, e1, e2);
            else { gather_tiled(P.m, P.ehalo_tiled, c, e1, e2); ++nfall; }
            P.out[(int64_t)p * kN + lane] = make_double2(e1, e2);
        }
    }
    if (nfall) atomicAdd(fallbacks, nfall);
}

// ---- deposits -----------------------------------------------------------------------------------------------------
// as in deposit_split<4>: a particle is spread over 4 lanes, 3 x 3 taps each
DEVINL void half_weights(double u, double (&w)[3]) {
    const double a = pow5(u), b = pow5(1.0 + u), c = pow5(2.0 + u), k = 1.0 / 120.0;
    w[0] = a * k; w[1] = fma(-6.0, a, b) * k; w[2] = fma(15.0, a, fma(-6.0, b, c)) * k;
}
DEVINL Cell deposit_cell(const Params &P, double2 xx, double2 vv, double dt) {
    double xw, yw;
    return cell_fast(P.m, P.f, xx.x + dt * vv.x, xx.y + dt * vv.y, kWrapFortran, xw, yw);
}

template <int V> __global__ void __launch_bounds__(kBlock, 2) k_deposit(Params P, double dt, double weight, double scale, unsigned long long *fallbacks) {
    // V = 0: global RED to CTA-private copies; 1: shared tile, fp64; 2: shared tile, int64 fixed point
    constexpr int kDW = 8 + 2 * 8 + 5, kDH = 8 + 2 * 16 + 5;      // deposit tile: the particle has moved by dt*v (2|v| / 4|v| cells)
    __shared__ double tile[V == 0 ? 1 : kDW * kDH];
    unsigned long long *itile = reinterpret_cast<unsigned long long *>(tile);
    const int lane = threadIdx.x & 31, g = lane & 3, pin = lane >> 2;
    const size_t nrho = (size_t)P.m.ld * (P.m.ny + 1);
    unsigned long long nfall = 0;
    if (V == 0) {
        double *rho = P.rho + (blockIdx.x & 7) * nrho;
        const int64_t per = (P.np + gridDim.x - 1) / gridDim.x;
        const int64_t lo = blockIdx.x * per, hi = min(lo + per, P.np);
        for (int64_t p0 = lo + (threadIdx.x >> 5) * 8; p0 < hi; p0 += (kBlock / 32) * 8) {
            const int64_t p = min(p0 + pin, hi - 1);
            const bool valid = p0 + pin < hi;
            const Cell c = deposit_cell(P, P.x[p], P.v[p], dt);
            double wx[3], wy[3];
            const bool hx = g & 1, hy = g & 2;
            half_weights(hx ? c.dpx : 1.0 - c.dpx, wx);
            half_weights(hy ? c.dpy : 1.0 - c.dpy, wy);
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int jy = wrap_fast(c.j, hy ? 3 - j : j - 2, P.m.ny) * P.m.ld;
#pragma unroll
                for (int i = 0; i < 3; ++i) atomicAdd(rho + wrap_fast(c.i, hx ? 3 - i : i - 2, P.m.nx) + jy, wx[i] * wy[j] * weight);
            }
        }
        return;
    }
    for (int w = blockIdx.x; w < P.nwork; w += gridDim.x) {
        const int bin = P.work[w].x, first = P.work[w].y;
        const int bx = bin % P.nbx, by = bin / P.nbx;
        const int i0 = 8 * bx - 8 - 2, j0 = 8 * by - 16 - 2;          // node index of tile (0, 0); may be negative (wrapped at the flush)
        __syncthreads();
        for (int q = threadIdx.x; q < kDW * kDH; q += kBlock) { if (V == 1) tile[q] = 0.0; else itile[q] = 0ull; }
        __syncthreads();
        const int last = min(first + kSlice, P.bin_start[bin + 1]);
        for (int p0 = first + (threadIdx.x >> 5) * 8; p0 < last; p0 += (kBlock / 32) * 8) {
            const int p = min(p0 + pin, last - 1);
            const bool valid = p0 + pin < last;
            const Cell c = deposit_cell(P, P.x[p], P.v[p], dt);
            double wx[3], wy[3];
            const bool hx = g & 1, hy = g & 2;
            half_weights(hx ? c.dpx : 1.0 - c.dpx, wx);
            half_weights(hy ? c.dpy : 1.0 - c.dpy, wy);
            if (!valid) continue;
            // tile coordinates of the lane's 3 x 3 block (unwrapped node indices; a particle whose block leaves the tile, or
            // whose cell wrapped around the period, goes to global memory)
            int ci = c.i, cj = c.j;
            if (ci - i0 > P.m.nx / 2) ci -= P.m.nx; else if (i0 - ci > P.m.nx / 2) ci += P.m.nx;
            if (cj - j0 > P.m.ny / 2) cj -= P.m.ny; else if (j0 - cj > P.m.ny / 2) cj += P.m.ny;
            const int ti = ci + (hx ? 1 : -2) - i0, tj = cj + (hy ? 1 : -2) - j0;
            const bool inside = ti >= 0 && ti + 3 <= kDW && tj >= 0 && tj + 3 <= kDH;
            if (!inside) ++nfall;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double val = wx[hx ? 2 - i : i] * wy[hy ? 2 - j : j] * weight;      // ascending node order inside the block
                    if (inside) {
                        if (V == 1) atomicAdd(&tile[(tj + j) * kDW + ti + i], val);
                        else atomicAdd(&itile[(tj + j) * kDW + ti + i], (unsigned long long)__double2ll_rn(val * scale));
                    } else {
                        const int gi = wrap_fast(c.i, (hx ? 1 : -2) + i, P.m.nx), gj = wrap_fast(c.j, (hy ? 1 : -2) + j, P.m.ny);
                        if (V == 1) atomicAdd(P.rho + gi + gj * P.m.ld, val);
                        else atomicAdd(P.rho_i + gi + gj * P.m.ld, (unsigned long long)__double2ll_rn(val * scale));
                    }
                }
            }
        }
        __syncthreads();
        // single flush of the tile (nodes wrap periodically)
        for (int q = threadIdx.x; q < kDW * kDH; q += kBlock) {
            const int tj = q / kDW, ti = q - tj * kDW;
            int gi = (i0 + ti) % P.m.nx, gj = (j0 + tj) % P.m.ny;
            gi += gi < 0 ? P.m.nx : 0; gj += gj < 0 ? P.m.ny : 0;
            if (V == 1) { const double t = tile[q]; if (t != 0.0) atomicAdd(P.rho + gi + gj * P.m.ld, t); }
            else { const unsigned long long t = itile[q]; if (t) atomicAdd(P.rho_i + gi + gj * P.m.ld, t); }
        }
    }
    if (nfall) atomicAdd(fallbacks, nfall);
}

template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv) {
    const int64_t np = argc > 1 ? atoll(argv[1]) : 2000000;
    const int nx = argc > 2 ? atoi(argv[2]) : 128, ny = argc > 3 ? atoi(argv[3]) : 128;
    const double eps = argc > 4 ? atof(argv[4]) : 0.1;
    const double pi = 3.14159265358979323846, dimx = 4 * pi, dimy = 2 * pi, dt = pi / 16;
    Params P{};
    P.m.xmin = 0; P.m.ymin = 0; P.m.dimx = dimx; P.m.dimy = dimy; P.m.dx = dimx / nx; P.m.dy = dimy / ny; P.m.nx = nx; P.m.ny = ny; P.m.ld = nx + 1;
    P.f.inv_dx = 1 / P.m.dx; P.f.inv_dy = 1 / P.m.dy; P.f.inv_nx = 1.0 / nx; P.f.inv_ny = 1.0 / ny; P.f.inv_dimx = 1 / dimx; P.f.inv_dimy = 1 / dimy;
    P.eps = eps; P.np = np;
    P.nbx = (nx + 7) >> kBinShift; P.nby = (ny + 7) >> kBinShift;
    const int nbins = P.nbx * P.nby;
    // ---- particles, sorted by bin ----
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double2> x(np), v(np);
    std::vector<int> bin(np);
    for (int64_t k = 0; k < np; ++k) {
        x[k] = make_double2(U(rng) * dimx, U(rng) * dimy);
        const double vr = std::sqrt(-2.0 * std::log(1.0 - U(rng))), th = 2 * pi * U(rng);
        v[k] = make_double2(vr * std::cos(th), vr * std::sin(th));
        const int i = std::min(nx - 1, (int)(x[k].x / P.m.dx)), j = std::min(ny - 1, (int)(x[k].y / P.m.dy));
        bin[k] = (j >> kBinShift) * P.nbx + (i >> kBinShift);
    }
    std::vector<int64_t> order(np);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return bin[a] < bin[b]; });
    std::vector<double2> xs(np), vs(np);
    std::vector<int> bin_start(nbins + 1, 0);
    for (int64_t k = 0; k < np; ++k) { xs[k] = x[order[k]]; vs[k] = v[order[k]]; bin_start[bin[order[k]] + 1]++; }
    for (int b = 0; b < nbins; ++b) bin_start[b + 1] += bin_start[b];
    std::vector<int2> work;
    for (int b = 0; b < nbins; ++b)
        for (int p = bin_start[b]; p < bin_start[b + 1]; p += kSlice) work.push_back(make_int2(b, p));
    // ---- field ----
    std::vector<double2> emesh((size_t)(nx + 1) * (ny + 1));
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) emesh[i + (size_t)(nx + 1) * j] = make_double2(U(rng) - 0.5, U(rng) - 0.5);
    const int ntx = (nx + 7) >> 1, nty = (ny + 9) >> 2;
    std::vector<double2> tiled((size_t)ntx * nty * 8), lin((size_t)(nx + 6) * (ny + 6));
    auto node = [&](int i, int j) { i %= nx; if (i < 0) i += nx; j %= ny; if (j < 0) j += ny; return emesh[i + (size_t)(nx + 1) * j]; };
    for (int J = 0; J < 4 * nty; ++J) for (int I = 0; I < 2 * ntx; ++I) tiled[(((J >> 2) * ntx + (I >> 1)) << 3) + ((J & 3) << 1) + (I & 1)] = node(I - 2, J - 2);
    for (int J = 0; J < ny + 6; ++J) for (int I = 0; I < nx + 6; ++I) lin[I + (size_t)(nx + 6) * J] = node(I - 2, J - 2);
    // ---- device ----
    double2 *dx, *dv, *dtiled, *dlin, *dout, *dout2;
    int *dbs; int2 *dwork; double *drho; unsigned long long *drhoi, *dfall;
    const size_t nrho = (size_t)(nx + 1) * (ny + 1);
    CK(cudaMalloc(&dx, 16 * np)); CK(cudaMalloc(&dv, 16 * np)); CK(cudaMalloc(&dout, 16 * np * kN)); CK(cudaMalloc(&dout2, 16 * np * kN));
    CK(cudaMalloc(&dtiled, 16 * tiled.size())); CK(cudaMalloc(&dlin, 16 * lin.size()));
    CK(cudaMalloc(&dbs, 4 * (nbins + 1))); CK(cudaMalloc(&dwork, 8 * work.size()));
    CK(cudaMalloc(&drho, 8 * 8 * nrho)); CK(cudaMalloc(&drhoi, 8 * nrho)); CK(cudaMalloc(&dfall, 8));
    CK(cudaMemcpy(dx, xs.data(), 16 * np, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dv, vs.data(), 16 * np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtiled, tiled.data(), 16 * tiled.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dlin, lin.data(), 16 * lin.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbs, bin_start.data(), 4 * (nbins + 1), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dwork, work.data(), 8 * work.size(), cudaMemcpyHostToDevice));
    P.x = dx; P.v = dv; P.bin_start = dbs; P.ehalo_tiled = dtiled; P.ehalo_lin = dlin; P.out = dout; P.rho = drho; P.rho_i = drhoi; P.work = dwork; P.nwork = (int)work.size();
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = 2 * sms, reps = 10;
    // ---- gathers ----
    const float g0 = time_ms([&] { k_gather_global<0><<<grid, kBlock>>>(P); }, reps);
    Params P2 = P; P2.out = dout2;
    const float g1 = time_ms([&] { k_gather_global<1><<<grid, kBlock>>>(P2); }, reps);
    std::vector<double2> r0((size_t)np * kN), r1((size_t)np * kN);
    CK(cudaMemcpy(r0.data(), dout, 16 * np * kN, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r1.data(), dout2, 16 * np * kN, cudaMemcpyDeviceToHost));
    const bool same01 = memcmp(r0.data(), r1.data(), 16 * (size_t)np * kN) == 0;
    CK(cudaMemset(dout2, 0, 16 * np * kN)); CK(cudaMemset(dfall, 0, 8));
    const float g2 = time_ms([&] { k_gather_tma<<<grid, kBlock>>>(P2, dfall); }, reps);
    unsigned long long fall_g = 0;
    CK(cudaMemcpy(&fall_g, dfall, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r1.data(), dout2, 16 * np * kN, cudaMemcpyDeviceToHost));
    const bool same02 = memcmp(r0.data(), r1.data(), 16 * (size_t)np * kN) == 0;
    // ---- deposits ----
    const double weight = dimx * dimy / (double)np, scale = std::ldexp(1.0, 54);
    CK(cudaMemset(drho, 0, 8 * 8 * nrho));
    const float d0 = time_ms([&] { k_deposit<0><<<grid, kBlock>>>(P, dt, weight, scale, dfall); }, reps);
    std::vector<double> h0(8 * nrho), h1(nrho);
    CK(cudaMemcpy(h0.data(), drho, 8 * 8 * nrho, cudaMemcpyDeviceToHost));
    double tot0 = 0; for (double t : h0) tot0 += t;
    CK(cudaMemset(drho, 0, 8 * nrho)); CK(cudaMemset(dfall, 0, 8));
    const float d1 = time_ms([&] { k_deposit<1><<<grid, kBlock>>>(P, dt, weight, scale, dfall); }, reps);
    unsigned long long fall_d = 0;
    CK(cudaMemcpy(&fall_d, dfall, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1.data(), drho, 8 * nrho, cudaMemcpyDeviceToHost));
    double tot1 = 0, maxd = 0;
    for (size_t q = 0; q < nrho; ++q) { tot1 += h1[q]; double s = 0; for (int c = 0; c < 8; ++c) s += h0[c * nrho + q]; maxd = std::max(maxd, std::fabs(s - h1[q])); }
    CK(cudaMemset(drhoi, 0, 8 * nrho));
    const float d2 = time_ms([&] { k_deposit<2><<<grid, kBlock>>>(P, dt, weight, scale, dfall); }, reps);
    const double samples = (double)np * kN;
    printf("{\"particles\": %lld, \"mesh\": [%d, %d], \"eps\": %g, \"work_items\": %d, "
           "\"gather_ms\": {\"G0_ldg128_tiled_global\": %.4f, \"G1_ldg256_pairs\": %.4f, \"G2_tma_smem_tile\": %.4f}, "
           "\"gather_samples_per_s\": {\"G0\": %.4e, \"G1\": %.4e, \"G2\": %.4e}, \"G1_bit_identical_to_G0\": %s, \"G2_bit_identical_to_G0\": %s, "
           "\"G2_fallback_share\": %.5f, \"G2_tile_nodes\": [%d, %d], "
           "\"deposit_ms\": {\"D0_red_global_8_copies\": %.4f, \"D1_smem_tile_fp64\": %.4f, \"D2_smem_tile_int64\": %.4f}, "
           "\"D1_fallback_share\": %.5f, \"deposit_total_rel_diff\": %.3e, \"deposit_max_abs_diff_per_launch\": %.3e}\n",
           (long long)np, nx, ny, eps, (int)work.size(), g0, g1, g2, samples / (g0 * 1e-3), samples / (g1 * 1e-3), samples / (g2 * 1e-3),
           same01 ? "true" : "false", same02 ? "true" : "false", (double)fall_g / ((reps + 1) * samples), kTW, kTH,
           d0, d1, d2, (double)fall_d / ((reps + 1) * 4.0 * np), std::fabs(tot0 - tot1) / std::fabs(tot0), maxd / (reps + 1));
    return 0;
}

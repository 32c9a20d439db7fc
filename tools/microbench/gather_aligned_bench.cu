// gather_aligned_bench.cu -- can the M6 gather reach the 4-wavefront floor by giving every QUARTER-WARP a copy of the E halo whose
// 2 x 4 tile grid is aligned to the bounding box of its 8 samples?  (8 copies, one per alignment (ox, oy) in 2 x 4; the copy is
// chosen per tap: alignment of (i_min + a, j_min + b).)  If the 8 cells of a quarter-warp fit a 2 x 4 box they then share ONE line:
// 1 wavefront per quarter-warp, 4 per instruction.  Cost: two 3-stage min-reductions (6 + 6 SHFL) per sample gather, 8x the L1
// footprint, more index arithmetic.  Variants: 0 = the product's gather_tiled (one copy), 1 = aligned copies.
// Same workload as gather_layout_bench.cu.  Prints ms; run under ncu for l1tex__data_pipe_lsu_wavefronts.
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
constexpr int kN = 32, kBlock = 256;

struct Params {
    MeshDev m; MeshFast f; double eps; int64_t np;
    const double2 *x, *v, *e1;     // e1: the product's tiled halo
    const double2 *e8;             // 8 aligned copies, copy c at e8 + c * stride8
    size_t stride8; int ntx8;      // tiles per tile-row of the padded copies
    double2 *out;
};
DEVINL void sample_pos(const Params &P, double2 xx, double2 vv, int n, double &px, double &py) {
    const double b = 1.0 + 0.5 * sin(xx.x) * sin(xx.y), rb = 1.0 / b;
    double s, c;
    sincospi(2.0 * (double)n / (double)kN, &s, &c);
    const double vxb = vv.x * rb, vyb = vv.y * rb;
    px = xx.x + P.eps * (s * vxb - c * vyb) + P.eps * vyb;
    py = xx.y + P.eps * (s * vyb + c * vxb) - P.eps * vxb;
}
// node (I, J) of the halo (I = i + 2) in the copy whose tile grid starts at (ox, oy): padded coordinates I + 2 - ox, J + 4 - oy
DEVINL int idx8(int ntx, int I, int J, int ox, int oy) {
    const int Ip = I + 2 - ox, Jp = J + 4 - oy;
    return (((Jp >> 2) * ntx + (Ip >> 1)) << 3) + ((Jp & 3) << 1) + (Ip & 1);
}
DEVINL void gather_aligned(const Params &P, const Cell &c, int imin, int jmin, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const int oy = (jmin + b) & 3;
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int ox = (imin + a) & 1;
            const double2 ev = __ldg(P.e8 + (size_t)(ox + 2 * oy) * P.stride8 + idx8(P.ntx8, c.i + a, c.j + b, ox, oy));
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
    }
    e1 = s1; e2 = s2;
}
template <int V> __global__ void __launch_bounds__(kBlock, 2) k_gather(Params P) {
    const int lane = threadIdx.x & 31, warps = kBlock / 32;
    const int64_t per = (P.np + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, P.np);
    for (int64_t p = lo + (threadIdx.x >> 5); p < hi; p += warps) {
        const double2 xx = P.x[p], vv = P.v[p];
        double px, py, xw, yw, e1, e2;
        sample_pos(P, xx, vv, lane, px, py);
        const Cell c = cell_fast(P.m, P.f, px, py, kWrapFortran, xw, yw);
        if (V == 0) {
            gather_tiled(P.m, P.e1, c, e1, e2);
        } else {
            int imin = c.i, jmin = c.j;
#pragma unroll
            for (int h = 1; h < 8; h <<= 1) { imin = min(imin, __shfl_xor_sync(kFull, imin, h)); jmin = min(jmin, __shfl_xor_sync(kFull, jmin, h)); }
            // cells that wrapped around the period sit far from the minimum: any alignment is as good as another for them
            gather_aligned(P, c, imin, jmin, e1, e2);
        }
        P.out[p * kN + lane] = make_double2(e1, e2);
    }
}
template <class F> float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int r = 0; r < reps; ++r) f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}
int main(int argc, char **argv) {
    const int64_t np = argc > 1 ? atoll(argv[1]) : 2000000;
    const int nx = argc > 2 ? atoi(argv[2]) : 128, ny = argc > 3 ? atoi(argv[3]) : 128;
    const double eps = argc > 4 ? atof(argv[4]) : 0.1;
    const double pi = 3.14159265358979323846, dimx = 4 * pi, dimy = 2 * pi;
    Params P{};
    P.m.xmin = 0; P.m.ymin = 0; P.m.dimx = dimx; P.m.dimy = dimy; P.m.dx = dimx / nx; P.m.dy = dimy / ny; P.m.nx = nx; P.m.ny = ny; P.m.ld = nx + 1;
    P.f.inv_dx = 1 / P.m.dx; P.f.inv_dy = 1 / P.m.dy; P.f.inv_nx = 1.0 / nx; P.f.inv_ny = 1.0 / ny; P.f.inv_dimx = 1 / dimx; P.f.inv_dimy = 1 / dimy;
    P.eps = eps; P.np = np;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double2> x(np), v(np);
    std::vector<int> bin(np);
    const int nbx = (nx + 7) >> 3;
    for (int64_t k = 0; k < np; ++k) {
        x[k] = make_double2(U(rng) * dimx, U(rng) * dimy);
        const double vr = std::sqrt(-2.0 * std::log(1.0 - U(rng))), th = 2 * pi * U(rng);
        v[k] = make_double2(vr * std::cos(th), vr * std::sin(th));
        bin[k] = (std::min(ny - 1, (int)(x[k].y / P.m.dy)) >> 3) * nbx + (std::min(nx - 1, (int)(x[k].x / P.m.dx)) >> 3);
    }
    std::vector<int64_t> order(np);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return bin[a] < bin[b]; });
    std::vector<double2> xs(np), vs(np);
    for (int64_t k = 0; k < np; ++k) { xs[k] = x[order[k]]; vs[k] = v[order[k]]; }
    std::vector<double2> emesh((size_t)nx * ny);
    for (auto &e : emesh) e = make_double2(U(rng) - 0.5, U(rng) - 0.5);
    auto node = [&](int i, int j) { i %= nx; if (i < 0) i += nx; j %= ny; if (j < 0) j += ny; return emesh[i + (size_t)nx * j]; };
    // the product's tiled halo
    const int ntx = (nx + 7) >> 1, nty = (ny + 9) >> 2;
    std::vector<double2> tiled((size_t)ntx * nty * 8);
    for (int J = 0; J < 4 * nty; ++J) for (int I = 0; I < 2 * ntx; ++I) tiled[(((J >> 2) * ntx + (I >> 1)) << 3) + ((J & 3) << 1) + (I & 1)] = node(I - 2, J - 2);
    // 8 aligned copies (padded by one tile each way)
    const int ntx8 = (nx + 6 + 2 + 1) / 2 + 1, nty8 = (ny + 6 + 4 + 3) / 4 + 1;
    const size_t stride8 = (size_t)ntx8 * nty8 * 8;
    std::vector<double2> c8(8 * stride8, make_double2(0, 0));
    for (int oy = 0; oy < 4; ++oy) for (int ox = 0; ox < 2; ++ox)
        for (int J = 0; J < ny + 6; ++J) for (int I = 0; I < nx + 6; ++I) {
            const int Ip = I + 2 - ox, Jp = J + 4 - oy;
            c8[(size_t)(ox + 2 * oy) * stride8 + ((((Jp >> 2) * ntx8 + (Ip >> 1)) << 3) + ((Jp & 3) << 1) + (Ip & 1))] = node(I - 2, J - 2);
        }
    double2 *dx, *dv, *dout, *d1, *d8;
    CK(cudaMalloc(&dx, 16 * np)); CK(cudaMalloc(&dv, 16 * np)); CK(cudaMalloc(&dout, 16 * np * kN));
    CK(cudaMalloc(&d1, 16 * tiled.size())); CK(cudaMalloc(&d8, 16 * c8.size()));
    CK(cudaMemcpy(dx, xs.data(), 16 * np, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dv, vs.data(), 16 * np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d1, tiled.data(), 16 * tiled.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d8, c8.data(), 16 * c8.size(), cudaMemcpyHostToDevice));
    P.x = dx; P.v = dv; P.out = dout; P.e1 = d1; P.e8 = d8; P.stride8 = stride8; P.ntx8 = ntx8;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<double2> r0((size_t)np * kN), r1((size_t)np * kN);
    const float t0 = time_ms([&] { k_gather<0><<<2 * sms, kBlock>>>(P); }, 5);
    CK(cudaMemcpy(r0.data(), dout, 16 * np * kN, cudaMemcpyDeviceToHost));
    const float t1 = time_ms([&] { k_gather<1><<<2 * sms, kBlock>>>(P); }, 5);
    CK(cudaMemcpy(r1.data(), dout, 16 * np * kN, cudaMemcpyDeviceToHost));
    const bool same = memcmp(r0.data(), r1.data(), 16 * (size_t)np * kN) == 0;
    printf("{\"particles\": %lld, \"mesh\": [%d, %d], \"eps\": %g, \"ms\": {\"one_copy_2x4\": %.4f, \"aligned_8_copies\": %.4f}, \"bit_identical\": %s}\n",
           (long long)np, nx, ny, eps, t0, t1, same ? "true" : "false");
    return 0;
}

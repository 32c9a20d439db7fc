// gather_layout_bench.cu -- which memory layout of the E halo costs the fewest L1 data-pipe wavefronts per M6 tap?
// Same workload as gather_deposit_bench.cu (config-3-shaped orbits, particles sorted by 8 x 8-cell bin, lane = tau sample).
// Layouts: tiles of TX x TY nodes per 128-byte line (2x4 = the product, 4x2, 8x1, 1x8), row-major arrays whose row pitch is
// p (mod 8) nodes (bank slot = (i + p j) mod 8), and a PER-QUARTER-WARP choice between the 2x4 and the 4x2 copy (the 8 lanes
// of a quarter-warp vote on the span of their cells).  Run under ncu for l1tex__data_pipe_lsu_wavefronts; prints ms per layout.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../../uapic.jl_b200/csrc/uapic_fast.cuh"
using namespace uapic;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int kN = 32, kBlock = 256;

struct Lay { int kind, tx, ty, ntx, pitch; };     // kind 0: tiles tx x ty (tx*ty = 8), ntx tiles per row of tiles ; kind 1: row-major, pitch nodes
__host__ __device__ inline int lay_index(const Lay &L, int I, int J) {
    if (L.kind == 0) return ((J / L.ty) * L.ntx + I / L.tx) * 8 + (J % L.ty) * L.tx + (I % L.tx);
    return J * L.pitch + I;
}
struct Params {
    MeshDev m; MeshFast f; double eps; int64_t np;
    const double2 *x, *v; const double2 *e0, *e1; Lay l0, l1; int mode;   // mode 0: layout l0 only; 1: per-quarter choice between l0 and l1
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
DEVINL void gather_lay(const Lay &L, const double2 *__restrict__ e, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double2 ev = __ldg(e + lay_index(L, c.i + a, c.j + b));
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
    }
    e1 = s1; e2 = s2;
}
__global__ void __launch_bounds__(kBlock, 2) k_gather(Params P) {
    const int lane = threadIdx.x & 31, warps = kBlock / 32;
    const int64_t per = (P.np + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, P.np);
    for (int64_t p = lo + (threadIdx.x >> 5); p < hi; p += warps) {
        const double2 xx = P.x[p], vv = P.v[p];
        double px, py, xw, yw, e1, e2;
        sample_pos(P, xx, vv, lane, px, py);
        const Cell c = cell_fast(P.m, P.f, px, py, kWrapFortran, xw, yw);
        bool second = false;
        if (P.mode == 1) {
            // span of the quarter-warp's cells (8 lanes): l0 = 2x4 tiles is conflict-free for spans <= 2 x 4, l1 = 4x2 for <= 4 x 2
            int imin = c.i, imax = c.i, jmin = c.j, jmax = c.j;
#pragma unroll
            for (int h = 1; h < 8; h <<= 1) {
                imin = min(imin, __shfl_xor_sync(kFull, imin, h)); imax = max(imax, __shfl_xor_sync(kFull, imax, h));
                jmin = min(jmin, __shfl_xor_sync(kFull, jmin, h)); jmax = max(jmax, __shfl_xor_sync(kFull, jmax, h));
            }
            second = (imax - imin >= 2) && (jmax - jmin <= 1);
        }
        if (second) gather_lay(P.l1, P.e1, c, e1, e2); else gather_lay(P.l0, P.e0, c, e1, e2);
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
    const int W = nx + 8, H = ny + 8;      // halo extent rounded up to whole tiles of every shape
    auto build = [&](const Lay &L, double2 **dptr) {
        size_t n = 0;
        for (int J = 0; J < H; ++J) for (int I = 0; I < W; ++I) n = std::max(n, (size_t)lay_index(L, I, J) + 1);
        std::vector<double2> h(n, make_double2(0, 0));
        for (int J = 0; J < H; ++J) for (int I = 0; I < W; ++I) h[lay_index(L, I, J)] = node(I - 2, J - 2);
        CK(cudaMalloc(dptr, 16 * n)); CK(cudaMemcpy(*dptr, h.data(), 16 * n, cudaMemcpyHostToDevice));
    };
    auto tiles = [&](int tx, int ty) { Lay L{0, tx, ty, (W + tx - 1) / tx, 0}; return L; };
    auto rows = [&](int p) { int pitch = W; while (pitch % 8 != p) ++pitch; Lay L{1, 0, 0, 0, pitch}; return L; };
    struct Case { const char *name; Lay l0, l1; int mode; };
    std::vector<Case> cases = {
        {"tile_2x4", tiles(2, 4), tiles(2, 4), 0}, {"tile_4x2", tiles(4, 2), tiles(4, 2), 0}, {"tile_8x1", tiles(8, 1), tiles(8, 1), 0},
        {"tile_1x8", tiles(1, 8), tiles(1, 8), 0}, {"rows_pitch1", rows(1), rows(1), 0}, {"rows_pitch2", rows(2), rows(2), 0},
        {"rows_pitch3", rows(3), rows(3), 0}, {"rows_pitch5", rows(5), rows(5), 0}, {"rows_pitch6", rows(6), rows(6), 0},
        {"quarter_choice_2x4_or_4x2", tiles(2, 4), tiles(4, 2), 1},
    };
    double2 *dx, *dv, *dout;
    CK(cudaMalloc(&dx, 16 * np)); CK(cudaMalloc(&dv, 16 * np)); CK(cudaMalloc(&dout, 16 * np * kN));
    CK(cudaMemcpy(dx, xs.data(), 16 * np, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dv, vs.data(), 16 * np, cudaMemcpyHostToDevice));
    P.x = dx; P.v = dv; P.out = dout;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<double2> ref, got((size_t)np * kN);
    printf("{\"particles\": %lld, \"mesh\": [%d, %d], \"eps\": %g, \"ms\": {", (long long)np, nx, ny, eps);
    for (size_t q = 0; q < cases.size(); ++q) {
        double2 *d0, *d1;
        build(cases[q].l0, &d0); build(cases[q].l1, &d1);
        Params Q = P; Q.e0 = d0; Q.e1 = d1; Q.l0 = cases[q].l0; Q.l1 = cases[q].l1; Q.mode = cases[q].mode;
        const float ms = time_ms([&] { k_gather<<<2 * sms, kBlock>>>(Q); }, 3);
        CK(cudaMemcpy(got.data(), dout, 16 * np * kN, cudaMemcpyDeviceToHost));
        if (q == 0) ref = got;
        bool same = true; for (size_t k = 0; k < got.size() && same; ++k) same = got[k].x == ref[k].x && got[k].y == ref[k].y;
        printf("%s\"%s\": %.4f", q ? ", " : "", cases[q].name, ms);
        if (!same) fprintf(stderr, "layout %s differs from tile_2x4!\n", cases[q].name);
        CK(cudaFree(d0)); CK(cudaFree(d1));
    }
    printf("}}\n");
    return 0;
}

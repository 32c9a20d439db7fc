// uapic_sort.cu -- spatial reordering of the particle arrays (counting sort by coarse mesh bin), sm_100a.
//
// Not part of the reference's algorithm (particles are independent; every sum over particles is either an atomic
// accumulation or order independent in fixed-point mode): it only makes consecutive particles -- the ones a warp, a CTA
// and an SM work on at the same time -- neighbours on the mesh, so that the 36-tap M6 gathers of the phase kernels hit
// in L1 (measured: +20 % on the gather-bound kernel).  The permutation is tracked (`perm[slot]` = index the particle had
// when it was uploaded or generated) and undone by the download entry points, so callers never see it.
//
//   k_sort_count   : bin of every particle (uint16) + global histogram (shared-memory pre-aggregation per CTA)
//   k_sort_scan    : exclusive scan of the histogram (one CTA)
//   k_sort_scatter : every CTA reserves a range per bin with ONE atomic per bin, then moves x, v, e, perm
//   k_unpermute2   : out[perm[i]] = in[i] for the downloads
#include "uapic_internal.h"
#include "uapic_fast.cuh"

namespace uapic {

namespace {

constexpr int kSortBlock = 256;
constexpr int kSortChunk = 4096;     // particles per CTA

struct SortGeom {
    double inv_dx, inv_dy, inv_nx, inv_ny;
    double fnx, fny;
    int nx, ny, shift, nbx, nbins;
};

DEVINL int bin_of(const SortGeom &g, double2 p) {
    const double px = modulo_fast(p.x * g.inv_dx, g.fnx, g.inv_nx), py = modulo_fast(p.y * g.inv_dy, g.fny, g.inv_ny);
    int i = __double2int_rd(px), j = __double2int_rd(py);
    i = min(max(i, 0), g.nx - 1); j = min(max(j, 0), g.ny - 1);
    return (j >> g.shift) * g.nbx + (i >> g.shift);
}

__global__ void __launch_bounds__(kSortBlock) k_sort_count(SortGeom g, int64_t np, const double2 *__restrict__ x,
                                                           uint16_t *__restrict__ binid, unsigned *__restrict__ hist) {
    extern __shared__ unsigned cnt[];
    for (int b = threadIdx.x; b < g.nbins; b += kSortBlock) cnt[b] = 0;
    __syncthreads();
    const int64_t first = (int64_t)blockIdx.x * kSortChunk;
    for (int k = threadIdx.x; k < kSortChunk; k += kSortBlock) {
        const int64_t i = first + k;
        if (i < np) {
            const int b = bin_of(g, x[i]);
            binid[i] = (uint16_t)b;
            atomicAdd(&cnt[b], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < g.nbins; b += kSortBlock)
        if (cnt[b]) atomicAdd(&hist[b], cnt[b]);
}

// hist[0..nbins) -> exclusive prefix sums in place; nbins <= 4096; one CTA of 1024 threads, 4 bins per thread
__global__ void __launch_bounds__(1024) k_sort_scan(int nbins, unsigned *hist) {
    __shared__ unsigned part[1024];
    const int t = threadIdx.x;
    unsigned v[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int b = 4 * t + q; v[q] = b < nbins ? hist[b] : 0u; s += v[q]; }
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const unsigned add = t >= off ? part[t - off] : 0u;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    unsigned run = part[t] - s;
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int b = 4 * t + q; if (b < nbins) hist[b] = run; run += v[q]; }
}

__global__ void __launch_bounds__(kSortBlock) k_sort_scatter(int nbins, int64_t np, const uint16_t *__restrict__ binid,
                                                             unsigned *__restrict__ cursor, const double2 *__restrict__ x,
                                                             const double2 *__restrict__ v, const double2 *__restrict__ ep,
                                                             const uint32_t *__restrict__ perm, double2 *__restrict__ x2,
                                                             double2 *__restrict__ v2, double2 *__restrict__ ep2,
                                                             uint32_t *__restrict__ perm2, uint32_t index_base) {
    extern __shared__ unsigned sm[];
    unsigned *cnt = sm, *base = sm + nbins;
    for (int b = threadIdx.x; b < nbins; b += kSortBlock) cnt[b] = 0;
    __syncthreads();
    const int64_t first = (int64_t)blockIdx.x * kSortChunk;
    for (int k = threadIdx.x; k < kSortChunk; k += kSortBlock) {
        const int64_t i = first + k;
        if (i < np) atomicAdd(&cnt[binid[i]], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += kSortBlock) {
        const unsigned c = cnt[b];
        base[b] = c ? atomicAdd(&cursor[b], c) : 0u;
        cnt[b] = 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kSortChunk; k += kSortBlock) {
        const int64_t i = first + k;
        if (i < np) {
            const int b = binid[i];
            const size_t d = (size_t)base[b] + atomicAdd(&cnt[b], 1u);
            x2[d] = x[i]; v2[d] = v[i]; ep2[d] = ep[i];
            perm2[d] = perm ? perm[i] : index_base + (uint32_t)i;
        }
    }
}

__global__ void k_unpermute2(int64_t np, const uint32_t *__restrict__ perm, const double2 *__restrict__ a,
                             double2 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[perm[i]] = a[i];
}

}  // namespace

int sort_bins(const MeshDev &m, int bin_cells_log2) {
    const int nbx = (m.nx + (1 << bin_cells_log2) - 1) >> bin_cells_log2, nby = (m.ny + (1 << bin_cells_log2) - 1) >> bin_cells_log2;
    return nbx * nby;
}

// x, v, ep, perm -> x2, v2, ep2, perm2 ordered by bin; binid: np uint16; hist: nbins unsigned (scratch).  perm may be
// null (identity: slot i gets index_base + i, so that a sub-range of a larger array can be sorted on its own).
// nbins = sort_bins(m, bin_cells_log2) must be <= 4096.
cudaError_t launch_sort_particles(const LaunchCtx &c, const MeshDev &m, int bin_cells_log2, int64_t np, const double2 *x,
                                  const double2 *v, const double2 *ep, const uint32_t *perm, double2 *x2, double2 *v2,
                                  double2 *ep2, uint32_t *perm2, uint16_t *binid, unsigned *hist, uint32_t index_base) {
    if (np <= 0) return cudaSuccess;
    SortGeom g;
    g.inv_dx = 1.0 / m.dx; g.inv_dy = 1.0 / m.dy; g.inv_nx = 1.0 / (double)m.nx; g.inv_ny = 1.0 / (double)m.ny;
    g.fnx = (double)m.nx; g.fny = (double)m.ny;
    g.nx = m.nx; g.ny = m.ny; g.shift = bin_cells_log2;
    g.nbx = (m.nx + (1 << bin_cells_log2) - 1) >> bin_cells_log2;
    g.nbins = sort_bins(m, bin_cells_log2);
    if (g.nbins > 4096 || np >= ((int64_t)1 << 32)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)g.nbins, c.stream);
    if (e != cudaSuccess) return e;
    const int grid = (int)((np + kSortChunk - 1) / kSortChunk);
    k_sort_count<<<grid, kSortBlock, sizeof(unsigned) * g.nbins, c.stream>>>(g, np, x, binid, hist);
    k_sort_scan<<<1, 1024, 0, c.stream>>>(g.nbins, hist);
    k_sort_scatter<<<grid, kSortBlock, 2 * sizeof(unsigned) * g.nbins, c.stream>>>(g.nbins, np, binid, hist, x, v, ep, perm, x2, v2, ep2, perm2, index_base);
    if (c.launches) *c.launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_unpermute(const LaunchCtx &c, int64_t np, const uint32_t *perm, const double2 *a, double2 *out) {
    if (np <= 0) return cudaSuccess;
    k_unpermute2<<<(unsigned)((np + 255) / 256), 256, 0, c.stream>>>(np, perm, a, out);
    if (c.launches) *c.launches += 1;
    return cudaGetLastError();
}

}  // namespace uapic

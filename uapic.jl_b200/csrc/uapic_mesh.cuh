// uapic_mesh.cuh -- device helpers shared by the mesh kernels (uapic_kernels.cu) and the 3D path (uapic_mrc3d.cu):
// fixed-order block reduction and the shared-memory FFT of one mesh line.
#pragma once

#include "uapic_device.cuh"

namespace uapic {

constexpr int kMeshBlock = 1024;

// deterministic block sum (fixed tree), result valid on every thread
DEVINL double block_sum(double v, double *sh) {
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// ---- shared-memory FFT of one line (power of two: radix-2; otherwise direct DFT) -----------------------------
DEVINL void line_twiddles(cd *tw, int n) {   // tw[k] = exp(-2 pi i k/n), k < n
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)n, &s, &c);
        tw[k] = mk(c, s);
    }
}

// in-place transform of a[0..n) ; tmp[0..n) scratch ; sign -1 forward / +1 backward ; ends with __syncthreads
DEVINL void line_fft(cd *a, cd *tmp, const cd *tw, int n, int sign) {
    const bool pow2 = (n & (n - 1)) == 0;
    __syncthreads();
    if (pow2) {
        int logn = 0;
        while ((1 << logn) < n) ++logn;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int r = (int)(__brev((unsigned)i) >> (32 - logn));
            if (logn > 0 && i < r) { const cd u = a[i]; a[i] = a[r]; a[r] = u; }
        }
        __syncthreads();
        for (int len = 2; len <= n; len <<= 1) {
            const int half = len >> 1, step = n / len;
            for (int bfly = threadIdx.x; bfly < n / 2; bfly += blockDim.x) {
                const int k = bfly & (half - 1);
                const int s = (bfly / half) * len;
                cd w = tw[k * step];
                if (sign > 0) w.im = -w.im;
                const cd u = a[s + k];
                const cd v = cmul(w, a[s + k + half]);
                a[s + k] = cadd(u, v);
                a[s + k + half] = csub(u, v);
            }
            __syncthreads();
        }
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            cd acc = mk(0.0, 0.0);
            for (int j = 0; j < n; ++j) {
                cd w = tw[(int)(((long long)k * j) % n)];
                if (sign > 0) w.im = -w.im;
                acc = cadd(acc, cmul(w, a[j]));
            }
            tmp[k] = acc;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) a[k] = tmp[k];
        __syncthreads();
    }
}

}  // namespace uapic

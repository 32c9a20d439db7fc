// uapic_fast.cuh -- reciprocal-based mesh helpers shared by the fused phase kernels (uapic_fused.cu) and the
// one-pass kernels (uapic_onepass.cu): periodic cell lookup, wrap-free halo gather, branch-free M6 weights.
#pragma once

#include "uapic_device.cuh"

namespace uapic {

// ---- mesh access with reciprocals -------------------------------------------------------------------------
struct MeshFast {
    double inv_dx, inv_dy, inv_nx, inv_ny, inv_dimx, inv_dimy;
};

DEVINL Cell cell_fast(const MeshDev &m, const MeshFast &f, double x, double y, int wrap, double &xw, double &yw) {
    double px, py;
    if (wrap == kWrapJulia) {
        const double xn = modulo_fast(x - m.xmin, m.dimx, f.inv_dimx);
        const double yn = modulo_fast(y - m.ymin, m.dimy, f.inv_dimy);
        px = xn * f.inv_dx; py = yn * f.inv_dy;
        xw = xn + m.xmin; yw = yn + m.ymin;
    } else {
        px = modulo_fast(x * f.inv_dx, (double)m.nx, f.inv_nx);
        py = modulo_fast(y * f.inv_dy, (double)m.ny, f.inv_ny);
        xw = x; yw = y;
    }
    Cell c;
    c.i = __double2int_rd(px); c.dpx = px - (double)c.i;
    c.j = __double2int_rd(py); c.dpy = py - (double)c.j;
    // memory safety only: a non-finite position (NaN input, b -> 0) must not turn into an out-of-range mesh address
    c.i = min(max(c.i, 0), m.nx - 1);
    c.j = min(max(c.j, 0), m.ny - 1);
    return c;
}

// i in [0,n], off in [-2,3], n >= 4: one conditional add and one conditional subtract replace the integer modulo.
// The centre keeps the reference's unwrapped index (compute_rho_m6.F90:102-116).
DEVINL int wrap_fast(int i, int off, int n) {
    if (off == 0) return i;
    int r = i + off;
    r += (r < 0) ? n : 0;
    r -= (r >= n) ? n : 0;
    return r;
}

// separable M6 gather on the periodic halo copy of E: no index wrap, one base address per row and immediate offsets
// for its 6 nodes; (ex,ey) pairs read as 16-byte words through the read-only path
DEVINL void gather_fast(const MeshDev &m, const double2 *__restrict__ ehalo, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    const int ldx = m.nx + 6;
    const double2 *row = ehalo + (c.j * ldx + c.i);      // node (i-2, j-2)
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double2 ev = __ldg(row + a);
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
        row += ldx;
    }
    e1 = s1; e2 = s2;
}

// ---- tiled halo copy of E for the one-pass kernels -------------------------------------------------------------
// Same nodes as the linear halo copy, i in [-2, nx+3], j in [-2, ny+3], but every 128-byte line holds a 2 (x) by 4 (y)
// block of nodes: node (I, J) = (i+2, j+2) lives at [ (J/4 * ntx + I/2) * 8 + (J%4) * 2 + I%2 ].
// Why: a gather instruction reads the SAME tap for the 32 tau samples of one particle, i.e. 32 points on its gyro-orbit;
// the L1 data pipe spends one wavefront per distinct line those points fall in (ncu: 8.5 tag lookups, 7.3 wavefronts per
// LDG.128 with 8 x 1 lines; a seeded model of the orbits reproduces 8.47).  Compact 2 x 4 blocks cut that to 5.2.
DEVINL int halo_tiled_ntx(const MeshDev &m) { return (m.nx + 6 + 1) >> 1; }
DEVINL int halo_tiled_nty(const MeshDev &m) { return (m.ny + 6 + 3) >> 2; }
DEVINL int halo_tiled_index(int ntx, int I, int J) { return (((J >> 2) * ntx + (I >> 1)) << 3) + ((J & 3) << 1) + (I & 1); }

#ifndef UAPIC_OP_TAP_EVICT_LAST
#define UAPIC_OP_TAP_EVICT_LAST 0     // 1: the taps ask L1 to keep their lines (ld.global.nc.L1::evict_last)
#endif
DEVINL double2 ldg_tap(const double2 *p) {
#if UAPIC_OP_TAP_EVICT_LAST
    double2 r;
    asm("ld.global.nc.L1::evict_last.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}

DEVINL void gather_tiled(const MeshDev &m, const double2 *__restrict__ ehalo, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    const int ntx8 = halo_tiled_ntx(m) << 3;
    const int p = c.i & 1;
    int q = c.j & 3;
    int roff = (c.j >> 2) * ntx8 + (q << 1) + ((c.i >> 1) << 3) + p;      // node (i-2, j-2)
    const int step_odd = p ? 7 : 1;                                        // from an even tap to the next (odd) one
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const double2 *re = ehalo + roff;          // taps 0, 2, 4 at +0, +8, +16
        const double2 *ro = re + step_odd;         // taps 1, 3, 5
        double r1 = 0.0, r2 = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double2 ev = ldg_tap(((a & 1) ? ro : re) + 8 * (a >> 1));
            r1 = fma(cx[a], ev.x, r1);
            r2 = fma(cx[a], ev.y, r2);
        }
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
        roff += (q == 3) ? ntx8 - 6 : 2;
        q = (q + 1) & 3;
    }
    e1 = s1; e2 = s2;
}

// ---- the same gather with 256-bit loads (sm_100a: LDG.E.256) -----------------------------------------------------
// In the 2 x 4 tiling the two nodes (I, J), (I+1, J) with I even are 32 contiguous, 32-byte-aligned bytes, so one
// ld.global.nc.v4.f64 fetches a PAIR of taps.  The L1 data pipe spends one wavefront per distinct 128-byte line an
// instruction touches, whatever the access width (B300_MICROARCH.md, "L1tex wavefront queue"), and both nodes of a pair
// sit in the same line: a row of 6 taps costs 3 pair loads (+ one predicated 16-byte load of the 7th node when the
// stencil starts on an odd node) instead of 6 loads -- 21 load instructions per gather on average instead of 36, at the
// same number of lines per instruction.  The 6 weights are shifted into a 7-wide window by the parity of c.i.
DEVINL void ldg256(const double2 *p, double2 &a, double2 &b) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}

DEVINL void gather_tiled_pairs(const MeshDev &m, const double2 *__restrict__ ehalo, const Cell &c, double &e1, double &e2) {
    double cx[6], cy[6];
    m6_weights_fast(c.dpx, cx);
    m6_weights_fast(c.dpy, cy);
    const int ntx8 = halo_tiled_ntx(m) << 3;
    const bool odd = c.i & 1;
    double w[7];
    w[0] = odd ? 0.0 : cx[0];
#pragma unroll
    for (int a = 1; a < 6; ++a) w[a] = odd ? cx[a - 1] : cx[a];
    w[6] = odd ? cx[5] : 0.0;
    int q = c.j & 3;
    int roff = (c.j >> 2) * ntx8 + (q << 1) + ((c.i >> 1) << 3);           // even node (i-2 or i-3, j-2)
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const double2 *r = ehalo + roff;
        double2 n0, n1, n2, n3, n4, n5, n6 = make_double2(0.0, 0.0);
        ldg256(r, n0, n1);
        ldg256(r + 8, n2, n3);
        ldg256(r + 16, n4, n5);
        if (odd) n6 = __ldg(r + 24);
        double r1 = w[0] * n0.x, r2 = w[0] * n0.y;
        r1 = fma(w[1], n1.x, r1); r2 = fma(w[1], n1.y, r2);
        r1 = fma(w[2], n2.x, r1); r2 = fma(w[2], n2.y, r2);
        r1 = fma(w[3], n3.x, r1); r2 = fma(w[3], n3.y, r2);
        r1 = fma(w[4], n4.x, r1); r2 = fma(w[4], n4.y, r2);
        r1 = fma(w[5], n5.x, r1); r2 = fma(w[5], n5.y, r2);
        r1 = fma(w[6], n6.x, r1); r2 = fma(w[6], n6.y, r2);
        s1 = fma(cy[b], r1, s1);
        s2 = fma(cy[b], r2, s2);
        roff += (q == 3) ? ntx8 - 6 : 2;
        q = (q + 1) & 3;
    }
    e1 = s1; e2 = s2;
}

// ---- CIC (bilinear), BUILD-DEFINED: the reference has no 2D CIC deposit and no CIC in the UA loop.
// Weights of performance/test_cic.F90:73-76 (= 2D restriction of fortran/compute_rho_cic.f90:46-53) on nodes (i,j), (i+1,j),
// (i+1,j+1), (i,j+1); wrap, ghost copy, /(dx dy) and neutralisation as on the M6 path.  oracle/uapic_oracle.c states the same.
DEVINL void gather_cic_tiled(const MeshDev &m, const double2 *__restrict__ ehalo, const Cell &c, double &e1, double &e2) {
    const int ntx = halo_tiled_ntx(m);
    const int I = c.i + 2, J = c.j + 2;                 // halo coordinates of node (i, j); (i+1, j+1) never needs a wrap there
    const double2 e00 = __ldg(ehalo + halo_tiled_index(ntx, I, J)), e10 = __ldg(ehalo + halo_tiled_index(ntx, I + 1, J));
    const double2 e01 = __ldg(ehalo + halo_tiled_index(ntx, I, J + 1)), e11 = __ldg(ehalo + halo_tiled_index(ntx, I + 1, J + 1));
    const double ax = 1.0 - c.dpx, ay = 1.0 - c.dpy;
    const double a1 = ax * ay, a2 = c.dpx * ay, a3 = c.dpx * c.dpy, a4 = ax * c.dpy;
    e1 = fma(a4, e01.x, fma(a3, e11.x, fma(a2, e10.x, a1 * e00.x)));
    e2 = fma(a4, e01.y, fma(a3, e11.y, fma(a2, e10.y, a1 * e00.y)));
}

// tap q of the 4-tap CIC deposit: 0 (i,j), 1 (i+1,j), 2 (i,j+1), 3 (i+1,j+1)
DEVINL void deposit_cic_tap(const MeshDev &m, const RhoAcc &r, const Cell &c, double weight, int q) {
    const int ox = q & 1, oy = q >> 1;
    const double wx = ox ? c.dpx : 1.0 - c.dpx, wy = oy ? c.dpy : 1.0 - c.dpy;
    rho_add(r, wrap_fast(c.i, ox, m.nx) + wrap_fast(c.j, oy, m.ny) * m.ld, wx * wy * weight);
}

// f_m6 without branches (q >= 0): the three pieces of compute_rho_m6.F90:33-41 are the same sum of truncated powers
DEVINL double f_m6_branchless(double q) {
    const double a = fmax(3.0 - q, 0.0), b = fmax(2.0 - q, 0.0), c = fmax(1.0 - q, 0.0);
    return fma(15.0, pow5(c), fma(-6.0, pow5(b), pow5(a))) * (1.0 / 120.0);
}

}  // namespace uapic

"""External-field two-scale program of the reference, numpy twin -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Second, independently written restatement of fortran/efd.f90 (the C one is `orc_efd_run` in uapic_oracle.c): all
particles at once, arrays (np, ntau), `numpy.fft` for the tau transforms.  It follows the FORTRAN conventions --
forward transforms carry 1/ntau (fft.f90:44-72) -- because the Julia twin test/test_efd.jl is an unfinished port
(unnormalised `mul!(tilde, ftau, r)` followed by `ave .= real(tilde[1,:])/ep`, test_efd.jl:121-124, and a `return` at
:197); its operation order differs from the C file's (vectorised primitives, complex `interv`), so agreement of the
two is a check of both.

PINNED by the reference: with init_particles_2d's own load it reproduces the constants printed on efd.f90:481
(tests/test_efd_oracle.py).
"""
from __future__ import annotations

import numpy as np


def _fft(a):
    return np.fft.fft(a, axis=1) / a.shape[1]


def _ifft(a):
    return np.fft.ifft(a, axis=1) * a.shape[1]


def efd_run(x, v, ntau=16, eps=1e-3, dt=np.pi / 16, tfinal=np.pi / 2, box=(0.0, 4 * np.pi, 0.0, 2 * np.pi), stages=None,
            real=np.float64):
    """x, v: (2, np).  Returns (x, v) at tfinal.  `stages`, if a dict, receives ave, xt, yt after the preparation.
    `real=np.longdouble` evaluates the same formulas in x87 extended precision (64-bit mantissa, numpy's FFT included) from the
    same double inputs: the REFEREE the double implementations are ranked against (tests/test_efd_oracle.py, test_gpu_efd.py)."""
    x1, x2, v1, v2 = (np.asarray(a, dtype=np.float64).astype(real)[:, None] for a in (x[0], x[1], v[0], v[1]))
    eps, dt, tfinal = real(eps), real(dt), real(tfinal)
    pi = real(np.pi) if real is np.float64 else 4 * np.arctan(real(1))
    nstep = int(round(float(tfinal / dt)))                                           # efd.f90:102
    m = ntau // 2
    ltau = np.concatenate([np.arange(0, m), np.arange(-m, 0)]).astype(real)[None, :]   # efd.f90:111-112
    tau = (np.arange(ntau).astype(real) * (2 * pi / ntau))[None, :]
    c, s = np.cos(tau), np.sin(tau)
    inv = np.zeros_like(ltau)
    inv[0, 1:] = 1.0 / ltau[0, 1:]
    time = real(0)
    bx = 1 + 0.5 * np.sin(x1) * np.sin(x2)                                           # efd.f90:139
    ds = dt * bx

    def prim(a):                      # zero-mean tau-primitive: -i a_l / l, l != 0 (efd.f90:183-188)
        return _ifft(-1j * _fft(a) * inv)

    def db(a, b):                     # (b(X) - b(x)) / b(x)
        return (1 + 0.5 * np.sin(a.real) * np.sin(b.real) - bx) / bx

    def first(a):
        return a[:, :1]

    # efd.f90:157-195
    h1, h2 = eps * (s * v1 / bx - c * v2 / bx), eps * (s * v2 / bx + c * v1 / bx)
    xt1, xt2 = x1 + h1 - first(h1), x2 + h2 - first(h2)
    e1 = (0.5 * np.cos(x1 / 2) * np.sin(x2)) * (1 + 0.5 * np.sin(time))
    e2 = (np.sin(x1 / 2) * np.cos(x2)) * (1 + 0.5 * np.sin(time))
    q = db(xt1, xt2)
    r1, r2 = q * v2 + 0j, -q * v1 + 0j
    ave1, ave2 = first(_fft(r1)).real / eps, first(_fft(r2)).real / eps
    r1, r2 = prim(r1), prim(r2)
    r1 = eps * (s * e1 + c * e2) / bx + r1
    r2 = eps * (s * e2 - c * e1) / bx + r2
    yt1, yt2 = v1 + (r1 - first(r1)), v2 + (r2 - first(r2))
    # efd.f90:200-222
    h1 = prim(eps * (c * yt1 + s * yt2) / bx) - eps ** 2 / bx * (-c * ave1 - s * ave2)
    h2 = prim(eps * (c * yt2 - s * yt1) / bx) - eps ** 2 / bx * (-c * ave2 + s * ave1)
    xt1, xt2 = x1 + h1 - first(h1), x2 + h2 - first(h2)
    # efd.f90:226-310
    e1 = (0.5 * np.cos(x1 / 2) * np.sin(x2)) * 0.5 * np.cos(time)
    e2 = (np.sin(x1 / 2) * np.cos(x2)) * 0.5 * np.cos(time)
    q = db(xt1, xt2)
    fx1, fx2 = q * ave2, -q * ave1
    fy1, fy2 = eps / bx * (s * ave1 - c * ave2), eps / bx * (c * ave1 + s * ave2)
    w = np.cos(x1) * np.sin(x2) * fy1 + np.sin(x1) * np.cos(x2) * fy2
    fy1, fy2 = w / bx / 2 * v2 + fx1, -w / bx / 2 * v1 + fx2
    fx1, fx2 = eps / bx ** 2 * (-s * e2 + c * e1), eps / bx ** 2 * (s * e1 + c * e2)
    t1, t2 = _fft(fy1 + fx1 + 0j), _fft(fy2 + fx2 + 0j)
    r1, r2 = -eps * _ifft(-t1 * inv ** 2), -eps * _ifft(-t2 * inv ** 2)
    fy1, fy2 = _ifft(-1j * t1 * inv), _ifft(-1j * t2 * inv)
    e1 = (0.5 * np.cos(xt1.real / 2) * np.sin(xt2.real)) * (1 + 0.5 * np.sin(time))
    e2 = (np.sin(xt1.real / 2) * np.cos(xt2.real)) * (1 + 0.5 * np.sin(time))
    q = db(xt1, xt2)
    t1 = q * yt2 + eps / bx * (-s * e2 + c * e1)
    t2 = -q * yt1 + eps / bx * (s * e1 + c * e2)
    yd1, yd2 = first(_fft(t1)) / eps, first(_fft(t2)) / eps
    r1, r2 = r1 + prim(t1), r2 + prim(t2)
    yt1, yt2 = v1 + r1 - first(r1), v2 + r2 - first(r2)
    # efd.f90:315-383
    m1, m2 = first(_fft((c * r1 + s * r2) / bx)), first(_fft((c * r2 - s * r1) / bx))
    w0 = np.cos(x1) * np.sin(x2) * m1 + np.sin(x1) * np.cos(x2) * m2
    g1, g2 = w0 / eps / bx * v2 / 2, -w0 / eps / bx * v1 / 2
    q0 = first(_fft(db(xt1, xt2) + 0j))
    g1, g2 = g1 + q0 / eps * ave2, g2 - q0 / eps * ave1
    yf1, yf2 = yd1 + fy1, yd2 + fy2
    t1, t2 = prim(c * yf1 + s * yf2), prim(c * yf2 - s * yf1)
    fy1 = t1 * eps / bx - eps ** 2 / bx * (-c * g1 - s * g2)
    fy2 = t2 * eps / bx - eps ** 2 / bx * (-c * g2 + s * g1)
    h1 = -eps * prim(fy1) + prim(eps * (c * yt1 + s * yt2) / bx)
    h2 = -eps * prim(fy2) + prim(eps * (c * yt2 - s * yt1) / bx)
    xt1, xt2 = x1 + h1 - first(h1), x2 + h2 - first(h2)
    if stages is not None:
        stages.update(ave=np.stack([ave1[:, 0], ave2[:, 0]]), xt=np.stack([xt1, xt2]), yt=np.stack([yt1, yt2]))

    def force(a1, a2, b1, b2, t):     # efd.f90:509-524
        ee1 = (0.5 * np.cos(a1.real / 2) * np.sin(a2.real)) * (1 + 0.5 * np.sin(t))
        ee2 = (np.cos(a2.real) * np.sin(a1.real / 2)) * (1 + 0.5 * np.sin(t))
        qq = db(a1, a2) / eps
        return (c * ee1 - s * ee2) / bx + qq * b2, (c * ee2 + s * ee1) / bx - qq * b1

    den = 1.0 + 1j * ds / 2 * ltau / eps
    num = 1.0 - 1j * ds / eps / 2 * ltau
    for _ in range(nstep):            # efd.f90:388-454
        fy1, fy2 = force(xt1, xt2, yt1, yt2, time)
        yf1, yf2 = _ifft(_fft(yt1 + ds / 2 * fy1) / den), _ifft(_fft(yt2 + ds / 2 * fy2) / den)
        fx1, fx2 = (c * yf1 + s * yf2) / bx, (c * yf2 - s * yf1) / bx
        xf1, xf2 = _ifft(_fft(xt1 + ds / 2 * fx1) / den), _ifft(_fft(xt2 + ds / 2 * fx2) / den)
        time = time + dt / 2
        fy1, fy2 = force(xf1, xf2, yf1, yf2, time)
        n1, n2 = _ifft((_fft(yt1) * num + ds * _fft(fy1)) / den), _ifft((_fft(yt2) * num + ds * _fft(fy2)) / den)
        yf1, yf2 = (n1 + yt1) / 2, (n2 + yt2) / 2
        yt1, yt2 = n1, n2
        fx1, fx2 = (c * yf1 + s * yf2) / bx, (c * yf2 - s * yf1) / bx
        xt1, xt2 = _ifft((_fft(xt1) * num + ds * _fft(fx1)) / den), _ifft((_fft(xt2) * num + ds * _fft(fx2)) / den)
        time = time + dt / 2
    # efd.f90:456-478, 526-544
    ph = np.exp(1j * ltau * tfinal * bx / eps)
    xo = np.stack([(_fft(xt1) * ph).sum(1).real, (_fft(xt2) * ph).sum(1).real])
    for d, (lo, hi) in enumerate(((real(box[0]), real(box[1])), (real(box[2]), real(box[3])))):
        span = hi - lo
        for _ in range(64):
            over, under = xo[d] > hi, xo[d] < lo
            if not (over.any() or under.any()):
                break
            xo[d] = np.where(over, xo[d] - span, np.where(under, xo[d] + span, xo[d]))
    w1, w2 = (_fft(yt1) * ph).sum(1), (_fft(yt2) * ph).sum(1)
    cb, sb = np.cos(tfinal * bx[:, 0] / eps), np.sin(tfinal * bx[:, 0] / eps)
    return xo, np.stack([(cb * w1 + sb * w2).real, (cb * w2 - sb * w1).real])

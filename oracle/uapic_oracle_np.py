"""numpy twin of the CPU oracle -- TEST INFRASTRUCTURE, not the product.

Written independently of ``uapic_oracle.c`` and following the *Julia* sources of
JuliaVlasov/UAPIC.jl (``src/*.jl`` + ``test/bupdate.jl``), whereas the C oracle follows the
Fortran.  The two conventions differ in wrap order, FFT normalisation and where ``yt`` lives
after the corrector; tests/test_oracle.py requires both to agree to
<= 1e-12, which is the mitigation for "parity unpinned by the reference".

numpy's pocketfft has FFTW's semantics for irfft on non-Hermitian input (imaginary parts of the
DC / Nyquist bins of the halved axis are ignored), which is what ``src/poisson.jl:72-73`` relies on.

Arrays are numpy arrays in *Julia index order*, e.g. ``xt[n, c, k]`` with shape (ntau, 2, np);
use ``order='F'`` buffers when exchanging memory with the C oracle / the CUDA library.
"""
from __future__ import annotations

import contextlib

import numpy as np

# Working precision.  float64 is the twin of the reference.  "longdouble" (x87 80-bit, 64-bit mantissa, eps 1.1e-19) turns
# this file into the EXTENDED-PRECISION REFEREE of tests/golden/make_referee.py: the same formulas evaluated from the same
# double inputs with 2000x smaller rounding, which tells which of two double implementations (C oracle, CUDA kernels) is
# closer to the exact result when they disagree -- at small eps every rounding of b(x) is multiplied by t/eps in the phase
# l*t/eps (ua_steps.F90:64,224,258,297).  numpy's pocketfft, sin/cos/exp and mod all run in long double for these dtypes.
REAL = np.float64
CPLX = np.complex128


def set_precision(name: str):
    global REAL, CPLX
    REAL, CPLX = {"double": (np.float64, np.complex128), "longdouble": (np.longdouble, np.clongdouble)}[name]


@contextlib.contextmanager
def precision(name: str):
    old = "double" if REAL is np.float64 else "longdouble"
    set_precision(name)
    try:
        yield
    finally:
        set_precision(old)


def _pi():
    return np.pi if REAL is np.float64 else REAL(4) * np.arctan(REAL(1))


# ----------------------------------------------------------------------------------------------
# types                                 src/meshfields.jl:3-45, src/ua_type.jl:17-41
# ----------------------------------------------------------------------------------------------
class Mesh:
    def __init__(self, xmin, xmax, nx, ymin, ymax, ny):
        self.xmin, self.xmax, self.nx = REAL(xmin), REAL(xmax), int(nx)
        self.ymin, self.ymax, self.ny = REAL(ymin), REAL(ymax), int(ny)
        self.dx = (self.xmax - self.xmin) / self.nx      # meshfields.jl:16
        self.dy = (self.ymax - self.ymin) / self.ny      # meshfields.jl:17


def ua_tables(ntau):
    """tau[i] = i*2pi/ntau ; ltau = [0:ntau/2-1 ; -ntau/2:-1]      src/ua_type.jl:19-25"""
    dtau = 2 * _pi() / ntau
    ltau = np.concatenate([np.arange(0, ntau // 2), np.arange(-ntau // 2, 0)]).astype(REAL)
    tau = np.array([i * dtau for i in range(ntau)], dtype=REAL)
    return tau, ltau


# ----------------------------------------------------------------------------------------------
# M6                                                        src/compute_rho.jl:10-25
# ----------------------------------------------------------------------------------------------
def f_m6(q):
    q = np.asarray(q, dtype=REAL)
    a = (3 - q) ** 5
    b = (2 - q) ** 5
    c = (1 - q) ** 5
    out = np.where(q < 1.0, a - 6 * b + 15 * c,
                   np.where(q < 2.0, a - 6 * b, np.where(q < 3.0, a, 0.0)))
    return out / 120


# shape function: "m6" (what the reference ships) or "cic" (BUILD-DEFINED: bilinear weights of
# performance/test_cic.F90:73-76 = 2D restriction of fortran/compute_rho_cic.f90:46-53, with the wrap / ghost copy /
# scaling / neutralisation of the M6 path; the reference has no 2D CIC deposit and no CIC in the UA loop)
SCHEME = "m6"


def _cell_weights(mesh, x, y):
    """src/interpolation.jl:19-64 / src/compute_rho.jl:63-110 : wrap, cell, 7+7 weights, wrapped indices.
    returns (xw, yw, ix[7,...], jy[7,...], cx[7,...], cy[7,...]) with 0-based node indices."""
    dimx = mesh.xmax - mesh.xmin
    dimy = mesh.ymax - mesh.ymin
    xn = np.mod(x - mesh.xmin, dimx)
    yn = np.mod(y - mesh.ymin, dimy)
    px = xn / mesh.dx
    py = yn / mesh.dy
    i = np.trunc(px).astype(np.int64)
    j = np.trunc(py).astype(np.int64)
    dpx = px - i
    dpy = py - j
    offs = np.arange(-3, 4)
    ix = np.stack([(i if a == 0 else np.mod(i + a, mesh.nx)) for a in offs])
    jy = np.stack([(j if a == 0 else np.mod(j + a, mesh.ny)) for a in offs])
    if SCHEME == "cic":
        z = np.zeros_like(dpx)
        cx = np.stack([z, z, z, 1 - dpx, dpx, z, z])
        cy = np.stack([z, z, z, 1 - dpy, dpy, z, z])
    else:
        cx = np.stack([f_m6(3 + dpx), f_m6(2 + dpx), f_m6(1 + dpx), f_m6(dpx), f_m6(1 - dpx), f_m6(2 - dpx), f_m6(3 - dpx)])
        cy = np.stack([f_m6(3 + dpy), f_m6(2 + dpy), f_m6(1 + dpy), f_m6(dpy), f_m6(1 - dpy), f_m6(2 - dpy), f_m6(3 - dpy)])
    return xn + mesh.xmin, yn + mesh.ymin, ix, jy, cx, cy


def _gather(mesh, e, ix, jy, cx, cy):
    """49-term sum, same order as src/interpolation.jl:66-116 (x offset outer, y offset inner)"""
    s1 = np.zeros(ix.shape[1:], dtype=REAL)
    s2 = np.zeros(ix.shape[1:], dtype=REAL)
    for a in range(7):
        for b in range(7):
            w = cx[a] * cy[b]
            s1 = s1 + w * e[0, ix[a], jy[b]]
            s2 = s2 + w * e[1, ix[a], jy[b]]
    return s1, s2


def _scatter_and_epilogue(mesh, rho, ix, jy, cx, cy, weight):
    nx, ny = mesh.nx, mesh.ny
    rho[:] = 0.0
    for a in range(7):
        for b in range(7):
            np.add.at(rho, (ix[a], jy[b]), cx[a] * cy[b] * weight)
    # src/compute_rho.jl:169-177
    rho[0:nx, ny] = rho[0:nx, 0]
    rho[nx, 0:ny] = rho[0, 0:ny]
    rho[nx, ny] = rho[0, 0]
    rho /= (mesh.dx * mesh.dy)
    rho_total = np.sum(rho[0:nx, 0:ny]) * mesh.dx * mesh.dy
    rho -= rho_total / (mesh.xmax - mesh.xmin) / (mesh.ymax - mesh.ymin)
    return rho_total


def compute_rho_m6(mesh, rho, x, w):
    """compute_rho_m6!(fields, particles)     src/compute_rho.jl:181-316   (x is (2,np), wrapped in place)"""
    xw, yw, ix, jy, cx, cy = _cell_weights(mesh, x[0], x[1])
    x[0], x[1] = xw, yw
    return _scatter_and_epilogue(mesh, rho, ix, jy, cx, cy, w)


def interpol_eb_m6(mesh, e, x, ep):
    """interpol_eb_m6!(particles, fields)     src/interpolation.jl:125-247"""
    xw, yw, ix, jy, cx, cy = _cell_weights(mesh, x[0], x[1])
    x[0], x[1] = xw, yw
    ep[0], ep[1] = _gather(mesh, e, ix, jy, cx, cy)


def interpol_eb_m6_tau(mesh, e, xt, et):
    """interpol_eb_m6!(e, fields, x, nbpart, ntau)     src/interpolation.jl:3-123"""
    _, _, ix, jy, cx, cy = _cell_weights(mesh, xt[:, 0, :].real, xt[:, 1, :].real)
    et[:, 0, :], et[:, 1, :] = _gather(mesh, e, ix, jy, cx, cy)


def compute_rho_m6_tau(mesh, rho, x, w, xt, t, eps):
    """compute_rho_m6!(fields, particles, xt, ua)     src/compute_rho.jl:29-179"""
    ntau = xt.shape[0]
    _, ltau = ua_tables(ntau)
    ph = np.exp(1j * ltau[:, None] * t[None, :] / eps) / ntau                    # :50, :58
    xt1 = np.real(np.sum(np.fft.fft(xt[:, 0, :], axis=0) * ph, axis=0))          # :47-52
    xt2 = np.real(np.sum(np.fft.fft(xt[:, 1, :], axis=0) * ph, axis=0))          # :55-60
    xw, yw, ix, jy, cx, cy = _cell_weights(mesh, xt1, xt2)
    x[0], x[1] = xw, yw                                                            # :69-70
    return _scatter_and_epilogue(mesh, rho, ix, jy, cx, cy, w)


# ----------------------------------------------------------------------------------------------
# Poisson                                                   src/poisson.jl:14-83
# ----------------------------------------------------------------------------------------------
class Poisson:
    def __init__(self, mesh):
        nx, ny = mesh.nx, mesh.ny
        kx0 = 2 * _pi() / (mesh.xmax - mesh.xmin)
        ky0 = 2 * _pi() / (mesh.ymax - mesh.ymin)
        kx = np.zeros((nx // 2 + 1, ny), dtype=REAL)
        ky = np.zeros((nx // 2 + 1, ny), dtype=REAL)
        for ik in range(nx // 2 + 1):
            kx[ik, :] = ik * kx0
        for jk in range(ny // 2):
            ky[:, jk] = jk * ky0
        for jk in range(ny // 2, ny):
            ky[:, jk] = (jk - ny) * ky0
        kx[0, 0] = 1.0
        k2 = kx * kx + ky * ky
        self.kx = kx / k2
        self.ky = ky / k2
        self.mesh = mesh

    def __call__(self, rho, e):
        """rfft halves the FIRST Julia dimension (x): numpy axes=(1,0) puts the halved axis last-transformed = axis 0"""
        m = self.mesh
        nx, ny = m.nx, m.ny
        rt = np.fft.fft(np.fft.rfft(rho[0:nx, 0:ny], axis=0), axis=1)           # poisson.jl:67
        ex = -1j * self.kx * rt
        ey = -1j * self.ky * rt
        e[0, 0:nx, 0:ny] = np.fft.irfft(np.fft.ifft(ex, axis=1), n=nx, axis=0)  # :72
        e[1, 0:nx, 0:ny] = np.fft.irfft(np.fft.ifft(ey, axis=1), n=nx, axis=0)  # :73
        e[0, nx, :] = e[0, 0, :]
        e[0, :, ny] = e[0, :, 0]
        e[1, nx, :] = e[1, 0, :]
        e[1, :, ny] = e[1, :, 0]
        return np.sum(e[0] * e[0] + e[1] * e[1]) * m.dx * m.dy                   # :80-81


# ----------------------------------------------------------------------------------------------
# UA stages                                                 src/ua_steps.jl
# ----------------------------------------------------------------------------------------------
def preparation(ntau, eps, dt, x, v, ep):
    """preparation!     src/ua_steps.jl:3-78 ; returns b, t, pl, ql, xt, yt"""
    tau, ltau = ua_tables(ntau)
    x1, x2 = x[0], x[1]
    b = 1 + 0.5 * np.sin(x1) * np.sin(x2)                                        # :20
    t = dt * b                                                                   # :21
    npart = x.shape[1]
    pl = np.zeros((ntau, npart), dtype=CPLX)
    ql = np.zeros((ntau, npart), dtype=CPLX)
    pl[0] = t                                                                    # :23
    ql[0] = t ** 2 / 2                                                           # :24
    l = ltau[1:, None]
    elt = np.exp(-1j * l * t[None, :] / eps)                                     # :30
    pl[1:] = eps * 1j * (elt - 1) / l                                            # :31
    ql[1:] = eps * (eps * (1 - elt) - 1j * l * t[None, :]) / l ** 2              # :32
    ex, ey = ep[0], ep[1]
    vx, vy = v[0], v[1]
    vxb, vyb = vx / b, vy / b
    st, ct = np.sin(tau)[:, None], np.cos(tau)[:, None]
    h1 = eps * (st * vxb - ct * vyb)                                             # :43
    h2 = eps * (st * vyb + ct * vxb)                                             # :44
    xt1 = x1 + h1 + eps * vyb                                                    # :46
    xt2 = x2 + h2 - eps * vxb                                                    # :47
    xt = np.zeros((ntau, 2, npart), dtype=CPLX)
    xt[:, 0, :], xt[:, 1, :] = xt1, xt2
    interv = (1 + 0.5 * np.sin(xt1) * np.sin(xt2) - b) / eps                     # :52
    exb = ((ct * vy - st * vx) * interv + ex) / b                                # :54
    eyb = ((-ct * vx - st * vy) * interv + ey) / b                               # :55
    r = np.zeros((ntau, 2, npart), dtype=CPLX)
    r[:, 0, :] = ct * exb - st * eyb                                             # :57
    r[:, 1, :] = st * exb + ct * eyb                                             # :58
    rt = np.fft.fft(r, axis=0)                                                   # :62
    rt[1:] = -1j * rt[1:] / ltau[1:, None, None]                                 # :64-67
    r = np.fft.ifft(rt, axis=0)                                                  # :69
    yt = np.zeros((ntau, 2, npart), dtype=CPLX)
    yt[:, 0, :] = vx + (r[:, 0, :] - r[0, 0, :]) * eps                           # :72
    yt[:, 1, :] = vy + (r[:, 1, :] - r[0, 1, :]) * eps                           # :73
    return b, t, pl, ql, xt, yt


def compute_f(eps, b, xt, yt, et):
    """compute_f!     src/ua_steps.jl:105-145  (unnormalised fft!)"""
    ntau = xt.shape[0]
    tau, _ = ua_tables(ntau)
    ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
    xt1, xt2 = xt[:, 0, :].real, xt[:, 1, :].real
    yt1, yt2 = yt[:, 0, :], yt[:, 1, :]
    fx = np.empty_like(xt)
    fy = np.empty_like(xt)
    fx[:, 0, :] = (ct * yt1 + st * yt2) / b                                      # :127
    fx[:, 1, :] = (-st * yt1 + ct * yt2) / b                                     # :128
    interv = (1 + 0.5 * np.sin(xt1) * np.sin(xt2) - b) / eps                     # :130
    tmp1 = et[:, 0, :] + (ct * yt2 - st * yt1) * interv                          # :132
    tmp2 = et[:, 1, :] + (-ct * yt1 - st * yt2) * interv                         # :133
    fy[:, 0, :] = (ct * tmp1 - st * tmp2) / b                                    # :135
    fy[:, 1, :] = (st * tmp1 + ct * tmp2) / b                                    # :136
    return np.fft.fft(fx, axis=0), np.fft.fft(fy, axis=0)                        # :142-143


def ua_step_predict(eps, t, pl, xf, fx):
    """ua_step! 5-arg (Fourier in, Fourier out)     src/ua_steps.jl:149-170"""
    ntau = xf.shape[0]
    _, ltau = ua_tables(ntau)
    elt = np.exp(-1j * ltau[:, None] * t[None, :] / eps)[:, None, :]
    return elt * xf + pl[:, None, :] * fx


def ua_step_correct(eps, t, pl, ql, xf, fx, gx):
    """ua_step! 6-arg     src/ua_steps.jl:172-200"""
    ntau = xf.shape[0]
    _, ltau = ua_tables(ntau)
    elt = np.exp(-1j * ltau[:, None] * t[None, :] / eps)[:, None, :]
    out = elt * xf + pl[:, None, :] * fx
    out = out + ql[:, None, :] * (gx - fx) / t[None, None, :]
    return out


def compute_v(eps, t, yt_fourier):
    """compute_v!  (consumes Fourier coefficients)     src/ua_steps.jl:204-224"""
    ntau = yt_fourier.shape[0]
    _, ltau = ua_tables(ntau)
    elt = np.exp(1j * ltau[:, None] * t[None, :] / eps)
    px = np.sum(yt_fourier[:, 0, :] / ntau * elt, axis=0)
    py = np.sum(yt_fourier[:, 1, :] / ntau * elt, axis=0)
    c, s = np.cos(t / eps), np.sin(t / eps)
    return np.stack([np.real(c * px + s * py), np.real(c * py - s * px)])


# ----------------------------------------------------------------------------------------------
# driver                                                    test/bupdate.jl:63-114
# ----------------------------------------------------------------------------------------------
def run_bupdate(mesh, ntau, eps, dt, nstep, x, v, w):
    """returns (x, v, energy[1+2*nstep], sumv[nstep,2]); x, v are (2,np) arrays, modified copies returned"""
    x = np.array(x, dtype=REAL, copy=True)
    v = np.array(v, dtype=REAL, copy=True)
    nx, ny = mesh.nx, mesh.ny
    rho = np.zeros((nx + 1, ny + 1), dtype=REAL)
    e = np.zeros((2, nx + 1, ny + 1), dtype=REAL)
    ep = np.zeros_like(x)
    poisson = Poisson(mesh)
    energy = []
    sumv = []
    compute_rho_m6(mesh, rho, x, w)                                              # :63
    energy.append(poisson(rho, e))                                               # :65
    interpol_eb_m6(mesh, e, x, ep)                                               # :67
    et = np.zeros((ntau, 2, x.shape[1]), dtype=REAL)
    for _ in range(nstep):
        b, t, pl, ql, xt, yt = preparation(ntau, eps, dt, x, v, ep)              # :71
        interpol_eb_m6_tau(mesh, e, xt, et)                                      # :73
        fx, fy = compute_f(eps, b, xt, yt, et)                                   # :77
        xf = np.fft.fft(xt, axis=0)                                              # :79
        xt = ua_step_predict(eps, t, pl, xf, fx)                                 # :80
        yf = np.fft.fft(yt, axis=0)                                              # :82
        yt = ua_step_predict(eps, t, pl, yf, fy)                                 # :83
        xt = np.fft.ifft(xt, axis=0)                                             # :85
        yt = np.fft.ifft(yt, axis=0)                                             # :86
        compute_rho_m6_tau(mesh, rho, x, w, xt, t, eps)                          # :88
        energy.append(poisson(rho, e))                                           # :90
        interpol_eb_m6_tau(mesh, e, xt, et)                                      # :92
        gx, gy = compute_f(eps, b, xt, yt, et)                                   # :96
        xt = ua_step_correct(eps, t, pl, ql, xf, fx, gx)                         # :98
        yt = ua_step_correct(eps, t, pl, ql, yf, fy, gy)                         # :100
        xt = np.fft.ifft(xt, axis=0)                                             # :102
        compute_rho_m6_tau(mesh, rho, x, w, xt, t, eps)                          # :104
        energy.append(poisson(rho, e))                                           # :106
        v = compute_v(eps, t, yt)                                                # :110
        sumv.append([np.sum(v[0]), np.sum(v[1])])                                # :112
    return x, v, np.array(energy), np.array(sumv), e

"""Host-side mirror of the UAPIC.jl API for the `bupdate` path (Julia is not installed in this image;
`julia/UAPIC.jl` holds the equivalent `ccall` wrappers).

Same names, argument order and mutation semantics as the reference's exported functions, minus the
trailing `!`.  Arrays are numpy arrays in Fortran (column-major) order with the reference's shapes, so
the bytes handed to the C ABI are exactly what Julia would hand over.  Every function runs on the GPU
through libuapic_b200.so; nothing is computed on the host.

    Mesh, MeshFields, Particles, UA, Poisson                 src/meshfields.jl, src/particles.jl, src/ua_type.jl, src/poisson.jl
    compute_rho_m6, interpol_eb_m6                           src/compute_rho.jl, src/interpolation.jl
    preparation, update_particles_e, update_particles_x,
    compute_f, ua_step, compute_v                            src/ua_steps.jl
    fft_tau, ifft_tau                                        mul!(x̃t, ftau, xt) / ifft!(xt, 1) of test/bupdate.jl
    integrate, errors                                        src/integrate.jl, src/gnuplot.jl:29-35
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MeshStruct, check, lib

_dp = C.POINTER(C.c_double)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.f_contiguous):
        raise TypeError("expected a Fortran-ordered float64 numpy array")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _c128(a, shape=None):
    if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags.f_contiguous):
        raise TypeError("expected a Fortran-ordered complex128 numpy array")
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


class Mesh:
    """src/meshfields.jl:3-23"""

    def __init__(self, xmin, xmax, nx, ymin, ymax, ny):
        self.xmin, self.xmax, self.nx = float(xmin), float(xmax), int(nx)
        self.ymin, self.ymax, self.ny = float(ymin), float(ymax), int(ny)
        self.dx = (self.xmax - self.xmin) / self.nx
        self.dy = (self.ymax - self.ymin) / self.ny

    def _struct(self) -> MeshStruct:
        return MeshStruct(self.xmin, self.xmax, self.ymin, self.ymax, self.nx, self.ny)


class MeshFields:
    """src/meshfields.jl:27-45 : e (2,nx+1,ny+1), rho (nx+1,ny+1) (the reference's field name is the Greek rho)"""

    def __init__(self, mesh: Mesh):
        self.mesh = mesh
        self.e = np.zeros((2, mesh.nx + 1, mesh.ny + 1), order="F")
        self.rho = np.zeros((mesh.nx + 1, mesh.ny + 1), order="F")


class Particles:
    """src/particles.jl:12-35"""

    def __init__(self, nbpart: int, w: float):
        self.nbpart = int(nbpart)
        self.x = np.zeros((2, nbpart), order="F")
        self.v = np.zeros((2, nbpart), order="F")
        self.e = np.zeros((2, nbpart), order="F")
        self.b = np.zeros(nbpart)
        self.t = np.zeros(nbpart)
        self.w = float(w)


class UA:
    """src/ua_type.jl:3-41 (the FFTW plans are replaced by the library's warp-shuffle FFT)"""

    def __init__(self, ntau: int, eps: float, nbpart: int, wrap=_lib.WRAP_JULIA, deposit_mode=_lib.DEPOSIT_FP64_ATOMIC):
        if ntau % 2 or not 2 <= ntau <= 256:
            raise ValueError("ntau must be even and in [2, 256] (powers of two <= 32 run on the fast kernels)")
        self.ntau = int(ntau)
        self.eps = float(eps)
        dtau = 2 * np.pi / ntau
        self.ltau = np.concatenate([np.arange(0, ntau // 2), np.arange(-ntau // 2, 0)]).astype(np.float64)
        self.tau = np.array([i * dtau for i in range(ntau)])
        self.pl = np.zeros((ntau, nbpart), dtype=np.complex128, order="F")
        self.ql = np.zeros((ntau, nbpart), dtype=np.complex128, order="F")
        # conventions of the library calls made with this object (Julia's by default)
        self.wrap = wrap
        self.deposit_mode = deposit_mode


class Poisson:
    """src/poisson.jl:14-83 : `poisson = Poisson(mesh); nrj = poisson(fields)`"""

    def __init__(self, mesh: Mesh):
        self.mesh = mesh

    def __call__(self, fields: MeshFields) -> float:
        m = self.mesh
        nrj = C.c_double(0.0)
        ms = m._struct()
        check(lib().uapic_poisson(C.byref(ms), _ptr(_f64(fields.rho, (m.nx + 1, m.ny + 1))),
                                  _ptr(_f64(fields.e, (2, m.nx + 1, m.ny + 1))), C.byref(nrj)))
        return nrj.value


def compute_rho_m6(fields: MeshFields, particles: Particles, xt=None, ua: UA | None = None, wrap=None, deposit_mode=None):
    """compute_rho_m6!(fields, particles)            src/compute_rho.jl:181-316
       compute_rho_m6!(fields, particles, xt, ua)    src/compute_rho.jl:29-179   (overwrites particles.x)
    returns rho_total (the value the reference prints)."""
    m = fields.mesh
    ms = m._struct()
    wrap = (ua.wrap if ua is not None else _lib.WRAP_JULIA) if wrap is None else wrap
    deposit_mode = (ua.deposit_mode if ua is not None else _lib.DEPOSIT_FP64_ATOMIC) if deposit_mode is None else deposit_mode
    tot = C.c_double(0.0)
    if xt is None:
        check(lib().uapic_compute_rho_m6(C.byref(ms), C.c_int64(particles.nbpart), _ptr(_f64(particles.x)),
                                         C.c_double(particles.w), _ptr(_f64(fields.rho)), C.c_int(wrap),
                                         C.c_int(deposit_mode), C.byref(tot)))
    else:
        if ua is None:
            raise TypeError("compute_rho_m6(fields, particles, xt, ua) needs ua")
        _c128(xt, (ua.ntau, 2, particles.nbpart))
        check(lib().uapic_compute_rho_m6_tau(C.byref(ms), C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(particles.nbpart),
                                             _ptr(xt), _ptr(particles.t), C.c_double(particles.w), _ptr(_f64(fields.rho)),
                                             _ptr(_f64(particles.x)), C.c_int(wrap), C.c_int(deposit_mode), C.byref(tot)))
    return tot.value


def interpol_eb_m6(*args, wrap=_lib.WRAP_JULIA):
    """interpol_eb_m6!(particles, fields)                 src/interpolation.jl:125-247
       interpol_eb_m6!(e, fields, x, nbpart, ntau)        src/interpolation.jl:3-123"""
    if len(args) == 2:
        particles, fields = args
        ms = fields.mesh._struct()
        check(lib().uapic_interpol_eb_m6(C.byref(ms), _ptr(_f64(fields.e)), C.c_int64(particles.nbpart),
                                         _ptr(_f64(particles.x)), _ptr(_f64(particles.e)), C.c_int(wrap)))
    elif len(args) == 5:
        e, fields, x, nbpart, ntau = args
        ms = fields.mesh._struct()
        _c128(x, (ntau, 2, nbpart))
        _f64(e, (ntau, 2, nbpart))
        check(lib().uapic_interpol_eb_m6_tau(C.byref(ms), _ptr(_f64(fields.e)), C.c_int(ntau), C.c_int64(nbpart), _ptr(x),
                                             _ptr(e), C.c_int(wrap)))
    else:
        raise TypeError("interpol_eb_m6(particles, fields) or interpol_eb_m6(e, fields, x, nbpart, ntau)")


def compute_rho_cic(fields: MeshFields, particles: Particles, wrap=_lib.WRAP_JULIA, deposit_mode=_lib.DEPOSIT_FP64_ATOMIC):
    """the plain deposit with the bilinear shape of SCHEME_CIC (build-defined, include/uapic_b200.h); returns rho_total"""
    ms = fields.mesh._struct()
    tot = C.c_double(0.0)
    check(lib().uapic_compute_rho_cic(C.byref(ms), C.c_int64(particles.nbpart), _ptr(_f64(particles.x)), C.c_double(particles.w),
                                      _ptr(_f64(fields.rho)), C.c_int(wrap), C.c_int(deposit_mode), C.byref(tot)))
    return tot.value


def interpol_eb_cic(particles: Particles, fields: MeshFields, wrap=_lib.WRAP_JULIA):
    """the plain gather with the bilinear shape of SCHEME_CIC (build-defined)"""
    ms = fields.mesh._struct()
    check(lib().uapic_interpol_eb_cic(C.byref(ms), _ptr(_f64(fields.e)), C.c_int64(particles.nbpart), _ptr(_f64(particles.x)),
                                      _ptr(_f64(particles.e)), C.c_int(wrap)))


def preparation(ua: UA, dt: float, particles: Particles, xt, yt):
    """preparation!(ua, dt, particles, xt, yt)     src/ua_steps.jl:3-78"""
    n = particles.nbpart
    _c128(xt, (ua.ntau, 2, n))
    _c128(yt, (ua.ntau, 2, n))
    check(lib().uapic_preparation(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_double(dt), C.c_int64(n), _ptr(_f64(particles.x)),
                                  _ptr(_f64(particles.v)), _ptr(_f64(particles.e)), _ptr(particles.b), _ptr(particles.t),
                                  _ptr(_c128(ua.pl, (ua.ntau, n))), _ptr(_c128(ua.ql, (ua.ntau, n))), _ptr(xt), _ptr(yt)))


def update_particles_e(particles: Particles, et, fields: MeshFields, ua: UA, xt):
    """update_particles_e!     src/ua_steps.jl:82-90"""
    interpol_eb_m6(et, fields, xt, particles.nbpart, ua.ntau, wrap=ua.wrap)


def update_particles_x(particles: Particles, fields: MeshFields, ua: UA, xt):
    """update_particles_x!     src/ua_steps.jl:94-101"""
    return compute_rho_m6(fields, particles, xt, ua)


def compute_f(fx, fy, ua: UA, particles: Particles, xt, yt, et, normalise=False):
    """compute_f!(fx, fy, ua, particles, xt, yt, et)     src/ua_steps.jl:105-145 (normalise=True: Fortran, ua_steps.F90:194-195)"""
    n = particles.nbpart
    shp = (ua.ntau, 2, n)
    check(lib().uapic_compute_f(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.b), _ptr(_c128(xt, shp)),
                                _ptr(_c128(yt, shp)), _ptr(_f64(et, shp)), _ptr(_c128(fx, shp)), _ptr(_c128(fy, shp)),
                                C.c_int(int(normalise))))


def fft_tau(out, ua_or_ntau, inp):
    """mul!(x̃t, ftau, xt)     test/bupdate.jl:79,82"""
    ntau = ua_or_ntau.ntau if isinstance(ua_or_ntau, UA) else int(ua_or_ntau)
    _c128(inp)
    _c128(out, inp.shape)
    check(lib().uapic_fft_tau(C.c_int(ntau), C.c_int64(inp.size // ntau), _ptr(inp), _ptr(out), C.c_int(-1), C.c_int(0)))


def ifft_tau(a, ua_or_ntau=None):
    """ifft!(xt, 1)     test/bupdate.jl:85-86,102 (in place, normalised)"""
    ntau = a.shape[0] if ua_or_ntau is None else (ua_or_ntau.ntau if isinstance(ua_or_ntau, UA) else int(ua_or_ntau))
    _c128(a)
    check(lib().uapic_fft_tau(C.c_int(ntau), C.c_int64(a.size // ntau), _ptr(a), _ptr(a), C.c_int(+1), C.c_int(1)))


def ua_step(xt, xft, ua: UA, particles: Particles, fx, gx=None):
    """ua_step!(xt, x̃t, ua, particles, fx)         src/ua_steps.jl:149-170
       ua_step!(xt, x̃t, ua, particles, fx, gx)     src/ua_steps.jl:172-200"""
    n = particles.nbpart
    shp = (ua.ntau, 2, n)
    if gx is None:
        check(lib().uapic_ua_step_predict(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.t), _ptr(ua.pl),
                                          _ptr(_c128(xft, shp)), _ptr(_c128(fx, shp)), _ptr(_c128(xt, shp))))
    else:
        check(lib().uapic_ua_step_correct(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.t), _ptr(ua.pl),
                                          _ptr(ua.ql), _ptr(_c128(xft, shp)), _ptr(_c128(fx, shp)), _ptr(_c128(gx, shp)),
                                          _ptr(_c128(xt, shp))))


def ua_step1(xt, xf, ua: UA, particles: Particles, fx):
    """Fortran ua_step1     ua_steps.F90:200-236"""
    n = particles.nbpart
    shp = (ua.ntau, 2, n)
    check(lib().uapic_ua_step1(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.t), _ptr(ua.pl),
                               _ptr(_c128(xt, shp)), _ptr(_c128(xf, shp)), _ptr(_c128(fx, shp))))


def ua_step2(xt, xf, ua: UA, particles: Particles, fx, gx):
    """Fortran ua_step2     ua_steps.F90:238-272"""
    n = particles.nbpart
    shp = (ua.ntau, 2, n)
    check(lib().uapic_ua_step2(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.t), _ptr(ua.pl), _ptr(ua.ql),
                               _ptr(_c128(xt, shp)), _ptr(_c128(xf, shp)), _ptr(_c128(fx, shp)), _ptr(_c128(gx, shp))))


def compute_v(yt, particles: Particles, ua: UA, yt_is_fourier=True):
    """compute_v!(yt, particles, ua)     src/ua_steps.jl:204-224 (yt in tau-Fourier space);
    yt_is_fourier=False is the Fortran form (ua_steps.F90:274-307: FFT first)."""
    n = particles.nbpart
    check(lib().uapic_compute_v(C.c_int(ua.ntau), C.c_double(ua.eps), C.c_int64(n), _ptr(particles.t),
                                _ptr(_c128(yt, (ua.ntau, 2, n))), C.c_int(int(yt_is_fourier)), _ptr(_f64(particles.v))))


def integrate(field: np.ndarray, mesh: Mesh) -> float:
    """src/integrate.jl:3-10 (diagnostic, host side)"""
    return float(np.sum(field[:mesh.nx, :mesh.ny]) * mesh.dx * mesh.dy)


def errors(computed: MeshFields, reference: MeshFields) -> float:
    """src/gnuplot.jl:29-35"""
    return float(np.max(np.abs(computed.e - reference.e)))


def gnuplot(filename: str, fields: MeshFields) -> None:
    """src/gnuplot.jl:4-27: one line `x  y  e1  e2  rho` per node, x outer / y inner, a blank line after every x column
    (host-side diagnostic dump, same record order and separators as the reference; numbers in Python's shortest repr)"""
    m = fields.mesh
    with open(filename, "w") as f:
        for i in range(m.nx + 1):
            for j in range(m.ny + 1):
                f.write(f"{i * m.dx!r}  {j * m.dy!r}  {float(fields.e[0, i, j])!r}  {float(fields.e[1, i, j])!r}  {float(fields.rho[i, j])!r}\n")
            f.write("\n")

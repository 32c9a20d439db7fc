"""CPU oracle package -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this package.  The product (``uapic.jl_b200``) never does.

``corc``  : ctypes binding of ``uapic_oracle.c`` (Fortran-following restatement)
``nporc`` : the numpy twin (Julia-following restatement)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import uapic_oracle_np as nporc  # noqa: F401
from . import efd_np  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

WRAP_FORTRAN = 0
WRAP_JULIA = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "uapic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _SO


class OrcMesh(C.Structure):
    _fields_ = [("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
                ("nx", C.c_int32), ("ny", C.c_int32)]

    @property
    def dx(self):
        return (self.xmax - self.xmin) / self.nx

    @property
    def dy(self):
        return (self.ymax - self.ymin) / self.ny


def mesh(xmin, xmax, nx, ymin, ymax, ny) -> OrcMesh:
    return OrcMesh(float(xmin), float(xmax), float(ymin), float(ymax), int(nx), int(ny))


_dp = C.POINTER(C.c_double)


def _p(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags.f_contiguous or a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


class _COracle:
    """thin ctypes front-end; arrays are Fortran-ordered numpy arrays in the reference's shapes"""

    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_f_m6.restype = C.c_double
        L.orc_f_m6.argtypes = [C.c_double]
        L.orc_compute_rho_m6.restype = C.c_double
        L.orc_compute_rho_m6_tau.restype = C.c_double
        L.orc_compute_rho_m6_fixed.restype = C.c_double
        L.orc_poisson.restype = C.c_double
        L.orc_run_bupdate.restype = C.c_int
        L.orc_plasma_from_uniforms.restype = C.c_int64
        L.orc_get_max_threads.restype = C.c_int
        L.orc_sim_create.restype = C.c_void_p
        L.orc_sim_init.restype = C.c_double

    # -- helpers -------------------------------------------------------------------------------
    def set_threads(self, n):
        self.lib.orc_set_threads(C.c_int(int(n)))

    def max_threads(self):
        return int(self.lib.orc_get_max_threads())

    def set_scheme(self, scheme):
        """0 / "m6": the reference's shape function; 1 / "cic": the build-defined bilinear variant (uapic_oracle.c)"""
        self.lib.orc_set_scheme(C.c_int({"m6": 0, "cic": 1}.get(scheme, scheme)))

    def scheme(self):
        return int(self.lib.orc_get_scheme())

    def f_m6(self, q):
        return float(self.lib.orc_f_m6(float(q)))

    def ua_tables(self, ntau):
        tau = np.zeros(ntau)
        ltau = np.zeros(ntau)
        self.lib.orc_ua_tables(C.c_int(ntau), _p(tau), _p(ltau))
        return tau, ltau

    def fft_tau(self, a, sign=-1, normalise=False):
        a = np.asfortranarray(a, dtype=np.complex128)
        out = np.empty_like(a, order="F")
        ntau = a.shape[0]
        nvec = a.size // ntau
        self.lib.orc_fft_tau(C.c_int(ntau), C.c_int64(nvec), a.ctypes.data_as(_dp), out.ctypes.data_as(_dp),
                             C.c_int(sign), C.c_int(int(normalise)))
        return out

    # -- mesh <-> particles --------------------------------------------------------------------
    def compute_rho_m6(self, m, x, w, rho, wrap=WRAP_FORTRAN):
        return self.lib.orc_compute_rho_m6(C.byref(m), C.c_int64(x.shape[1]), _p(x), C.c_double(w), _p(rho), C.c_int(wrap))

    def compute_rho_m6_fixed(self, m, x, w, scale, rho, wrap=WRAP_FORTRAN):
        return self.lib.orc_compute_rho_m6_fixed(C.byref(m), C.c_int64(x.shape[1]), _p(x), C.c_double(w), C.c_double(scale),
                                                 _p(rho), C.c_int(wrap))

    def interpol_eb_m6(self, m, e, x, ep, wrap=WRAP_FORTRAN):
        self.lib.orc_interpol_eb_m6(C.byref(m), _p(e), C.c_int64(x.shape[1]), _p(x), _p(ep), C.c_int(wrap))

    def interpol_eb_m6_tau(self, m, e, xt, et, wrap=WRAP_FORTRAN):
        ntau, _, npart = xt.shape
        self.lib.orc_interpol_eb_m6_tau(C.byref(m), _p(e), C.c_int(ntau), C.c_int64(npart), xt.ctypes.data_as(_dp), _p(et), C.c_int(wrap))

    def compute_rho_m6_tau(self, m, eps, xt, t, w, rho, x, wrap=WRAP_FORTRAN):
        ntau, _, npart = xt.shape
        return self.lib.orc_compute_rho_m6_tau(C.byref(m), C.c_int(ntau), C.c_double(eps), C.c_int64(npart),
                                               xt.ctypes.data_as(_dp), _p(t), C.c_double(w), _p(rho), _p(x), C.c_int(wrap))

    def poisson(self, m, rho, e):
        return self.lib.orc_poisson(C.byref(m), _p(rho), _p(e))

    # -- UA stages -----------------------------------------------------------------------------
    def preparation(self, ntau, eps, dt, x, v, ep):
        npart = x.shape[1]
        b = np.zeros(npart)
        t = np.zeros(npart)
        pl = np.zeros((ntau, npart), dtype=np.complex128, order="F")
        ql = np.zeros((ntau, npart), dtype=np.complex128, order="F")
        xt = np.zeros((ntau, 2, npart), dtype=np.complex128, order="F")
        yt = np.zeros((ntau, 2, npart), dtype=np.complex128, order="F")
        self.lib.orc_preparation(C.c_int(ntau), C.c_double(eps), C.c_double(dt), C.c_int64(npart), _p(x), _p(v), _p(ep),
                                 _p(b), _p(t), pl.ctypes.data_as(_dp), ql.ctypes.data_as(_dp),
                                 xt.ctypes.data_as(_dp), yt.ctypes.data_as(_dp))
        return b, t, pl, ql, xt, yt

    def compute_f(self, eps, b, xt, yt, et, normalise=True):
        ntau, _, npart = xt.shape
        fx = np.zeros_like(xt, order="F")
        fy = np.zeros_like(xt, order="F")
        self.lib.orc_compute_f(C.c_int(ntau), C.c_double(eps), C.c_int64(npart), _p(b), xt.ctypes.data_as(_dp),
                               yt.ctypes.data_as(_dp), _p(et), fx.ctypes.data_as(_dp), fy.ctypes.data_as(_dp),
                               C.c_int(int(normalise)))
        return fx, fy

    def ua_step1(self, eps, t, pl, xt, fx):
        """in place on xt; returns xf"""
        ntau, _, npart = xt.shape
        xf = np.zeros_like(xt, order="F")
        self.lib.orc_ua_step1(C.c_int(ntau), C.c_double(eps), C.c_int64(npart), _p(t), pl.ctypes.data_as(_dp),
                              xt.ctypes.data_as(_dp), xf.ctypes.data_as(_dp), fx.ctypes.data_as(_dp))
        return xf

    def ua_step2(self, eps, t, pl, ql, xt, xf, fx, gx):
        ntau, _, npart = xt.shape
        self.lib.orc_ua_step2(C.c_int(ntau), C.c_double(eps), C.c_int64(npart), _p(t), pl.ctypes.data_as(_dp),
                              ql.ctypes.data_as(_dp), xt.ctypes.data_as(_dp), xf.ctypes.data_as(_dp),
                              fx.ctypes.data_as(_dp), gx.ctypes.data_as(_dp))

    def compute_v(self, eps, t, yt, v):
        ntau, _, npart = yt.shape
        yf = np.zeros_like(yt, order="F")
        self.lib.orc_compute_v(C.c_int(ntau), C.c_double(eps), C.c_int64(npart), _p(t), yt.ctypes.data_as(_dp),
                               yf.ctypes.data_as(_dp), _p(v))
        return yf

    # -- driver --------------------------------------------------------------------------------
    def run_bupdate(self, m, ntau, eps, dt, nstep, x, v, w, wrap=WRAP_FORTRAN, faithful=False):
        """x, v: (2,np) Fortran-ordered float64, updated in place.  returns (energy, sumv, e_part, e_mesh)"""
        npart = x.shape[1]
        energy = np.zeros(1 + 2 * nstep)
        sumv = np.zeros((nstep, 2))
        ep = np.zeros((2, npart), order="F")
        emesh = np.zeros((2, m.nx + 1, m.ny + 1), order="F")
        rc = self.lib.orc_run_bupdate(C.byref(m), C.c_int(ntau), C.c_double(eps), C.c_double(dt), C.c_int(nstep),
                                      C.c_int64(npart), C.c_double(w), _p(x), _p(v), _p(ep), _p(emesh), _p(energy),
                                      _p(sumv), C.c_int(wrap), C.c_int(int(faithful)))
        if rc != 0:
            raise RuntimeError(f"orc_run_bupdate failed rc={rc}")
        return energy, sumv, ep, emesh

    def sim(self, m, ntau, eps, dt, x, v, w, wrap=WRAP_FORTRAN, faithful=True):
        """persistent-state driver (init / step separately timed by bench.py's CPU baseline)"""
        return _Sim(self, m, ntau, eps, dt, x, v, w, wrap, faithful)

    def generate(self, m, kind, seed, npart, np_global=None, first=0, stride=1, alpha=0.05, kx=0.5):
        """the counter-based synthetic loads of the benchmark configs, same particles as the device generator
        (kind "plasma"/0: particles.F90:68-103 densities; "landau"/1: intent of src/landau.jl:19-43)"""
        kind = {"plasma": 0, "landau": 1}.get(kind, kind)
        x = np.zeros((2, npart), order="F")
        v = np.zeros((2, npart), order="F")
        self.lib.orc_generate(C.byref(m), C.c_int(kind), C.c_uint64(seed), C.c_int64(first), C.c_int64(stride), C.c_int64(npart),
                              C.c_int64(np_global if np_global is not None else npart), C.c_double(alpha), C.c_double(kx), _p(x), _p(v))
        return x, v

    def efd_run(self, x, v, ntau=16, eps=1e-3, dt=np.pi / 16, tfinal=np.pi / 2, box=(0.0, 4 * np.pi, 0.0, 2 * np.pi)):
        """fortran/efd.f90 over all particles (the text behind its `stop`); returns (x, v) at tfinal, inputs untouched"""
        xo, vo = np.array(x, order="F", dtype=np.float64), np.array(v, order="F", dtype=np.float64)
        b = np.array(box, dtype=np.float64)
        n = self.lib.orc_efd_run(C.c_int(ntau), C.c_int64(xo.shape[1]), C.c_double(eps), C.c_double(dt), C.c_double(tfinal), _p(b), _p(xo), _p(vo))
        if n < 0:
            raise ValueError("ntau must be even, 2..256")
        return xo, vo

    def plasma_from_uniforms(self, m, npart, alpha, kx, u):
        x = np.zeros((2, npart), order="F")
        v = np.zeros((2, npart), order="F")
        used = self.lib.orc_plasma_from_uniforms(C.byref(m), C.c_int64(npart), C.c_double(alpha), C.c_double(kx),
                                                 _p(u), C.c_int64(u.size), _p(x), _p(v))
        if used < 0:
            raise RuntimeError("not enough uniform deviates")
        return x, v, int(used)


class _Sim:
    def __init__(self, orc, m, ntau, eps, dt, x, v, w, wrap, faithful):
        self.orc, self.x, self.v = orc, x, v
        self.h = orc.lib.orc_sim_create(C.byref(m), C.c_int(ntau), C.c_double(eps), C.c_double(dt), C.c_int64(x.shape[1]),
                                        C.c_double(w), _p(x), _p(v), C.c_int(wrap), C.c_int(int(faithful)))
        if not self.h:
            raise MemoryError("orc_sim_create failed")

    def init(self):
        return float(self.orc.lib.orc_sim_init(C.c_void_p(self.h)))

    def step(self):
        e2 = np.zeros(2)
        sv = np.zeros(2)
        self.orc.lib.orc_sim_step(C.c_void_p(self.h), _p(e2), _p(sv))
        return e2, sv

    def close(self):
        if self.h:
            self.orc.lib.orc_sim_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()


_corc = None


def corc() -> _COracle:
    global _corc
    if _corc is None:
        _corc = _COracle()
    return _corc


# ---- sibling scheme: the 3D rotation-push PIC of fortran/uapic3d.f90 (uapic_oracle.c, orc3_*) ---------------------------
class Orc3Mesh(C.Structure):
    _fields_ = [("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("n", C.c_int32 * 3)]


def mesh3(xmin, xmax, n) -> Orc3Mesh:
    return Orc3Mesh((C.c_double * 3)(*map(float, xmin)), (C.c_double * 3)(*map(float, xmax)), (C.c_int32 * 3)(*map(int, n)))


class _COracle3:
    """arrays: x, v, e_part (3, np); rho (nx+1, ny+1, nz+1); e (3, nx+1, ny+1, nz+1), Fortran order"""

    def __init__(self):
        self.lib = corc().lib
        self.lib.orc3_run.restype = C.c_int64

    def compute_rho_cic(self, m, x, w):
        rho = np.zeros((m.n[0] + 1, m.n[1] + 1, m.n[2] + 1), order="F")
        self.lib.orc3_compute_rho_cic(C.byref(m), C.c_int64(x.shape[1]), _p(x), C.c_double(w), _p(rho))
        return rho

    def poisson(self, m, rho):
        e = np.zeros((3, m.n[0] + 1, m.n[1] + 1, m.n[2] + 1), order="F")
        self.lib.orc3_poisson(C.byref(m), _p(rho), _p(e))
        return e

    def interpolate_eb_cic(self, m, e, x):
        ep = np.zeros((3, x.shape[1]), order="F")
        self.lib.orc3_interpolate_eb_cic(C.byref(m), _p(e), C.c_int64(x.shape[1]), _p(x), _p(ep))
        return ep

    def generate(self, m, seed, npart, first=0):
        x = np.zeros((3, npart), order="F")
        v = np.zeros((3, npart), order="F")
        self.lib.orc3_generate(C.byref(m), C.c_uint64(seed), C.c_int64(first), C.c_int64(npart), _p(x), _p(v))
        return x, v

    def run(self, m, x, v, w, ep, delta, nmrc, nmrcm, tfinal, max_outer=0, index_quirk=1):
        """x, v updated in place; returns (substeps, e_part, e, rho)"""
        npart = x.shape[1]
        epart = np.zeros((3, npart), order="F")
        e = np.zeros((3, m.n[0] + 1, m.n[1] + 1, m.n[2] + 1), order="F")
        rho = np.zeros((m.n[0] + 1, m.n[1] + 1, m.n[2] + 1), order="F")
        n = self.lib.orc3_run(C.byref(m), C.c_int64(npart), _p(x), _p(v), _p(epart), C.c_double(w), C.c_double(ep), C.c_double(delta),
                              C.c_int(nmrc), C.c_int(nmrcm), C.c_double(tfinal), C.c_int(max_outer), C.c_int(index_quirk), _p(e), _p(rho))
        return int(n), epart, e, rho


_corc3 = None


def corc3() -> _COracle3:
    global _corc3
    if _corc3 is None:
        _corc3 = _COracle3()
    return _corc3

"""Host-side mirror of the reference's external-field program (fortran/efd.f90, test/test_efd.jl) over `uapic_efd_run`.

``efd_run(x, v, ...)`` integrates every particle in the prescribed field of efd.f90:166-167 (one CUDA kernel, no mesh in the
loop); ``efd(ntau, nbpart)`` is the whole program with the reference's own constants -- load, run, the two numbers it prints
(efd.f90:481), and the final M6 deposit (efd.f90:484).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import WRAP_FORTRAN, check, lib
from .api import Mesh, MeshFields, Particles, compute_rho_m6

_dp = C.POINTER(C.c_double)

# the constants the reference's program prints its sum(v) against (efd.f90:481, test/test_efd.jl:468)
REFERENCE_SUM_V = (-857.95049281063064, -593.40700170710875)


class EfdConfigStruct(C.Structure):
    _fields_ = [("ntau", C.c_int32), ("nstep", C.c_int32), ("eps", C.c_double), ("dt", C.c_double), ("tfinal", C.c_double),
                ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double)]


def _config(ntau, eps, dt, tfinal, box, nstep):
    return EfdConfigStruct(int(ntau), int(nstep), float(eps), float(dt), float(tfinal), *map(float, box))


def efd_run(x, v, ntau=16, eps=1e-3, dt=np.pi / 16, tfinal=np.pi / 2, box=(0.0, 4 * np.pi, 0.0, 2 * np.pi), nstep=0):
    """x, v: (2, nbpart) float64.  Returns (x, v) at tfinal (efd.f90:133-478 over all particles); the inputs are untouched.
    nstep = 0 takes nint(tfinal/dt) (efd.f90:102)."""
    x = np.asfortranarray(x, dtype=np.float64)
    v = np.asfortranarray(v, dtype=np.float64)
    if x.ndim != 2 or x.shape[0] != 2 or v.shape != x.shape:
        raise ValueError("x and v must both have shape (2, nbpart)")
    xo, vo = np.empty_like(x, order="F"), np.empty_like(v, order="F")
    cfg = _config(ntau, eps, dt, tfinal, box, nstep)
    check(lib().uapic_efd_run(C.byref(cfg), C.c_int64(x.shape[1]), x.ctypes.data_as(_dp), v.ctypes.data_as(_dp),
                              xo.ctypes.data_as(_dp), vo.ctypes.data_as(_dp)))
    return xo, vo


def efd_run_device(x, v, x_out=None, v_out=None, ntau=16, eps=1e-3, dt=np.pi / 16, tfinal=np.pi / 2,
                   box=(0.0, 4 * np.pi, 0.0, 2 * np.pi), nstep=0, stream=None):
    """the same on torch CUDA tensors of shape (nbpart, 2) (= column-major (2, nbpart)); asynchronous on `stream`
    (default: torch's current stream).  Returns (x_out, v_out); pass x_out = x, v_out = v to update in place."""
    import torch
    for t in (x, v):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.ndim == 2 and t.shape[1] == 2):
            raise ValueError("expected contiguous float64 CUDA tensors of shape (nbpart, 2)")
    x_out = torch.empty_like(x) if x_out is None else x_out
    v_out = torch.empty_like(v) if v_out is None else v_out
    st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else int(stream)
    cfg = _config(ntau, eps, dt, tfinal, box, nstep)
    with torch.cuda.device(x.device):
        check(lib().uapic_efd_run_device(C.byref(cfg), C.c_int64(x.shape[0]), C.c_void_p(x.data_ptr()), C.c_void_p(v.data_ptr()),
                                         C.c_void_p(x_out.data_ptr()), C.c_void_p(v_out.data_ptr()), C.c_void_p(st)))
    return x_out, v_out


def efd(ntau: int = 16, nbpart: int | None = None, particles: Particles | None = None, eps: float = 1e-3):
    """`efd(ntau, nbpart)` of test/test_efd.jl:8 / `program efd`: 128 x 64 mesh on [0,4pi] x [0,2pi], dt = pi/16, tfinal = pi/2,
    the first `nbpart` particles of the load integrated (all of them by default), then compute_rho_m6.  `particles` defaults
    to init_particles_2d's load (`plasma(..., use_gfortran=True)`: the reference's own stream when libgfortran is loadable).
    Returns (particles, fields, (sum(v1) - ref1, sum(v2) - ref2)) -- the pair the program prints."""
    from .loaders import plasma
    kx, ky = 0.5, 1.0
    mesh = Mesh(0.0, 2 * np.pi / kx, 128, 0.0, 2 * np.pi / ky, 64)                  # efd.f90:64-65,90-98,117-122
    p = plasma(mesh, 204800, use_gfortran=True) if particles is None else particles
    n = p.nbpart if nbpart is None else int(nbpart)
    if not 0 <= n <= p.nbpart:
        raise ValueError("nbpart exceeds the load")
    xo, vo = efd_run(p.x[:, :n], p.v[:, :n], ntau=ntau, eps=eps, dt=np.pi / 2 / 2.0 ** 3, tfinal=np.pi / 2,
                     box=(mesh.xmin, mesh.xmax, mesh.ymin, mesh.ymax))
    p.x[:, :n], p.v[:, :n] = xo, vo
    f = MeshFields(mesh)
    compute_rho_m6(f, p, wrap=WRAP_FORTRAN)                                         # efd.f90:484
    return p, f, (p.v[0].sum() - REFERENCE_SUM_V[0], p.v[1].sum() - REFERENCE_SUM_V[1])

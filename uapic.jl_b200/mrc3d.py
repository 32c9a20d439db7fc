"""Host-side mirror of the reference's 3D path (fortran/uapic3d.f90 and its modules) over the uapic3d_* C ABI.

Same names and argument meaning as the Fortran routines: ``compute_rho_cic(fields, particles)``, ``solve_poisson(fields)``,
``interpolate_eb_cic(particles, fields)``, and ``Session3D`` / ``run_uapic3d`` for the whole program with the state in HBM.
Arrays are Fortran-ordered: x, v, e (3, nbpart); rho (nx+1, ny+1, nz+1); e (3, nx+1, ny+1, nz+1).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib

_dp = C.POINTER(C.c_double)


class Mesh3Struct(C.Structure):
    _fields_ = [("xmin", C.c_double * 3), ("xmax", C.c_double * 3), ("n", C.c_int32 * 3)]


class Config3Struct(C.Structure):
    _fields_ = [("mesh", Mesh3Struct), ("nbpart", C.c_int64), ("nbpart_global", C.c_int64), ("weight", C.c_double), ("ep", C.c_double),
                ("delta", C.c_double), ("deposit_mode", C.c_int32), ("index_quirk", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p)]


class Mesh3D:
    """init_mesh (3D), meshfields.F90:80-107"""

    def __init__(self, xmin, xmax, n):
        self.xmin, self.xmax, self.n = tuple(map(float, xmin)), tuple(map(float, xmax)), tuple(map(int, n))
        self.d = tuple((b - a) / k for a, b, k in zip(self.xmin, self.xmax, self.n))

    def _struct(self):
        return Mesh3Struct((C.c_double * 3)(*self.xmin), (C.c_double * 3)(*self.xmax), (C.c_int32 * 3)(*self.n))

    @property
    def node_shape(self):
        return tuple(k + 1 for k in self.n)


class Fields3D:
    """fields_3d_t, meshfields.F90:30-36"""

    def __init__(self, mesh: Mesh3D):
        self.mesh = mesh
        self.e = np.zeros((3,) + mesh.node_shape, order="F")
        self.rho = np.zeros(mesh.node_shape, order="F")


def _f(a, shape):
    if a.dtype != np.float64 or a.shape != shape or not a.flags.f_contiguous:
        raise ValueError(f"expected a Fortran-ordered float64 array of shape {shape}")
    return a.ctypes.data_as(_dp)


def compute_rho_cic(fields: Fields3D, x: np.ndarray, w: float):
    """compute_rho_cic(f, p)           compute_rho_cic.f90:11-79"""
    ms = fields.mesh._struct()
    check(lib().uapic3d_compute_rho_cic(C.byref(ms), C.c_int64(x.shape[1]), _f(x, (3, x.shape[1])), C.c_double(w), _f(fields.rho, fields.mesh.node_shape)))


def solve_poisson(fields: Fields3D):
    """solve_poisson(poisson, fields)  poisson_3d.f90:47-191"""
    ms = fields.mesh._struct()
    check(lib().uapic3d_poisson(C.byref(ms), _f(fields.rho, fields.mesh.node_shape), _f(fields.e, (3,) + fields.mesh.node_shape)))


def interpolate_eb_cic(x: np.ndarray, fields: Fields3D) -> np.ndarray:
    """interpolate_eb_cic(p, f)        interpolation_cic.f90:10-66 ; returns p%e (3, nbpart)"""
    ms = fields.mesh._struct()
    ep = np.zeros((3, x.shape[1]), order="F")
    check(lib().uapic3d_interpolate_eb_cic(C.byref(ms), _f(fields.e, (3,) + fields.mesh.node_shape), C.c_int64(x.shape[1]), _f(x, (3, x.shape[1])),
                                           _f(ep, ep.shape)))
    return ep


def write_data(istep: int, fields: Fields3D, directory: str = ".") -> str:
    """write_data(istep, fields)          output.f90:10-110 -- the field dump the three 3D programs make (uapic3d.f90:82,
    test_pic_3d.f90:56, test_poisson_3d.f90:48): `fields-NNNN.xmf`, an XDMF 2.2 description of a 3DCoRectMesh with node
    attributes ex, ey, ez, rho.  The reference keeps the heavy data in `fields-NNNN.h5`; no HDF5 library exists in this image,
    so the arrays go to `fields-NNNN.bin` and the DataItems say Format='Binary' with a byte offset (little-endian doubles,
    x fastest, exactly the bytes the HDF5 datasets would hold) -- ParaView / VisIt read both forms.  Host-side diagnostic;
    returns the path of the .xmf file."""
    import os
    if not 0 <= istep < 10 ** 4:                          # int2string, output.f90:113-133
        raise ValueError("istep must be in 0..9999")
    name = f"fields-{istep:04d}"
    m = fields.mesh
    nx1, ny1, nz1 = m.node_shape
    arrays = [("ex", fields.e[0]), ("ey", fields.e[1]), ("ez", fields.e[2]), ("rho", fields.rho)]
    nbytes = 8 * nx1 * ny1 * nz1
    with open(os.path.join(directory, name + ".bin"), "wb") as f:
        for _, a in arrays:
            f.write(np.asfortranarray(a, dtype="<f8").tobytes(order="F"))
    dims = f"{nz1:5d}{ny1:5d}{nx1:5d}"
    lines = ["<?xml version='1.0' ?>", "<!DOCTYPE Xdmf SYSTEM 'Xdmf.dtd' []>",
             "<Xdmf xmlns:xi='http://www.w3.org/2003/XInclude' Version='2.2'>", "<Domain>", "<Grid Name='mesh' GridType='Uniform'>",
             f"<Topology TopologyType='3DCoRectMesh' NumberOfElements='{dims}'/>", "<Geometry GeometryType='ORIGIN_DXDYDZ'>",
             "<DataItem Dimensions='3' NumberType='Float' Format='XML'>", "".join(f"{c:12.5f}" for c in m.xmin), "</DataItem>",
             "<DataItem Dimensions='3' NumberType='Float' Format='XML'>", "".join(f"{c:12.5f}" for c in m.d), "</DataItem>",
             "</Geometry>"]
    for k, (attr, _) in enumerate(arrays):
        lines += [f"<Attribute Name='{attr}' AttributeType='Scalar' Center='Node'>",
                  f"<DataItem Dimensions='{dims}' NumberType='Float' Precision='8' Format='Binary' Endian='Little' Seek='{k * nbytes}'>",
                  name + ".bin", "</DataItem>", "</Attribute>"]
    lines += ["</Grid>", "</Domain>", "</Xdmf>"]
    path = os.path.join(directory, name + ".xmf")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


class Session3D:
    """device-resident state of fortran/uapic3d.f90"""

    def __init__(self, mesh: Mesh3D, nbpart: int, ep: float = 0.5 ** 10, delta: float = 3e-3, weight: float | None = None,
                 nbpart_global: int | None = None, deposit_mode=_lib.DEPOSIT_FP64_ATOMIC, index_quirk: bool = True, device: int = 0,
                 stream: int | None = None):
        self.mesh, self.nbpart = mesh, int(nbpart)
        npg = int(nbpart_global if nbpart_global is not None else nbpart)
        vol = np.prod([b - a for a, b in zip(mesh.xmin, mesh.xmax)])
        self.weight = float(weight) if weight is not None else float(vol) / npg            # particles.F90:133
        cfg = Config3Struct()
        cfg.mesh = mesh._struct()
        cfg.nbpart, cfg.nbpart_global, cfg.weight, cfg.ep, cfg.delta = self.nbpart, npg, self.weight, float(ep), float(delta)
        cfg.deposit_mode, cfg.index_quirk, cfg.device, cfg.stream = deposit_mode, int(index_quirk), int(device), C.c_void_p(stream or 0)
        self._h = C.c_void_p()
        check(lib().uapic3d_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().uapic3d_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def init_nccl(self, unique_id: bytes, nranks: int, rank: int):
        """collective: the library sums the raw rho nodes over the ranks itself (uapic3d_init_nccl)"""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(lib().uapic3d_init_nccl(self._h, buf, C.c_int(nranks), C.c_int(rank)))

    def upload_particles(self, x, v):
        check(lib().uapic3d_upload_particles(self._h, _f(x, (3, self.nbpart)), _f(v, (3, self.nbpart))))

    def generate_particles(self, seed: int = 20190101, first_global_index: int = 0):
        check(lib().uapic3d_generate_particles(self._h, C.c_uint64(seed), C.c_int64(first_global_index)))

    def init_fields(self):
        check(lib().uapic3d_init_fields(self._h))

    def substep(self, kind: int, dt: float, coef: float = 1.0, count: int = 1):
        check(lib().uapic3d_substep(self._h, C.c_int(kind), C.c_double(dt), C.c_double(coef), C.c_int(count)))

    def run(self, nmrc: int = 2 ** 7, nmrcm: int = 2 ** 7, tfinal: float = np.pi, max_outer: int = 0) -> int:
        n = C.c_int64(0)
        check(lib().uapic3d_run(self._h, C.c_int(nmrc), C.c_int(nmrcm), C.c_double(tfinal), C.c_int(max_outer), C.byref(n)))
        return n.value

    def download_particles(self):
        x, v, ep = (np.zeros((3, self.nbpart), order="F") for _ in range(3))
        check(lib().uapic3d_download_particles(self._h, _f(x, x.shape), _f(v, v.shape), _f(ep, ep.shape)))
        return x, v, ep

    def download_fields(self):
        f = Fields3D(self.mesh)
        check(lib().uapic3d_download_fields(self._h, _f(f.e, f.e.shape), _f(f.rho, f.rho.shape)))
        return f

    @property
    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(lib().uapic3d_launch_count(self._h, C.byref(n)))
        return n.value


def run_uapic3d(mesh: Mesh3D, x, v, ep=0.5 ** 10, delta=3e-3, nmrc=2 ** 7, nmrcm=2 ** 7, tfinal=np.pi, max_outer=0, weight=None, **kw):
    """the program fortran/uapic3d.f90 on one GPU: returns (x, v, e_particles, fields, substeps)"""
    with Session3D(mesh, x.shape[1], ep=ep, delta=delta, weight=weight, **kw) as s:
        s.upload_particles(np.asfortranarray(x), np.asfortranarray(v))
        s.init_fields()
        n = s.run(nmrc, nmrcm, tfinal, max_outer)
        xo, vo, eo = s.download_particles()
        return xo, vo, eo, s.download_fields(), n

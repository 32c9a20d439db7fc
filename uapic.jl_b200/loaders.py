"""Particle loads and the `particles.dat` text format (host side; the loop never touches these).

    read_particles / write_particles     src/read_particles.jl:3-35  (one line per particle: ix iy dpx dpy vx vy)
    plasma                               src/plasma.jl:3-52, fortran/particles.F90:21-105 (rejection sampling)
    plasma3d                             fortran/particles.F90:107-190 (init_particles_3d, the 3D programs' load)
    landau_sampling                      the intent of src/landau.jl:5-47 (that file references undefined names)

`test/particles.dat` is missing from the reference checkout (.MISSING_LARGE_BLOBS), so `make_particles_dat`
regenerates a 204 800-line file with the Fortran program's densities; when the bundled libgfortran is
loadable its RNG is driven with the reference's 33-word seed (particles.F90:54-66) so the draw sequence is
the one `init_particles_2d` would make, otherwise a seeded numpy stream is used.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

from .api import Mesh, Particles

# fortran/particles.F90:57-64
FORTRAN_SEED = [-1584649339, -1457681104, 1579121008, -819547200, 249798090, -517237887, 177452147, -981503238,
                1418301473, 1989625004, 2065424384, -296364178, 1658790794, -435188152, -1643185032, 1461389312,
                1869073641, 1321930686, 483734018, 1269936416, -1999561453, 906251506, 782514880, 428753705,
                -2031262823, 263953581, 1026600222, -1118515860, 1633712916, -464192498, -1860714528, 1436611533, 0]


def read_particles(filename: str, mesh: Mesh) -> Particles:
    """src/read_particles.jl:3-35"""
    data = np.loadtxt(filename, dtype=np.float64, ndmin=2)
    nbpart = data.shape[0]
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    p = Particles(nbpart, (dimx * dimy) / nbpart)
    ix, iy, dpx, dpy = data[:, 0], data[:, 1], data[:, 2], data[:, 3]
    p.v[0], p.v[1] = data[:, 4], data[:, 5]
    p.x[0] = (dpx + ix) * mesh.dx
    p.x[1] = (dpy + iy) * mesh.dy
    return p


def write_particles(filename: str, mesh: Mesh, x: np.ndarray, v: np.ndarray) -> None:
    """inverse of read_particles: `ix iy dpx dpy vx vy` with 17 significant digits"""
    px, py = (x[0] - 0.0) / mesh.dx, (x[1] - 0.0) / mesh.dy
    ix, iy = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
    with open(filename, "w") as f:
        for k in range(x.shape[1]):
            f.write(f"{ix[k]:d} {iy[k]:d} {px[k] - ix[k]:.17g} {py[k] - iy[k]:.17g} {v[0, k]:.17g} {v[1, k]:.17g}\n")


class _GfDesc(C.Structure):
    """gfortran array descriptor of a rank-1 array (GCC >= 8 layout)"""
    _fields_ = [("base", C.c_void_p), ("offset", C.c_size_t), ("elem_len", C.c_size_t), ("version", C.c_int),
                ("rank", C.c_int8), ("type", C.c_int8), ("attr", C.c_int16), ("span", C.c_ssize_t),
                ("stride", C.c_ssize_t), ("lbound", C.c_ssize_t), ("ubound", C.c_ssize_t)]


class _Uniforms:
    """stream of U[0,1) deviates: gfortran's random_number with the reference seed, or a numpy generator.
    `take(n)` hands out the next n deviates, `give_back(k)` returns the last k unused ones to the stream, so a
    vectorised rejection loop consumes exactly the deviates the one-trial-at-a-time Fortran loop does."""

    def __init__(self, seed=None, use_gfortran=False):
        self.gf = None
        self._buf, self._pos = np.empty(0), 0
        if use_gfortran:
            import scipy
            cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libgfortran*.so*"))
            for c in sorted(cands):
                try:
                    lib = C.CDLL(c)
                    arr = (C.c_int32 * 33)(*FORTRAN_SEED)
                    # random_seed(put = seed): _gfortran_random_seed_i4(size, put, get), put a rank-1 int32 array (type 1)
                    d = _GfDesc(C.addressof(arr), C.c_size_t(-1 & (2 ** 64 - 1)), 4, 0, 1, 1, 0, 4, 1, 1, 33)
                    lib._gfortran_random_seed_i4(None, C.byref(d), None)
                    lib._gfortran_arandom_r8.argtypes = [C.POINTER(_GfDesc)]
                    lib._gfortran_arandom_r8.restype = None
                    self.gf = lib
                    break
                except (OSError, AttributeError):
                    continue
        self.rng = np.random.default_rng(20190101 if seed is None else seed)

    @property
    def source(self):
        return "libgfortran random_number, seed of particles.F90:57-64" if self.gf else "numpy PCG64"

    def _fresh(self, n):
        if self.gf is None:
            return self.rng.random(n)
        out = np.empty(n)
        # random_number on a real(8) rank-1 array (type 3) draws the same sequence as n scalar calls
        d = _GfDesc(out.ctypes.data, C.c_size_t(-1 & (2 ** 64 - 1)), 8, 0, 1, 3, 0, 8, 1, 1, n)
        self.gf._gfortran_arandom_r8(C.byref(d))
        return out

    def take(self, n):
        left = self._buf.size - self._pos
        if left < n:
            self._buf = np.concatenate([self._buf[self._pos:], self._fresh(max(n - left, 1 << 20))])
            self._pos = 0
        out = self._buf[self._pos:self._pos + n]
        self._pos += n
        return out

    def give_back(self, k):
        self._pos -= k

    def draw(self, n):
        return self.take(n).copy()


def plasma(mesh: Mesh, nbpart: int, seed=None, alpha=0.05, kx=0.5, use_gfortran=False, return_source=False):
    """src/plasma.jl:3-52 / fortran/particles.F90:68-103: x from 1+sin(y)+alpha*cos(kx*x) (bound 2+alpha),
    v from the two-bump Maxwellian on [-5,5]^2.  Three deviates per trial (xi, yi, zi), consumed strictly in the
    reference's order: the position loop stops at its nbpart-th accepted trial and the velocity loop starts at the
    very next deviate.  With `use_gfortran` the stream is libgfortran's under the seed of particles.F90:57-64, and
    the load is the one the reference's Fortran programs run on: the Fortran external-field program's printed
    sum(v) after its run (efd.f90:481) is reproduced from it to 13 digits (tests/test_efd_oracle.py)."""
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    p = Particles(nbpart, (dimx * dimy) / nbpart)
    u = _Uniforms(seed, use_gfortran)

    def fill(trial, rate, target):
        k = 0
        while k < nbpart:
            n = max(1024, int((nbpart - k) * 1.1 / rate))
            d = u.take(3 * n).reshape(n, 3)
            a, b, ok = trial(d)
            idx = np.flatnonzero(ok)
            if idx.size > nbpart - k:           # the loop ends at this trial: the later ones were never drawn
                idx = idx[:nbpart - k]
                u.give_back(3 * (n - 1 - int(idx[-1])))
            target[0, k:k + idx.size], target[1, k:k + idx.size] = a[idx], b[idx]
            k += idx.size

    def trial_x(d):                              # particles.F90:68-83
        xi, yi, zi = d[:, 0] * dimx, d[:, 1] * dimy, (2.0 + alpha) * d[:, 2]
        return xi, yi, (1.0 + np.sin(yi) + alpha * np.cos(kx * xi)) >= zi

    def trial_v(d):                              # particles.F90:85-103
        xi, yi, zi = (d[:, 0] - 0.5) * 10.0, (d[:, 1] - 0.5) * 10.0, d[:, 2]
        temm = (np.exp(-((xi - 2.0) ** 2 + yi ** 2) / 2.0) + np.exp(-((xi + 2.0) ** 2 + yi ** 2) / 2.0)) / 2.0
        return xi, yi, temm >= zi

    fill(trial_x, 1.0 / (2.0 + alpha), p.x)
    fill(trial_v, 2 * np.pi / 100.0, p.v)
    if return_source:
        return p, u.source
    return p


def plasma3d(xmin, xmax, nbpart: int, seed=None, use_gfortran=False, return_source=False):
    """init_particles_3d, fortran/particles.F90:107-190 (the load of uapic3d.f90 and test_pic_3d.f90): `random_number` on the
    whole (3, nbpart) position array first (z keeps it, scaled into the box), then x, y by rejection from the ring density
    (1 + 0.02 cos 4 theta) exp(-5 (r - 4.8)^2) around (9, 9) -- three deviates per trial -- then v from exp(-2 |v|^2) on
    [-4, 4]^3 -- four deviates per trial (0.4 % accepted).  Deviates are consumed strictly in the reference's order; with
    `use_gfortran` the stream is libgfortran's under the seed of particles.F90:142-151, i.e. the Fortran program's own particles.
    Returns x, v as Fortran-ordered (3, nbpart) arrays."""
    u = _Uniforms(seed, use_gfortran)
    x = np.zeros((3, nbpart), order="F")
    v = np.zeros((3, nbpart), order="F")
    x[:] = u.take(3 * nbpart).reshape((3, nbpart), order="F")                      # :153
    x[2] = xmin[2] + (xmax[2] - xmin[2]) * x[2]                                    # :154

    def fill(trial, width, rate, rows, target):
        k = 0
        while k < nbpart:
            n = min(max(1024, int((nbpart - k) * 1.1 / rate)), 1 << 22)
            d = u.take(width * n).reshape(n, width)
            vals, ok = trial(d)
            idx = np.flatnonzero(ok)
            if idx.size > nbpart - k:           # the loop ends at this trial: the later ones were never drawn
                idx = idx[:nbpart - k]
                u.give_back(width * (n - 1 - int(idx[-1])))
            for r, a in zip(rows, vals):
                target[r, k:k + idx.size] = a[idx]
            k += idx.size

    def trial_x(d):                              # :156-170
        xi, yi, zi = 9.0 * d[:, 0], 2.0 * np.pi * d[:, 1], (1.0 + 0.02) * d[:, 2]
        temm = (1.0 + 0.02 * np.cos(4.0 * yi)) * np.exp(-5.0 * (xi - 4.8) ** 2)
        return (np.cos(yi) * xi + 9.0, np.sin(yi) * xi + 9.0), temm >= zi

    def trial_v(d):                              # :172-188
        xi, yi, wi, zi = (d[:, 0] - 0.5) * 8.0, (d[:, 1] - 0.5) * 8.0, (d[:, 2] - 0.5) * 8.0, d[:, 3]
        return (xi, yi, wi), np.exp(-2.0 * (xi ** 2 + yi ** 2 + wi ** 2)) >= zi

    fill(trial_x, 3, np.sqrt(np.pi / 5.0) / 9.0 / 1.02, (0, 1), x)
    fill(trial_v, 4, (np.pi / 2.0) ** 1.5 / 512.0, (0, 1, 2), v)
    if return_source:
        return x, v, u.source
    return x, v


def landau_sampling(mesh: Mesh, nbpart: int, seed=20190102, alpha=0.05, kx=0.5):
    """the load src/landau.jl:19-43 describes: x1 by inverse CDF of 1+alpha*cos(kx*x) (Newton, tol 1e-12), x2 uniform,
    |v| = sqrt(-2 ln((i-0.5)/nbpart)), theta uniform.  A seeded pseudo-random stream stands in for the Sobol sequence."""
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    p = Particles(nbpart, (dimx * dimy) / nbpart)
    rng = np.random.default_rng(seed)
    r = rng.random((nbpart, 3))
    target = r[:, 1] * 2 * np.pi / kx
    x0 = target.copy()
    for _ in range(50):
        pf = x0 + alpha * np.sin(kx * x0) / kx
        f = 1 + alpha * np.cos(kx * x0)
        xn = x0 - (pf - target) / f
        done = np.max(np.abs(xn - x0)) <= 1e-12
        x0 = xn
        if done:
            break
    vv = np.sqrt(-2 * np.log((np.arange(1, nbpart + 1) - 0.5) / nbpart))
    th = r[:, 0] * 2 * np.pi
    p.x[0], p.x[1] = mesh.xmin + x0, mesh.ymin + r[:, 2] * dimy
    p.v[0], p.v[1] = vv * np.cos(th), vv * np.sin(th)
    return p


def make_particles_dat(filename: str, nbpart=204800, use_gfortran=True):
    """regenerate the missing test/particles.dat fixture (mesh of test/test_particles.jl:7: 128 x 64 on [0,4pi]x[0,2pi])"""
    mesh = Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)
    p, src = plasma(mesh, nbpart, use_gfortran=use_gfortran, return_source=True)
    write_particles(filename, mesh, p.x, p.v)
    return src

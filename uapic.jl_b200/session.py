"""Device-resident session: the whole `bupdate` loop with state kept in HBM (the performance path).

One Session per GPU / per process.  For several GPUs each rank owns a contiguous slice of the particle
index space; the only exchange is the sum of the raw rho mesh, done by `torch.distributed.all_reduce`
(NCCL over NVLink) through the C ABI's allreduce hook (see dist.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ALLREDUCE_FN, ConfigStruct, check, lib
from .api import Mesh

_dp = C.POINTER(C.c_double)


# storage_mode=None picks the fastest kernels for the given ntau: the one-pass kernels in their lean layout (48 B per
# particle-tau across the one field barrier of a step) for ntau = 8, 16, 32, the two-barrier kernels otherwise.
# Tests that target the two-barrier kernels set this to _lib.STORE_FULL.
DEFAULT_STORAGE = None


def default_storage(ntau: int) -> int:
    if DEFAULT_STORAGE is not None:
        return DEFAULT_STORAGE
    return _lib.STORE_ONEPASS_LEAN if int(ntau) in (8, 16, 32) else _lib.STORE_FULL


class Session:
    def __init__(self, mesh: Mesh, ntau: int, eps: float, dt: float, nbpart: int, weight: float | None = None,
                 nbpart_global: int | None = None, wrap=_lib.WRAP_FORTRAN, deposit_mode=_lib.DEPOSIT_FP64_ATOMIC,
                 scheme=_lib.SCHEME_M6, storage_mode: int | None = None, device: int = 0, stream: int | None = None):
        if storage_mode is None:
            storage_mode = default_storage(ntau)
        self.storage_mode = storage_mode
        self.mesh, self.ntau, self.eps, self.dt = mesh, int(ntau), float(eps), float(dt)
        self.nbpart = int(nbpart)
        self.nbpart_global = int(nbpart_global if nbpart_global is not None else nbpart)
        dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
        self.weight = float(weight) if weight is not None else dimx * dimy / self.nbpart_global     # particles.F90:52
        cfg = ConfigStruct()
        cfg.mesh = mesh._struct()
        cfg.ntau, cfg.wrap, cfg.deposit_mode, cfg.scheme, cfg.storage_mode = int(ntau), wrap, deposit_mode, scheme, storage_mode
        cfg.device = int(device)
        cfg.eps, cfg.dt = float(eps), float(dt)
        cfg.nbpart = self.nbpart
        cfg.weight = self.weight
        cfg.total_mass = self.weight * self.nbpart_global
        cfg.stream = C.c_void_p(stream or 0)
        self._h = C.c_void_p()
        self._cb = None
        check(lib().uapic_session_create(C.byref(cfg), C.byref(self._h)))

    # ---- lifetime --------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            if getattr(self, "_peer_group", False) is not False:
                # nobody may unmap or free an exchange buffer a peer still reads: barrier, unmap, barrier, then free
                import torch.distributed as dist
                group, self._peer_group = self._peer_group, False
                dist.barrier(group=group)
                lib().uapic_session_close_peers(self._h)
                dist.barrier(group=group)
            lib().uapic_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- data ------------------------------------------------------------------------------------------
    def upload_particles(self, x: np.ndarray, v: np.ndarray):
        """x, v: (2, nbpart) Fortran-ordered float64 (numpy array, or a pinned torch tensor's numpy view)"""
        for a in (x, v):
            if a.dtype != np.float64 or a.shape != (2, self.nbpart) or not a.flags.f_contiguous:
                raise ValueError("x and v must be Fortran-ordered float64 arrays of shape (2, nbpart)")
        check(lib().uapic_session_upload_particles(self._h, x.ctypes.data_as(_dp), v.ctypes.data_as(_dp)))

    def upload_particles_ptr(self, x_ptr: int, v_ptr: int):
        check(lib().uapic_session_upload_particles(self._h, C.cast(x_ptr, _dp), C.cast(v_ptr, _dp)))

    def upload_particle_e(self, ep: np.ndarray):
        check(lib().uapic_session_upload_particle_e(self._h, ep.ctypes.data_as(_dp)))

    def upload_particle_e_ptr(self, ep_ptr: int):
        check(lib().uapic_session_upload_particle_e(self._h, C.cast(ep_ptr, _dp)))

    def set_sort(self, interval: int, bin_cells_log2: int = 3):
        """reorder the particle arrays by coarse mesh bin every `interval` steps (0 = never); invisible to the caller"""
        check(lib().uapic_session_set_sort(self._h, C.c_int(interval), C.c_int(bin_cells_log2)))

    def set_fusion(self, enable: bool = True):
        """run phase B of a step inside the first kernel of the next one (lean one-pass layout); results do not change"""
        check(lib().uapic_session_set_fusion(self._h, C.c_int(int(enable))))

    def enable_timing(self, enable: bool = True):
        check(lib().uapic_session_enable_timing(self._h, C.c_int(int(enable))))

    def phase_times(self):
        """(ms in phase A, ms in phase B, steps covered) since the last call; CUDA-event timed on the session stream"""
        a, b, n = C.c_double(0), C.c_double(0), C.c_int64(0)
        check(lib().uapic_session_phase_times(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def field_barrier_time(self) -> float:
        """ms between phase A and phase B (fold + all-reduce + field solve) over the steps the last phase_times() reported"""
        f = C.c_double(0)
        check(lib().uapic_session_field_barrier_time(self._h, C.byref(f)))
        return f.value

    def generate_particles(self, kind: str = "plasma", seed: int = 20190101, first_global_index: int = 0, alpha=0.05, kx=0.5,
                           index_stride: int = 1):
        """device-side load of this shard: global particle indices first, first+stride, ... (stride = world size and
        first = rank interleaves the shards, which balances the index-stratified |v| of the Landau load)"""
        k = {"plasma": 0, "landau": 1}[kind]
        check(lib().uapic_session_generate_particles_strided(self._h, C.c_int(k), C.c_uint64(seed), C.c_int64(first_global_index),
                                                             C.c_int64(index_stride), C.c_double(alpha), C.c_double(kx)))

    def download_particles(self):
        x = np.zeros((2, self.nbpart), order="F")
        v = np.zeros((2, self.nbpart), order="F")
        check(lib().uapic_session_download_particles(self._h, x.ctypes.data_as(_dp), v.ctypes.data_as(_dp)))
        return x, v

    def download_particles_ptr(self, x_ptr: int, v_ptr: int):
        check(lib().uapic_session_download_particles(self._h, C.cast(x_ptr, _dp), C.cast(v_ptr, _dp)))

    def download_particle_e(self):
        ep = np.zeros((2, self.nbpart), order="F")
        check(lib().uapic_session_download_particle_e(self._h, ep.ctypes.data_as(_dp)))
        return ep

    def download_fields(self):
        e = np.zeros((2, self.mesh.nx + 1, self.mesh.ny + 1), order="F")
        rho = np.zeros((self.mesh.nx + 1, self.mesh.ny + 1), order="F")
        check(lib().uapic_session_download_fields(self._h, e.ctypes.data_as(_dp), rho.ctypes.data_as(_dp)))
        return e, rho

    # ---- the loop --------------------------------------------------------------------------------------
    def init_fields(self):
        """compute_rho_m6_real -> solve_poisson -> interpolate_eb_m6_real     bupdate.F90:89-93"""
        check(lib().uapic_session_init_fields(self._h))

    def step(self, nsteps: int = 1):
        """nsteps UA steps (bupdate.F90:97-123); asynchronous on the session's stream"""
        check(lib().uapic_session_step(self._h, C.c_int(nsteps)))

    def step_host(self, x_in: np.ndarray, v_in: np.ndarray, e_in: np.ndarray | None, x_out: np.ndarray, v_out: np.ndarray):
        """one UA step with the particles in host memory ((2, nbpart) Fortran-ordered float64, ideally pinned): copies and
        kernels are pipelined chunk by chunk inside the library; returns when x_out, v_out are complete"""
        for a in (x_in, v_in, x_out, v_out) + ((e_in,) if e_in is not None else ()):
            if a.dtype != np.float64 or a.shape != (2, self.nbpart) or not a.flags.f_contiguous:
                raise ValueError("arrays must be Fortran-ordered float64 of shape (2, nbpart)")
        check(lib().uapic_session_step_host(self._h, x_in.ctypes.data_as(_dp), v_in.ctypes.data_as(_dp),
                                            e_in.ctypes.data_as(_dp) if e_in is not None else None,
                                            x_out.ctypes.data_as(_dp), v_out.ctypes.data_as(_dp)))

    def step_host_ptr(self, x_in: int, v_in: int, e_in: int, x_out: int, v_out: int):
        check(lib().uapic_session_step_host(self._h, C.cast(x_in, _dp), C.cast(v_in, _dp), C.cast(e_in, _dp) if e_in else None,
                                            C.cast(x_out, _dp), C.cast(v_out, _dp)))

    def synchronize(self):
        check(lib().uapic_session_synchronize(self._h))

    def energy_history(self) -> np.ndarray:
        n = C.c_int64(0)
        check(lib().uapic_session_energy_history(self._h, None, C.c_int64(0), C.byref(n)))
        out = np.zeros(n.value)
        if n.value:
            check(lib().uapic_session_energy_history(self._h, out.ctypes.data_as(_dp), C.c_int64(n.value), C.byref(n)))
        return out

    def sum_v(self) -> np.ndarray:
        out = np.zeros(2)
        check(lib().uapic_session_sum_v(self._h, out.ctypes.data_as(_dp)))
        return out

    @property
    def launch_count(self) -> int:
        n = C.c_int64(0)
        check(lib().uapic_session_launch_count(self._h, C.byref(n)))
        return n.value

    @property
    def device_bytes(self) -> int:
        n = C.c_int64(0)
        check(lib().uapic_session_device_bytes(self._h, C.byref(n)))
        return n.value

    # ---- multi-GPU ---------------------------------------------------------------------------------------
    def peer_handle(self) -> bytes:
        """64-byte CUDA IPC handle of this session's exchange buffer (uapic_session_peer_handle)"""
        buf = C.create_string_buffer(64)
        check(lib().uapic_session_peer_handle(self._h, buf))
        return buf.raw

    def init_peers(self, handles: bytes, nranks: int, rank: int):
        """map the exchange buffers of all ranks (handles concatenated in rank order): the sum of the raw rho meshes over the
        ranks then happens inside the field-solve kernel, over NVLink, without a collective call"""
        if len(handles) != 64 * nranks:
            raise ValueError("need one 64-byte handle per rank")
        buf = C.create_string_buffer(bytes(handles), len(handles))
        check(lib().uapic_session_init_peers(self._h, buf, C.c_int(nranks), C.c_int(rank)))

    def init_nccl(self, unique_id: bytes, nranks: int, rank: int):
        """collective: create this session's NCCL communicator inside the library (uapic_session_init_nccl); from then on
        the library itself sums the raw rho meshes with ncclAllReduce on the session's stream"""
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(lib().uapic_session_init_nccl(self._h, buf, C.c_int(nranks), C.c_int(rank)))

    def set_allreduce(self, fn):
        """fn(buf_ptr: int, count: int, dtype: int, stream: int) -> int ; dtype 0 = float64, 1 = int64"""
        if fn is None:
            self._cb = None
            check(lib().uapic_session_set_allreduce(self._h, C.cast(None, ALLREDUCE_FN), None))
            return

        def tramp(_ctx, buf, count, dtype, stream):
            try:
                return int(fn(buf or 0, int(count), int(dtype), stream or 0) or 0)
            except Exception:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        self._cb = ALLREDUCE_FN(tramp)
        check(lib().uapic_session_set_allreduce(self._h, self._cb, None))


def run_bupdate(mesh: Mesh, ntau: int, eps: float, dt: float, nstep: int, x: np.ndarray, v: np.ndarray, w: float | None = None,
                wrap=_lib.WRAP_FORTRAN, deposit_mode=_lib.DEPOSIT_FP64_ATOMIC, device: int = 0, storage_mode: int | None = None):
    """the whole program fortran/bupdate.F90:89-128 on one GPU: returns (x, v, energy[1+2*nstep], e_mesh)"""
    nbpart = x.shape[1]
    with Session(mesh, ntau, eps, dt, nbpart, weight=w, wrap=wrap, deposit_mode=deposit_mode, device=device,
                 storage_mode=storage_mode) as s:
        s.upload_particles(np.asfortranarray(x), np.asfortranarray(v))
        s.init_fields()
        s.step(nstep)
        s.synchronize()
        xo, vo = s.download_particles()
        e, _ = s.download_fields()
        return xo, vo, s.energy_history(), e

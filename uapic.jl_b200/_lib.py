"""ctypes binding of libuapic_b200.so (include/uapic_b200.h).

There is no fallback: if the shared library is missing, or a compute call is made without an
sm_100 device, this module raises.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UAPIC_B200_LIB lets a developer A/B-test another build of the SAME library (still CUDA-only, no fallback)
LIB_PATH = os.environ.get("UAPIC_B200_LIB") or os.path.join(_HERE, "libuapic_b200.so")

OK = 0
WRAP_FORTRAN, WRAP_JULIA = 0, 1
DEPOSIT_FP64_ATOMIC, DEPOSIT_FIXED_POINT = 0, 1
SCHEME_M6, SCHEME_CIC = 0, 1
STORE_FULL, STORE_HYBRID, STORE_ONEPASS, STORE_ONEPASS_LEAN = 0, 1, 2, 3

ERROR_NAMES = {-1: "UAPIC_EINVAL", -2: "UAPIC_ENODEVICE", -3: "UAPIC_ECUDA", -4: "UAPIC_ENOMEM", -5: "UAPIC_ESTATE",
               -6: "UAPIC_EUNSUPPORTED"}

# every symbol include/uapic_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "uapic_last_error", "uapic_version", "uapic_compiled_arch", "uapic_device_count", "uapic_fixed_point_scale", "uapic_probe_fp64_peak",
    "uapic_compute_rho_m6", "uapic_interpol_eb_m6", "uapic_poisson", "uapic_preparation", "uapic_interpol_eb_m6_tau",
    "uapic_compute_f", "uapic_fft_tau", "uapic_ua_step_predict", "uapic_ua_step_correct", "uapic_ua_step1",
    "uapic_ua_step2", "uapic_compute_rho_m6_tau", "uapic_compute_v",
    "uapic_session_create", "uapic_session_destroy", "uapic_session_set_allreduce", "uapic_nccl_unique_id", "uapic_nccl_version",
    "uapic_session_init_nccl", "uapic_session_set_nccl_comm", "uapic_session_peer_handle", "uapic_session_init_peers",
    "uapic_session_close_peers", "uapic_session_upload_particles", "uapic_session_upload_particle_e",
    "uapic_session_set_fusion", "uapic_session_enable_timing", "uapic_session_phase_times", "uapic_session_field_barrier_time", "uapic_session_set_sort",
    "uapic_session_generate_particles", "uapic_session_generate_particles_strided", "uapic_session_init_fields", "uapic_session_step", "uapic_session_step_host", "uapic_session_synchronize",
    "uapic_session_download_particles", "uapic_session_download_particle_e", "uapic_session_download_fields",
    "uapic_session_energy_history", "uapic_session_sum_v", "uapic_session_launch_count", "uapic_session_device_bytes",
    "uapic3d_create", "uapic3d_destroy", "uapic3d_init_nccl", "uapic3d_upload_particles", "uapic3d_generate_particles", "uapic3d_init_fields", "uapic3d_substep",
    "uapic3d_run", "uapic3d_download_particles", "uapic3d_download_fields", "uapic3d_launch_count", "uapic3d_compute_rho_cic",
    "uapic3d_poisson", "uapic3d_interpolate_eb_cic",
    "uapic_efd_run", "uapic_efd_run_device", "uapic_compute_rho_cic", "uapic_interpol_eb_cic",
]


class UapicError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class MeshStruct(C.Structure):
    _fields_ = [("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
                ("nx", C.c_int32), ("ny", C.c_int32)]


class ConfigStruct(C.Structure):
    _fields_ = [("mesh", MeshStruct), ("ntau", C.c_int32), ("wrap", C.c_int32), ("deposit_mode", C.c_int32),
                ("scheme", C.c_int32), ("storage_mode", C.c_int32), ("device", C.c_int32), ("eps", C.c_double),
                ("dt", C.c_double), ("nbpart", C.c_int64), ("weight", C.c_double), ("total_mass", C.c_double),
                ("stream", C.c_void_p)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p)

_lib = None


def lib() -> C.CDLL:
    """load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"or `make -C {os.path.join(_HERE, 'csrc')}`; uapic_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.uapic_last_error.restype = C.c_char_p
        for name in EXPORTS:
            if name != "uapic_last_error":
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise UapicError(rc, lib().uapic_last_error().decode())


def device_count() -> int:
    n = C.c_int(0)
    lib().uapic_device_count(C.byref(n))
    return n.value


def fixed_point_scale(total_mass: float) -> float:
    s = C.c_double(0.0)
    check(lib().uapic_fixed_point_scale(C.c_double(total_mass), C.byref(s)))
    return s.value


def probe_fp64_peak(device: int = 0, launches: int = 20):
    """(DFMA instructions/s of the whole chip in a pure FMA loop, ms per launch) -- bench.py --peaks"""
    r, ms = C.c_double(0.0), C.c_double(0.0)
    check(lib().uapic_probe_fp64_peak(C.c_int(device), C.c_int(launches), C.byref(r), C.byref(ms)))
    return r.value, ms.value

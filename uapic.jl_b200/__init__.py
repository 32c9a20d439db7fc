"""uapic_b200 -- B200-native UA-PIC time step behind the UAPIC.jl API (host-side mirror in Python; the Julia
`ccall` wrappers are in `julia/`).  The directory is named `uapic.jl_b200`; import it as `uapic_b200`
(the repo root has a small `uapic_b200.py` loader because a dot cannot appear in a package name).

All compute happens in `libuapic_b200.so` (hand-written sm_100a CUDA).  There is no CPU fallback.
"""
from ._lib import (DEPOSIT_FIXED_POINT, DEPOSIT_FP64_ATOMIC, LIB_PATH, SCHEME_CIC, SCHEME_M6, STORE_FULL, STORE_HYBRID,  # noqa: F401
                   STORE_ONEPASS, STORE_ONEPASS_LEAN, WRAP_FORTRAN, WRAP_JULIA, EXPORTS, UapicError, device_count, lib, probe_fp64_peak)
from .api import (UA, Mesh, MeshFields, Particles, Poisson, compute_f, compute_rho_cic, compute_rho_m6, compute_v, errors, fft_tau,  # noqa: F401
                  gnuplot, ifft_tau, integrate, interpol_eb_cic, interpol_eb_m6, preparation, ua_step, ua_step1, ua_step2, update_particles_e,
                  update_particles_x)
from .loaders import landau_sampling, make_particles_dat, plasma, plasma3d, read_particles, write_particles  # noqa: F401
from .session import Session, run_bupdate  # noqa: F401
from . import dist  # noqa: F401
from . import mrc3d  # noqa: F401
from .mrc3d import Fields3D, Mesh3D, Session3D, run_uapic3d, write_data  # noqa: F401
from . import external_field  # noqa: F401
from .external_field import efd, efd_run, efd_run_device  # noqa: F401

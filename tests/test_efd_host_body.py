"""The per-particle body of the CUDA external-field kernel (uapic.jl_b200/csrc/uapic_efd_body.cuh) compiled for the HOST with a
one-thread tau policy (tests/efd_host_body.cpp) and held to the oracle: the arithmetic the device runs is checked here without
a GPU; the device-side policies (shuffle FFT / shared-memory DFT) and the launch are what tests/test_gpu_efd.py adds."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle

_HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_body(tmp_path_factory):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("efd_host") / "libefd_host_body.so")
    subprocess.check_call([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", so,
                           os.path.join(_HERE, "efd_host_body.cpp")])
    lib = C.CDLL(so)
    lib.efd_host_body.restype = C.c_int
    return lib


@pytest.mark.parametrize("ntau", [4, 12, 16, 32])
@pytest.mark.parametrize("eps", [1e-1, 1e-3])
def test_device_body_on_the_host_matches_the_oracle(host_body, ntau, eps):
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(ntau)
    n = 200
    x = np.asfortranarray(rng.random((2, n)) * [[4 * np.pi], [2 * np.pi]])
    v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
    box = np.array([0.0, 4 * np.pi, 0.0, 2 * np.pi])
    xh, vh = x.copy(order="F"), v.copy(order="F")
    rc = host_body.efd_host_body(C.c_int(ntau), C.c_int64(n), C.c_double(eps), C.c_double(np.pi / 16), C.c_double(np.pi / 2), C.c_int(8),
                                 box.ctypes.data_as(dp), xh.ctypes.data_as(dp), vh.ctypes.data_as(dp))
    assert rc == 0
    xo, vo = oracle.corc().efd_run(x, v, ntau=ntau, eps=eps)
    # one ulp on the input moves v by ~1e-12 * 0.1/eps (tests/test_gpu_efd.py): same bound here
    assert np.abs(xh - xo).max() < 1e-12 * 4 * np.pi
    assert np.abs(vh - vo).max() < 1e-12 * max(1.0, 0.1 / eps) * max(1.0, np.abs(vo).max())

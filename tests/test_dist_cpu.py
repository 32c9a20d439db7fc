"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: contiguous particle sharding and the all-reduce of the
rho mesh.  The per-shard deposits come from the oracle (the product cannot run without a GPU); what is under test is
uapic_b200.dist, not the deposit."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import uapic_b200 as ub

from conftest import seeded_load


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 204800, 10**8 + 3):
        for world in (1, 2, 3, 8):
            edges = [ub.dist.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_shard_partitions_exactly():
    for n in (0, 1, 7, 204800, 10**8 + 3):
        for world in (1, 2, 3, 8):
            parts = [ub.dist.interleaved_shard(n, r, world) for r in range(world)]
            assert sum(c for _, _, c in parts) == n
            if n <= 204800:
                ids = np.concatenate([f + st * np.arange(c) for f, st, c in parts]) if n else np.zeros(0)
                assert np.array_equal(np.sort(ids), np.arange(n))
            assert max(c for _, _, c in parts) - min(c for _, _, c in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, npart, out_dir):
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    om, x, _ = seeded_load(npart, seed=4)
    w = 8 * np.pi ** 2 / npart
    lo, hi = ub.dist.shard_range(npart, rank, world)
    rho = np.zeros((om.nx + 1, om.ny + 1), order="F")
    # the epilogue (scale, subtract mean, ghost copy) is linear, so summing per-shard meshes equals the global mesh
    oracle.corc().compute_rho_m6(om, np.asfortranarray(x[:, lo:hi]), w, rho)
    total = ub.dist.allreduce_host(rho)
    if rank == 0:
        np.save(os.path.join(out_dir, "rho_sum.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_sharded_deposits_equals_global_deposit(tmp_path):
    import oracle
    npart, world = 5001, 2
    mp.spawn(_worker, args=(world, _free_port(), npart, str(tmp_path)), nprocs=world, join=True)
    om, x, _ = seeded_load(npart, seed=4)
    rho = np.zeros((om.nx + 1, om.ny + 1), order="F")
    oracle.corc().compute_rho_m6(om, x.copy(order="F"), 8 * np.pi ** 2 / npart, rho)
    got = np.load(tmp_path / "rho_sum.npy")
    assert np.abs(got - rho).max() < 1e-12 * np.abs(rho).max()


def test_library_loads_and_exports_every_declared_symbol():
    """C-ABI contract on a CPU-only box: the .so loads, exports everything include/uapic_b200.h declares, and every
    compute entry point fails loudly (no CPU fallback)."""
    import re
    lib = ub.lib()
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "uapic_b200.h")).read()
    declared = set(re.findall(r"\b(uapic(?:3d)?_[a-z0-9_]+)\s*\(", hdr)) - {"uapic_allreduce_fn"}
    assert declared == set(ub.EXPORTS), declared ^ set(ub.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert ub.lib().uapic_compiled_arch() == 100
    assert ub._lib.fixed_point_scale(8 * np.pi ** 2) == 2.0 ** 54
    if ub.device_count() == 0:
        m = ub.Mesh(0, 1, 8, 0, 1, 8)
        with pytest.raises(ub.UapicError) as e:
            ub.Poisson(m)(ub.MeshFields(m))
        assert e.value.code == -2          # UAPIC_ENODEVICE
        with pytest.raises(ub.UapicError):
            ub.Session(m, 16, 0.1, 0.1, 10)
        with pytest.raises(ub.UapicError) as e3:
            ub.Session3D(ub.Mesh3D((0, 0, 0), (1, 1, 1), (8, 8, 4)), 10)
        assert e3.value.code == -2


def test_nccl_is_bound_at_run_time_without_a_link_dependency():
    """uapic_nccl_*: the library dlopens libnccl.so.2 (no DT_NEEDED entry); the unique id is 128 opaque bytes"""
    import ctypes as C
    import subprocess
    so = ub.LIB_PATH
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "nccl" not in needed.lower()
    v = C.c_int(0)
    rc = ub.lib().uapic_nccl_version(C.byref(v))
    if rc != 0:
        pytest.skip("no libnccl.so.2 on this machine: " + ub.lib().uapic_last_error().decode())
    assert v.value >= 20000
    a, b = ub.dist.nccl_unique_id(), ub.dist.nccl_unique_id()
    assert len(a) == 128 and a != b


def test_particles_dat_roundtrip(tmp_path):
    """src/read_particles.jl format: `ix iy dpx dpy vx vy`; test/test_particles.jl:4-14 pins nbpart == 204800 on the real file"""
    mesh = ub.Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)
    p = ub.plasma(mesh, 2000, seed=1)
    f = tmp_path / "particles.dat"
    ub.write_particles(str(f), mesh, p.x, p.v)
    q = ub.read_particles(str(f), mesh)
    assert q.nbpart == 2000 and q.w == pytest.approx(8 * np.pi ** 2 / 2000)
    assert np.abs(q.x - p.x).max() < 1e-13 and np.array_equal(q.v, p.v)
    # densities of fortran/particles.F90:77,94
    big = ub.plasma(mesh, 100000, seed=2)
    assert abs(np.sin(big.x[1]).mean() - 0.5) < 0.01 and abs((big.v[0] ** 2).mean() - 5.0) < 0.1


def test_gfortran_stream_reproduces_the_reference_seed():
    """fortran/particles.F90:54-66: the bundled libgfortran accepts the reference's 33-word seed; the draw is deterministic"""
    mesh = ub.Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)
    a, src = ub.plasma(mesh, 300, use_gfortran=True, return_source=True)
    if "libgfortran" not in src:
        pytest.skip("libgfortran not loadable here")
    b = ub.plasma(mesh, 300, use_gfortran=True)
    assert np.array_equal(a.x, b.x) and np.array_equal(a.v, b.v)
    assert 0 <= a.x[0].min() and a.x[0].max() < 4 * np.pi


def test_write_data_xdmf_dump(tmp_path):
    """output.f90:10-110: fields-NNNN.xmf describing ex, ey, ez, rho on the (nx+1, ny+1, nz+1) nodes; heavy data beside it"""
    import xml.etree.ElementTree as ET
    mesh = ub.Mesh3D((0, 0, 0), (18, 18, 1), (6, 4, 2))
    f = ub.Fields3D(mesh)
    rng = np.random.default_rng(0)
    f.e[...] = rng.normal(size=f.e.shape)
    f.rho[...] = rng.normal(size=f.rho.shape)
    path = ub.write_data(1, f, str(tmp_path))
    assert path.endswith("fields-0001.xmf")
    root = ET.parse(path).getroot()
    grid = root.find("Domain/Grid")
    assert grid.find("Topology").get("NumberOfElements").split() == ["3", "5", "7"]          # nz+1, ny+1, nx+1
    geo = [list(map(float, d.text.split())) for d in grid.findall("Geometry/DataItem")]
    assert geo == [[0.0, 0.0, 0.0], [3.0, 4.5, 0.5]]
    raw = np.fromfile(tmp_path / "fields-0001.bin", dtype="<f8")
    want = {"ex": f.e[0], "ey": f.e[1], "ez": f.e[2], "rho": f.rho}
    for a in grid.findall("Attribute"):
        item = a.find("DataItem")
        off = int(item.get("Seek")) // 8
        got = raw[off:off + f.rho.size].reshape(f.rho.shape, order="F")
        assert item.text.strip() == "fields-0001.bin" and np.array_equal(got, want[a.get("Name")])
    with pytest.raises(ValueError):
        ub.write_data(10000, f, str(tmp_path))


def test_plasma3d_consumes_the_gfortran_stream_in_the_reference_order():
    """fortran/particles.F90:107-190 (init_particles_3d): the vectorised loader must hand out exactly the particles a
    one-deviate-at-a-time loop over the same libgfortran stream produces"""
    import ctypes as C
    from importlib import import_module
    n = 700
    x, v, src = ub.plasma3d((0, 0, 0), (18, 18, 1), n, use_gfortran=True, return_source=True)
    if "libgfortran" not in src:
        pytest.skip("libgfortran not loadable here")
    u = import_module(ub.plasma.__module__)._Uniforms(None, True)
    u.gf._gfortran_random_r8.argtypes = [C.POINTER(C.c_double)]
    tmp = C.c_double()

    def rn():
        u.gf._gfortran_random_r8(C.byref(tmp))
        return tmp.value

    xs, vs = np.zeros((3, n), order="F"), np.zeros((3, n), order="F")
    for m in range(n):
        for c in range(3):
            xs[c, m] = rn()
    m = 0
    while m < n:
        xi, yi, zi = 9.0 * rn(), 2.0 * np.pi * rn(), 1.02 * rn()
        if (1.0 + 0.02 * np.cos(4.0 * yi)) * np.exp(-5.0 * (xi - 4.8) ** 2) >= zi:
            xs[0, m], xs[1, m] = np.cos(yi) * xi + 9.0, np.sin(yi) * xi + 9.0
            m += 1
    m = 0
    while m < n:
        xi, yi, wi, zi = (rn() - 0.5) * 8.0, (rn() - 0.5) * 8.0, (rn() - 0.5) * 8.0, rn()
        if np.exp(-2.0 * (xi ** 2 + yi ** 2 + wi ** 2)) >= zi:
            vs[:, m] = xi, yi, wi
            m += 1
    assert np.array_equal(x, xs) and np.array_equal(v, vs)
    # densities with a seeded numpy stream: ring of radius ~4.8 around (9, 9), <v_i^2> = 1/4
    xb, vb = ub.plasma3d((0, 0, 0), (18, 18, 1), 50000, seed=5)
    r = np.hypot(xb[0] - 9.0, xb[1] - 9.0)
    assert abs(r.mean() - 4.8) < 0.02 and np.abs((vb ** 2).mean(axis=1) - 0.25).max() < 0.01 and 0 <= xb[2].min() and xb[2].max() < 1

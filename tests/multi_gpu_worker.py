"""Worker for tests/test_gpu_multi.py: launched by torch.distributed.run with one rank per GPU.
Runs the sharded session (NCCL all-reduce of the raw rho mesh through the C ABI hook) and, on rank 0, the same
problem on a single GPU; writes the comparison to a JSON file."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import uapic_b200 as ub  # noqa: E402
from conftest import seeded_load  # noqa: E402


def run(mode, npart, ntau, nstep, x0, v0, mesh, w, rank, world, local, reducer="nccl"):
    lo, hi = ub.dist.shard_range(npart, rank, world)
    s = ub.Session(mesh, ntau, 0.1, np.pi / 16, hi - lo, weight=w, nbpart_global=npart, deposit_mode=mode, device=local,
                   stream=torch.cuda.current_stream().cuda_stream)
    if world > 1:
        if reducer == "peer":
            ub.dist.attach_peer_exchange(s)     # no collective: the solve kernel sums the ranks' buffers over NVLink
        elif reducer == "nccl":
            ub.dist.attach_nccl(s)              # in-library ncclAllReduce (uapic_session_init_nccl)
        else:
            ub.dist.attach_torch_allreduce(s)   # host callback hook (uapic_session_set_allreduce)
    s.upload_particles(np.asfortranarray(x0[:, lo:hi]), np.asfortranarray(v0[:, lo:hi]))
    s.init_fields()
    s.step(nstep)
    s.synchronize()
    x, v = s.download_particles()
    e, _ = s.download_fields()
    en = s.energy_history()
    s.close()
    return x, v, en, e


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    npart, ntau, nstep = 40001, 16, 3
    _, x0, v0 = seeded_load(npart, seed=55)
    mesh = ub.Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)
    w = 8 * np.pi ** 2 / npart
    res = {}
    for name, mode, reducer in (("fixed", ub.DEPOSIT_FIXED_POINT, "nccl"), ("fp64", ub.DEPOSIT_FP64_ATOMIC, "nccl"),
                                ("fixed_callback", ub.DEPOSIT_FIXED_POINT, "torch"), ("fixed_peer", ub.DEPOSIT_FIXED_POINT, "peer"),
                                ("fp64_peer", ub.DEPOSIT_FP64_ATOMIC, "peer")):
        x, v, en, e = run(mode, npart, ntau, nstep, x0, v0, mesh, w, rank, world, local, reducer)
        # gather the shards on rank 0
        xs = [None] * world
        vs = [None] * world
        dist.all_gather_object(xs, x)
        dist.all_gather_object(vs, v)
        ens = [None] * world
        dist.all_gather_object(ens, en)
        if rank == 0:
            xg, vg = np.concatenate(xs, axis=1), np.concatenate(vs, axis=1)
            x1, v1, en1, e1 = run(mode, npart, ntau, nstep, x0, v0, mesh, w, 0, 1, local)
            res[name] = {
                "bit_identical_x": bool(np.array_equal(xg, x1)), "bit_identical_v": bool(np.array_equal(vg, v1)),
                "bit_identical_energy": bool(np.array_equal(en, en1)), "bit_identical_emesh": bool(np.array_equal(e, e1)),
                "ranks_agree_on_energy": bool(all(np.array_equal(ens[0], q) for q in ens)),
                "max_abs_dx": float(np.abs(xg - x1).max()), "max_abs_dv": float(np.abs(vg - v1).max()),
                "max_rel_denergy": float(np.abs(en - en1).max() / np.abs(en1).max()),
            }
        dist.barrier()
    # ---- the 3D program (fortran/uapic3d.f90), particles sharded, fixed-point deposits: bit-identical to one GPU ----
    import oracle
    mesh3 = ub.Mesh3D((0, 0, 0), (18, 18, 1), (16, 16, 4))
    np3 = 20001
    x3, v3 = oracle.corc3().generate(oracle.mesh3((0, 0, 0), (18, 18, 1), (16, 16, 4)), 77, np3)
    lo, hi = ub.dist.shard_range(np3, rank, world)

    def run3(lo, hi, sharded):
        # index_quirk off: p%x(m,1) indexes the flattened GLOBAL particle array (uapic3d.f90:179,182), which a shard does not hold
        with ub.Session3D(mesh3, hi - lo, nbpart_global=np3, deposit_mode=ub.DEPOSIT_FIXED_POINT, device=local, index_quirk=False) as s3:
            if sharded:
                ub.dist.attach_nccl(s3)
            s3.upload_particles(np.asfortranarray(x3[:, lo:hi]), np.asfortranarray(v3[:, lo:hi]))
            s3.init_fields()
            s3.run(4, 4, np.pi, 2)
            xa, va, _ = s3.download_particles()
            return xa, va, s3.download_fields().e

    xa, va, ea = run3(lo, hi, world > 1)
    xs, vs = [None] * world, [None] * world
    dist.all_gather_object(xs, xa)
    dist.all_gather_object(vs, va)
    if rank == 0:
        x1, v1, e1 = run3(0, np3, False)
        res["mrc3d_fixed"] = {"bit_identical_x": bool(np.array_equal(np.concatenate(xs, axis=1), x1)),
                              "bit_identical_v": bool(np.array_equal(np.concatenate(vs, axis=1), v1)),
                              "bit_identical_e": bool(np.array_equal(ea, e1))}
    dist.barrier()
    if rank == 0:
        with open(out, "w") as f:
            json.dump({"world": world, **res}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- throughput of the UA-PIC time step (BASELINE.json metric: particle-tau updates per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config3|config3-shard|config2|config5] [--impl b200|reference]

One process per GPU (the driver launches N>1 through torch.distributed.run).  A "step" is one full UA step
(fortran/bupdate.F90:97-123: predictor + corrector, two deposits, two Poisson solves) over every particle-tau
sample of the workload.  Default workload = BASELINE config 3, the case the metric is quoted on: 4D Landau load,
1e8 particles, ntau = 32, 128 x 128 mesh, M6 -- the SAME 1e8-particle problem at every N (strong scaling).
--storage auto picks the one-pass kernels (one field barrier per step) in their lean layout, 48 B per
particle-tau across the barrier (1e8 particles on ONE GPU = 154 GB of 180), else the legacy hybrid layout (16 B per
particle-tau, predictor recomputed).

Prints ONE JSON line (rank 0).  `value` times K steps with all state resident in HBM; `e2e` times the same step
through the host-facing API with the particles living in pinned HOST memory (x, v, e copied up, x, v and the energy
copied down, every step); `roofline` is for the dominant fused kernel, CUDA-event timed inside the run;
`cpu_baseline` is the CPU oracle (a port of the Fortran reference, OpenMP on all host cores) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIMX, DIMY = 4 * np.pi, 2 * np.pi
DT = np.pi / 16          # bupdate.F90:66
EPS = 0.1                # bupdate.F90:18

WORKLOADS = {
    # name: (load, ntau, nx, ny, particles, "total" (strong scaling) | "per_gpu" (weak scaling), description)
    "config3": ("landau", 32, 128, 128, 100_000_000, "total",
                "BASELINE config 3: 4D Landau load, 1e8 particles, ntau=32, 128x128 mesh, M6, strong scaling over the GPUs"),
    "config3-shard": ("landau", 32, 128, 128, 12_500_000, "per_gpu",
                      "per-GPU shard of BASELINE config 3 (12.5e6 particles/GPU = 1e8 at 8 GPUs), weak scaling"),
    "config2": ("plasma", 16, 128, 64, 1_000_000, "total", "BASELINE config 2: bupdate case, M6, 1e6 particles, ntau=16, 128x64"),
    "config4": ("plasma", 16, 128, 64, 10_000_000, "total",
                "BASELINE config 4: eps sweep case, bupdate densities, 1e7 particles, ntau=16, 128x64, M6 (pass --eps)"),
    "config5": ("landau", 32, 256, 256, 15_625_000, "per_gpu",
                "BASELINE config 5 weak point: 5e8 particle-tau samples/GPU, ntau=32, 256x256, M6"),
    "config5-strong": ("landau", 32, 256, 256, 31_250_000, "total",
                       "BASELINE config 5 strong: 1e9 particle-tau samples in total, ntau=32, 256x256, M6"),
}
SEED = 20190101


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # "under load": samples in the upper half of the power range seen
            pw = np.array(power)
            load = pw >= (pw.min() + 0.5 * (pw.max() - pw.min())) if pw.max() > pw.min() else np.ones_like(pw, bool)
            out.update(sm_mhz=float(np.median(np.array(sm)[load])), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(pw.max()))
        return out


def cpu_oracle_rate(workload, steps, warmup, sample_particles, np_global, eps, threads=None, faithful=True):
    """particle-tau updates/s of the CPU oracle (port of the Fortran reference) on a bounded sample of the workload: the
    SAME particles the device generator makes (oracle.generate = k_generate, same seed), every (np_global/sample)-th one,
    so the sample spans the |v| strata of the Landau load like the full problem does"""
    import oracle
    load, ntau, nx, ny = WORKLOADS[workload][:4]
    orc = oracle.corc()
    nthreads = threads or orc.max_threads()
    orc.set_threads(nthreads)
    om = oracle.mesh(0, DIMX, nx, 0, DIMY, ny)
    n = min(sample_particles, np_global)
    x, v = orc.generate(om, load, SEED, n, np_global=np_global, first=0, stride=max(1, np_global // n))
    sim = orc.sim(om, ntau, eps, DT, x, v, DIMX * DIMY / n, faithful=faithful)
    sim.init()
    for _ in range(warmup):
        sim.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    sim.close()
    orc.set_threads(1)
    return n * ntau * steps / dt, dt / steps * 1e3, nthreads


def cpu_baseline_block(workload, np_global, eps, sample, steps=2, warmup=1):
    """all host cores, and one core on an eighth of the sample (B1 = the stand-in for the README's
    single-core Fortran figure)"""
    ntau = WORKLOADS[workload][1]
    rate, _, threads = cpu_oracle_rate(workload, steps, warmup, sample, np_global, eps)
    s1 = max(2000, sample // 8)
    rate1, _, _ = cpu_oracle_rate(workload, steps, warmup, s1, np_global, eps, threads=1)
    # the Fortran's third interpolation per step is dead work (bupdate.F90:121, result never read) that the GPU path skips:
    # the same port without it, so the ratio can be read both ways ("inflates the ratio by roughly 15 %")
    rate_lean, _, _ = cpu_oracle_rate(workload, steps, warmup, sample, np_global, eps, faithful=False)
    what = "C port of fortran/bupdate.F90 (3 gathers/step as the Fortran does), same generated particles as the GPU arm (every k-th)"
    return {"value": rate, "unit": "particle-tau updates/s", "cores": threads, "kind": "port",
            "sample": f"{min(sample, np_global)} particles x ntau={ntau} x {steps} steps of the same workload; {what}",
            "one_core": {"value": rate1, "cores": 1, "sample": f"{min(s1, np_global)} particles x ntau={ntau} x {steps} steps"},
            "without_dead_third_gather": {"value": rate_lean, "cores": threads, "note": "same port, 2 gathers/step like the GPU path"}}


def run_peaks(args):
    """--peaks: measure the chip's DFMA issue rate (uapic_probe_fp64_peak) and write profiles/fp64_peak.json"""
    import torch
    import uapic_b200 as ub
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local)
    best, best_ms = 0.0, 0.0
    for _ in range(5):
        r, ms = ub.probe_fp64_peak(local, 20)
        if r > best:
            best, best_ms = r, ms
    clocks = sampler.stop()
    out = {"dfma_per_s": best, "fp64_tflops": 2 * best / 1e12, "ms_per_launch": best_ms,
           "how": "pure DFMA loop, 8 independent chains/thread, 256 threads x 8 CTAs/SM, 2^16 iterations, best of 5 x 20 launches (CUDA events)",
           "gpu": torch.cuda.get_device_name(local), "clocks": clocks,
           "nominal_dfma_per_s": 148 * 64 * 1.965e9}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for d in ("profiles", "gpurun_out"):
        with open(os.path.join(ROOT, d, "fp64_peak.json"), "w") as f:
            json.dump(out, f, indent=1)
    emit(out)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Neither Julia nor a Fortran compiler nor
    FFTW exist in this image, so it is the line-by-line C port (oracle/uapic_oracle.c, three gathers per step as the
    Fortran does), OpenMP over all host cores, on a bounded particle sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    load, ntau, nx, ny, np_cfg, mode, desc = WORKLOADS[args.workload]
    np_global = args.particles or (np_cfg * args.gpus if mode == "per_gpu" else np_cfg)
    sample = args.cpu_sample or 200_000
    rate, ms, threads = cpu_oracle_rate(args.workload, args.steps, args.warmup, sample, np_global, args.eps)
    line = {
        "impl": "reference", "metric": "particle-tau updates/sec", "value": rate, "unit": "particle-tau updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if mode == "total" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "load": load, "ntau": ntau, "mesh": [nx, ny], "eps": args.eps, "dt": DT, "scheme": "M6",
                   "particles_total": np_global,
                   "note": "CPU port timed on a bounded sample of the same generated particles (every k-th of the full load); cost is linear in the particle count"},
        "cpu_baseline": {"value": rate, "unit": "particle-tau updates/s", "cores": threads, "kind": "port",
                         "sample": f"{min(sample, np_global)} particles x ntau={ntau} x {args.steps} steps of the same workload (Fortran-faithful: 3 gathers/step)"},
        "e2e": {"value": rate, "unit": "particle-tau updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """the contract is ONE JSON line on stdout: libraries that print there from C (NCCL's "NCCL version ..." banner when a
    communicator is created under NCCL_DEBUG=VERSION/WARN) are sent to stderr; emit() writes the line to the real stdout"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--particles-per-gpu", type=int, default=0)
    ap.add_argument("--deposit", default="fp64", choices=["fp64", "fixed"])
    ap.add_argument("--storage", default="auto", choices=["auto", "full", "hybrid", "onepass", "onepass-lean"],
                    help="what crosses the field barrier per particle-tau: one-pass kernels 72 B (onepass) / 48 B (onepass-lean); "
                         "legacy two-barrier kernels 128 B (full) / 16 B + recompute (hybrid)")
    ap.add_argument("--scheme", default="m6", choices=["m6", "cic"],
                    help="shape function: m6 = what the reference ships (the metric is quoted on it); cic = build-defined bilinear variant")
    ap.add_argument("--reduce", default="nccl", choices=["nccl", "torch", "peer"],
                    help="N > 1: who sums the raw rho meshes -- nccl: ncclAllReduce enqueued by the library (default); torch: host callback; "
                         "peer: inside the field-solve kernel, out of the ranks' IPC-mapped buffers over NVLink (no collective)")
    ap.add_argument("--eps", type=float, default=EPS, help="the small parameter (bupdate.F90:18: 0.1); BASELINE config 4 sweeps 1e-1 .. 1e-5")
    ap.add_argument("--particles", type=int, default=0, help="override the workload's TOTAL particle count (strong scaling)")
    ap.add_argument("--peaks", action="store_true", help="measure the fp64 DFMA peak of the device, write profiles/fp64_peak.json, exit")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    quiet_stdout()

    if args.impl == "reference":
        run_reference(args)
        return
    if args.peaks:
        run_peaks(args)
        return

    import torch
    import uapic_b200 as ub

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: uapic_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    load, ntau, nx, ny, np_cfg, mode, desc = WORKLOADS[args.workload]
    if args.particles_per_gpu:
        mode, np_cfg = "per_gpu", args.particles_per_gpu
        desc += f" [overridden: {np_cfg} particles per GPU, weak scaling]"
    if args.particles:
        mode, np_cfg = "total", args.particles
        desc += f" [overridden: {np_cfg} particles in total]"
    eps = args.eps
    np_global = np_cfg * world if mode == "per_gpu" else np_cfg
    np_gpu = np_global // world
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    first, stride, n_mine = ub.dist.interleaved_shard(np_global, rank, world)   # global index = rank + k*world (see dist.py)
    lo, hi = 0, n_mine
    stream = torch.cuda.current_stream().cuda_stream
    free_b, _ = torch.cuda.mem_get_info()
    PER_TAU = {"full": 128, "hybrid": 16, "onepass": 72, "onepass-lean": 48}
    MODES = {"full": ub.STORE_FULL, "hybrid": ub.STORE_HYBRID, "onepass": ub.STORE_ONEPASS, "onepass-lean": ub.STORE_ONEPASS_LEAN}
    storage = args.storage
    sort_on = True
    if storage == "auto":
        storage = "hybrid"
        # the lean one-pass layout is also the faster one (W_n and interv are cheaper to redo).  Per particle: the store,
        # 128 B of particle arrays and records, 58 B of reordering buffers; + 1.5 GiB of slack.  If only the reordering
        # buffers do not fit, run without reordering (-9 %) rather than fall back to the two-barrier kernels (-55 %).
        if ntau in (8, 16, 32):
            base = (hi - lo) * (ntau * PER_TAU["onepass-lean"] + 128) + (3 << 29)
            if base + (hi - lo) * 58 < free_b:
                storage = "onepass-lean"
            elif base < free_b:
                storage, sort_on = "onepass-lean", False
    s = ub.Session(mesh, ntau, eps, DT, hi - lo, nbpart_global=np_global, device=local, stream=stream,
                   deposit_mode=ub.DEPOSIT_FIXED_POINT if args.deposit == "fixed" else ub.DEPOSIT_FP64_ATOMIC,
                   storage_mode=MODES[storage], scheme=ub.SCHEME_CIC if args.scheme == "cic" else ub.SCHEME_M6)
    if not sort_on:
        s.set_sort(0)
    if world > 1:
        if args.reduce == "peer":
            ub.dist.attach_peer_exchange(s)        # no collective: the field-solve kernel sums the ranks' deposit meshes over NVLink
        elif args.reduce == "nccl":
            ub.dist.attach_nccl(s)                 # the library owns the communicator and enqueues ncclAllReduce itself
        else:
            ub.dist.attach_torch_allreduce(s)      # host-callback hook (torch.distributed)
    # interleaved shards (global index = rank + k*world): the Landau load stratifies |v| by particle index
    s.generate_particles(load, seed=SEED, first_global_index=first, index_stride=stride)
    s.init_fields()
    s.step(args.warmup)
    s.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None

    # ---- device-resident throughput ------------------------------------------------------------------
    s.enable_timing(True)
    launches0 = s.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    s.step(args.steps)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = s.launch_count - launches0
    ms_a, ms_b, nt = s.phase_times()
    s.enable_timing(False)
    ms_step = ms_total / args.steps
    units_step = np_global * ntau
    value = units_step / (ms_step * 1e-3)

    # ---- end to end through the host-facing API, particles in pinned host memory ------------------------
    e2e = None
    if not args.no_e2e:
        n_loc = hi - lo
        hx = torch.empty((n_loc, 2), dtype=torch.float64).pin_memory()
        hv = torch.empty((n_loc, 2), dtype=torch.float64).pin_memory()
        he = torch.empty((n_loc, 2), dtype=torch.float64).pin_memory()
        s.download_particles_ptr(hx.data_ptr(), hv.data_ptr())
        he.numpy()[:] = s.download_particle_e().T
        e2e_steps = max(3, min(args.steps, 5))

        def one():
            # ONE C-ABI call per step: host buffers in, host buffers out; the library pipelines copies and kernels
            s.step_host_ptr(hx.data_ptr(), hv.data_ptr(), he.data_ptr(), hx.data_ptr(), hv.data_ptr())
            return s.energy_history()[-1]

        for _ in range(2):
            one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nrj = one()
        torch.cuda.synchronize()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        e2e = {"value": units_step / (ms_e2e * 1e-3), "unit": "particle-tau updates/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(3 * 16 * n_loc), "d2h_bytes_per_step": int(2 * 16 * n_loc + 8 * s.energy_history().size),
               "steps": e2e_steps, "last_energy": float(nrj),
               "note": "uapic_session_step_host: x, v, e (2,np) copied up from pinned host memory and x, v + the energy history copied back every step, per rank; copies of one chunk overlap the kernels of its neighbours"}
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        hbm, hbm_src = peaks()
        n_loc = hi - lo
        # ALGORITHMIC bytes: 128 B per particle-tau written by the first kernel and read back by the
        # second, plus 64 / 48 B per particle; B_alg = 256 + 112/ntau per particle-tau update for the whole step.  The
        # kernels are graded on this figure whatever they really move (impl_bytes: 72 / 48 / 16 / 128 B per particle-tau).
        onepass = storage.startswith("onepass")
        bytes_a = n_loc * ntau * 128 + n_loc * 64
        bytes_b = n_loc * ntau * 128 + n_loc * 48
        impl_a = n_loc * ntau * PER_TAU[storage] + n_loc * ((48 + 16 + 64) if onepass else 64)
        impl_b = n_loc * ntau * PER_TAU[storage] + n_loc * ((64 + 16) if onepass else 48)
        per_a, per_b = ms_a / max(nt, 1), ms_b / max(nt, 1)
        ka, kb = ("uapic::k_onepass_a", "uapic::k_onepass_b") if onepass else ("uapic::k_phase_a", "uapic::k_phase_b")
        dom = (kb, bytes_b, per_b, impl_b) if per_b >= per_a else (ka, bytes_a, per_a, impl_a)
        achieved = dom[1] / (dom[2] * 1e-3) / 1e9 if dom[2] > 0 else 0.0
        b_alg = 256 + 112 / ntau
        # DRAM traffic of the dominant kernel: dram__bytes_read + dram__bytes_write of an ncu capture of the same kernels at
        # 12.5e6 particles (profiles/r2e_traffic.json; r1n_traffic.json = 2e6 particles as fallback), scaled to this launch's
        # particle-tau count (traffic is linear in it: no reuse across particles)
        traffic, traffic_src = None, None
        for tfile in ("r2e_traffic.json", "r1n_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", tfile)) as f:
                    tj = json.load(f)
                kshort = dom[0].split("::")[-1]
                if onepass and storage == "onepass-lean" and ntau == tj["ntau"] and kshort in tj:
                    per_unit = (tj[kshort]["dram_bytes_read"] + tj[kshort]["dram_bytes_write"]) / (tj["particles"] * tj["ntau"])
                    traffic = int(per_unit * n_loc * ntau)
                    traffic_src = f"ncu capture at {tj['particles']} particles scaled by particle-tau count ({per_unit:.1f} B each); profiles/{tfile}"
                    break
            except (OSError, KeyError, ValueError):
                pass
        # fp64 pipe beside HBM: executed fp64 lane-instructions per update (ncu opcode census of both kernels,
        # profiles/r2e_sass_mix.txt -> profiles/fp64_instr.json) x rate / the DFMA rate measured on this pool (bench.py --peaks -> profiles/fp64_peak.json)
        fp64 = None
        try:
            with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
                pk = json.load(f)
            with open(os.path.join(ROOT, "profiles", "fp64_instr.json")) as f:
                fi = json.load(f)
            per_update = fi["fp64_lane_instr_per_update"] if (onepass and args.scheme == "m6" and ntau == fi["ntau"]) else None
            if per_update:
                fp64 = {"fp64_lane_instr_per_update": per_update, "peak_dfma_per_s": pk["dfma_per_s"], "peak_source": "measured DFMA loop (profiles/fp64_peak.json)",
                        "frac": value / world * per_update / pk["dfma_per_s"]}
        except (OSError, KeyError, ValueError):
            pass
        ms_field = s.field_barrier_time() / max(nt, 1) if onepass else None
        line = {
            "metric": "particle-tau updates/sec", "value": value, "unit": "particle-tau updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if mode == "total" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "load": load, "ntau": ntau, "mesh": [nx, ny], "eps": eps, "dt": DT, "scheme": args.scheme.upper(),
                       "particles_per_gpu": np_gpu, "particles_total": np_global, "deposit": args.deposit,
                       "storage": {"full": "store-full (128 B per particle-tau across the intra-step barrier, two barriers per step)",
                                   "hybrid": "hybrid (16 B per particle-tau across the barrier, predictor recomputed in phase B)",
                                   "onepass": "one-pass (one field barrier per step, 72 B per particle-tau across it)",
                                   "onepass-lean": "one-pass lean (one field barrier per step, 48 B per particle-tau across it)"}[storage],
                       "hbm_free_gb_before_alloc": round(free_b / 1e9, 1), "particle_reordering": "every step, 8x8-cell bins" if (sort_on and onepass) else "off",
                       "l2": f"inputs larger than L2: {s.device_bytes / 1e9:.1f} GB of particle state per GPU streamed every step",
                       "parallelism": f"particle shards x{world}, allreduce(rho) over NCCL ({ {'nccl': 'in-library ncclAllReduce', 'torch': 'torch.distributed callback', 'peer': 'summed inside the field-solve kernel over NVLink peer memory, no collective'}[args.reduce] })" if world > 1 else "single GPU"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": hbm, "unit": "GB/s",
                         "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                         "ms_per_launch": dom[2], "algorithmic_bytes_per_launch": int(dom[1]),
                         "impl_bytes_per_launch": int(dom[3]), "impl_gbs": dom[3] / (dom[2] * 1e-3) / 1e9 if dom[2] > 0 else 0.0,
                         "phase_a_ms": per_a, "phase_b_ms": per_b,
                         "whole_step_frac_of_hbm": value * b_alg / 1e9 / hbm / world,
                         "algorithmic_bytes_per_update": b_alg,
                         "achieved_is": "ALGORITHMIC bytes (SURVEY 8d: 128 B per particle-tau each way) / kernel time -- the graded figure, not DRAM throughput; traffic_gbs is what DRAM really moves",
                         "traffic_gbs": (traffic / (dom[2] * 1e-3) / 1e9) if (traffic and dom[2] > 0) else None,
                         "field_barrier_ms": ms_field,
                         "fp64": fp64,
                         "binding_unit": "L1TEX LSU data pipe (register write-back, 128 B/clk/SM): 76-87 % busy in both kernels (profiles/README.md); DRAM 11-18 % busy"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            # rank 0 only, after every timed region (the other ranks wait at the final barrier)
            line["cpu_baseline"] = cpu_baseline_block(args.workload, np_global, eps, args.cpu_sample or (100_000 if world == 1 else 50_000))
        else:
            line["cpu_baseline"] = None
        emit(line)
    s.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

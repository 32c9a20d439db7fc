"""A/B timing of the phase kernels: python time_phases.py NP mode[,mode] -> ms of A and B per step (CUDA events)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
import numpy as np
import uapic_b200 as ub
DT = np.pi / 16
DIMX, DIMY = 4 * np.pi, 2 * np.pi
npart = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ntau = int(os.environ.get('NTAU', '32')); nx = int(os.environ.get('NX', '128')); ny = int(os.environ.get('NY', '128'))
mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
modes = {"onepass": ub.STORE_ONEPASS, "lean": ub.STORE_ONEPASS_LEAN, "full128": ub.STORE_FULL, "hybrid": ub.STORE_HYBRID}
tag = os.path.basename(os.environ.get("UAPIC_B200_LIB", "default"))
for mname in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["onepass", "lean"]):
    with ub.Session(mesh, ntau, 0.1, DT, npart, storage_mode=modes[mname]) as s:
        if os.environ.get("SORT") is not None:
            s.set_sort(int(os.environ["SORT"]), int(os.environ.get("SORTLOG", "3")))
        s.generate_particles("landau", seed=1)
        s.init_fields()
        s.step(3); s.synchronize()
        s.enable_timing(True)
        s.step(4); s.synchronize()
        a, b, n = s.phase_times()
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.step(4); e1.record(); s.synchronize(); torch.cuda.synchronize()
        print(f"   whole step {e0.elapsed_time(e1) / 4:.3f} ms -> {npart * ntau / (e0.elapsed_time(e1) / 4 * 1e-3):.3e} upd/s")
        e = s.energy_history()
        print(f"{tag:28s} {mname:8s} np={npart} A={a / n:.3f} B={b / n:.3f} total={(a + b) / n:.3f} ms  ({npart * ntau / ((a + b) / n * 1e-3):.3e} upd/s kernels only) energy[-1]={e[-1]:.12f}", flush=True)

"""configs[3] on ONE GPU (or under torchrun on N): a subset of the 64 s-values (every `stride`-th point plus the two next
to s = 0) with the dynamic schedule, for several (workers, stagnation) settings.  Prints points/hour and the energies."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import mpo as M, scan

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl")
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 4
settings = json.loads(sys.argv[2]) if len(sys.argv) > 2 else [{"workers": 1, "stagnation": 30}, {"workers": 3, "stagnation": 30}]
s_all = np.linspace(-0.5, 0.5, 64)
idx = sorted(set(list(range(0, 64, stride)) + [31, 32]))
s_vals = s_all[idx]
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
order = [int(i) for i in np.argsort(np.abs(s_vals), kind="stable")]
for opts in settings:
    st = {}
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_vals, 256, rank, world, dist, maxIter=3, tol=1e-8, stats=st, schedule="dynamic", order=order, seed=0, **opts)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"opts": opts, "n_gpus": world, "points": len(idx), "wall_sec": round(dt, 2), "points_per_hour": round(len(idx) / dt * 3600),
                          "rank0_points": len(res["points"]), "rank0_sec": [round(v, 1) for v in res["sec"]], "rank0_matvecs": st.get("n_matvec"),
                          "E": [round(float(np.real(e)), 9) for e in res["E"]], "E_imag_max": float(np.max(np.abs(np.imag(res["E"]))))}), flush=True)
if dist is not None:
    dist.destroy_process_group()

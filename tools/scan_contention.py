"""Host-contention probe for the latency-bound scan (to run under torchrun with 1, 2, 4, 8 ranks):
every rank runs the same small-chi Davidson-heavy stage on its own GPU and reports its mat-vec cycles per
second, the host core count and its affinity.  If cycles/s per rank drops as ranks are added, the ranks are
starving each other on the host (the 8-GPU scan of round 1 ran 4.7x slower per point than one process alone).

    python -m torch.distributed.run --nproc-per-node 8 tools/scan_contention.py [--pin] [--blas1]
"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr, lw = int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("LOCAL_WORLD_SIZE", world))
if "--pin" in sys.argv and hasattr(os, "sched_setaffinity"):
    cores = sorted(os.sched_getaffinity(0)); per = max(1, len(cores) // lw)
    os.sched_setaffinity(0, set(cores[lr * per:(lr + 1) * per] or cores))
if "--blas1" in sys.argv:                       # hypothesis: spinning BLAS worker pools (one per rank, all cores each)
    import scipy.linalg  # noqa: F401  (only loaded pools can be limited)
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=1)
torch.cuda.set_device(lr)
from pydmrg_b200 import dmrg, mpo as M
mpo = M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
np.random.seed(0)
dmrg.run_dmrg(M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)), mbd=8, maxIter=0, twoSite=True)     # start-up
st = {}
torch.cuda.synchronize(); t0 = time.perf_counter()
dmrg.run_dmrg(mpo, mbd=16, tol=1e-8, maxIter=0, twoSite=True, stats=st, res_tol=1e-5)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("rank %d/%d: %d cycles in %.2f s = %.0f cycles/s | host cores %s, affinity %d, torch threads %d, OMP_NUM_THREADS=%s"
      % (rank, world, st["cycles"], dt, st["cycles"] / dt, os.cpu_count(),
         len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else -1, torch.get_num_threads(),
         os.environ.get("OMP_NUM_THREADS")), flush=True)

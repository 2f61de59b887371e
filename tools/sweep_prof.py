"""Kernel-level breakdown (torch.profiler) of ONE converged two-site sweep of the TASEP N=100 chain at chi
(development / evidence tool; output summarised in profiles/README.md)."""
import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import core, dmrg, env, mpo as M, mps as mps_tools
from torch.profiler import profile, ProfilerActivity

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 100
ctx = core.Context("c128")
mpo = M.asep_mpo(N, (0.35, 0, 1, 0, 2 / 3, 0, -0.5))
np.random.seed(0)
tmp = tempfile.mkdtemp()
sched = [c for c in (16, 64, 256) if c < chi]
dmrg.run_dmrg(mpo, mbd=sched, tol=1e-9, maxIter=1, fname=os.path.join(tmp, "st"), twoSite=True, ctx=ctx)
mpsL, _ = mps_tools.load_mps(os.path.join(tmp, "st_mbd%d" % (len(sched) - 1)))
mpsL = mps_tools.increase_all_mbd(mpsL, chi)
mpsL = [[ctx.dev(t) for t in mpsL[0]]]
F = env.calc_env(mpsL[0], mpo, chi, gaugeSite=0, ctx=ctx)
cache = {}
for _ in range(2):
    E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, recycle=cache)
torch.cuda.synchronize(); t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, recycle=cache)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev) / 1e6
print("chi=%d N=%d: sweep wall %.2f s (under the profiler), GPU busy %.2f s, %d kernel launches" % (chi, N, wall, tot, sum(e.count for e in ev)))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:22]:
    print("  %6.2f %%  %8.1f ms  x%-6d %s" % (100 * e.device_time_total / 1e6 / tot, e.device_time_total / 1e3, e.count, e.key[:100]))

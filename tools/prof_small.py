import sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import dmrg, mpo as M
mpo = M.asep_mpo(10, (0.35, 0, 1, 0, 2 / 3, 0, -0.5))
np.random.seed(0); dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10)   # warm
np.random.seed(0)
st = {}
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
r = dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10, stats=st)
pr.disable()
print("cfg0 sec", time.perf_counter() - t0, "E", r[0], "n_mv", st["n_matvec"], "cycles", st["cycles"])
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

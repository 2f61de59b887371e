"""Which TASEP s-ensemble points are entangled (so that a chi=1024 sweep does real work)?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import dmrg, mpo as M
for prm in [(0.35, 0, 1, 0, 2 / 3, 0, 0.5), (0.35, 0, 1, 0, 2 / 3, 0, 0.1), (0.35, 0, 1, 0, 2 / 3, 0, -0.1), (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)]:
    mpo = M.asep_mpo(100, prm)
    guess = None
    for chi in (16, 64, 128):
        st = {}
        np.random.seed(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = dmrg.run_dmrg(mpo, mbd=chi, tol=1e-9, maxIter=2, twoSite=True, initGuess=guess, returnState=True, returnEntSpec=True, stats=st)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        guess = out[-1]
        ees = np.sort(np.asarray(out[3]))[::-1]
        print(prm, "chi", chi, "E", complex(out[0]), "EE", float(np.real(out[1])), "%.1f s" % dt, "matvecs", st["n_matvec"], "per bond-visit %.1f" % (st["n_matvec"] / (198 * 3)),
              "min EEs", ees[-1], flush=True)

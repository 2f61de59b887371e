import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import mpo as M, scan
s_all = np.linspace(-0.5, 0.5, 64)
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
pts = [31, 32, 27, 36, 0, 63]
for opts in ({"nStates": 2, "workers": 3}, {"nStates": 2, "workers": 3, "precond": "diag"}):
    st = {}
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_all, 256, maxIter=10, tol=1e-8, stats=st, schedule="dynamic", order=pts, seed=0, **opts)
    print(opts, "%.1f s" % (time.perf_counter() - t0), "per point", [round(v, 1) for v in res["sec"]], [complex(round(res["E"][i].real, 7), round(res["E"][i].imag, 7)) for i in pts], "matvecs", st.get("n_matvec"), flush=True)

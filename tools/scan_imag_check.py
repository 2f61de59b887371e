import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import mpo as M, scan
s_all = np.linspace(-0.5, 0.5, 64)
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
idx = [0, 9, 18, 27, 36, 45, 54, 63]
s_vals = s_all[idx]
order = [int(i) for i in np.argsort(np.abs(s_vals), kind="stable")]
for opts in ({"stagnation": 30}, {"stagnation": 0}):
    st = {}
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_vals, 256, maxIter=10, tol=1e-8, stats=st, schedule="dynamic", order=order, seed=0, workers=3, **opts)
    print(opts, "%.1f s" % (time.perf_counter() - t0), [complex(round(e.real, 7), round(e.imag, 7)) for e in res["E"]], "EE", [round(float(np.real(e)), 4) for e in res["EE"]], flush=True)

"""Are the off-trend energies of a few scan points (seen at N=8: points 5, 13, 16 of 64) a property of the start / the
stagnation stop, or of running several workers?  Same seeds as the 64-point scan (seed + index)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import mpo as M, scan
s_all = np.linspace(-0.5, 0.5, 64)
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
pts = [4, 5, 6, 13, 16, 21, 31, 32]
for opts in ({"workers": 3, "stagnation": 30}, {"workers": 3, "stagnation": 60}):
    kw = dict(opts)
    mi = kw.pop("maxIter", 3)
    st = {}
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_all, 256, maxIter=mi, tol=1e-8, stats=st, schedule="dynamic", order=pts, seed=0, **kw)
    print(opts, "%.1f s" % (time.perf_counter() - t0), [round(float(np.real(res["E"][i])), 6) for i in pts], "matvecs", st.get("n_matvec"), flush=True)

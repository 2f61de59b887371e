"""Where does a cold scan point go?  Runs every 8th point of BASELINE configs[3] (ASEP N=50, chi=256, 64 s-values)
on ONE GPU and prints per-point wall time, mat-vecs, Davidson cycles and (second pass, synchronising profile) the
split solver / truncation / block update / everything else.  Run with and without OMP_NUM_THREADS=1 (torchrun sets it)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import mpo as M, scan
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s_all = np.linspace(-0.5, 0.5, 64)
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
print("OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads(), flush=True)
tot = 0.0
for prof in (False, True):
    for i in range(0, 64, stride):
        st = {"profile": True} if prof else {}
        np.random.seed(0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = scan.run_scan(mpo_of, [s_all[i]], 256, maxIter=3, tol=1e-8, stats=st)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if not prof:
            tot += dt
        print("profile" if prof else "plain  ", "point %2d s=%+.3f  %.2f s  E=%.10f  matvecs %d cycles %d" % (i, s_all[i], dt, r["E"][0].real, st.get("n_matvec", 0), st.get("cycles", 0)),
              {k: round(v, 2) for k, v in st.items() if k.startswith("t_")}, {k: v for k, v in st.items() if k.startswith("n_") and k != "n_matvec"}, flush=True)
    if not prof:
        print("mean sec/point %.2f -> %.0f points/hour/GPU" % (tot / len(range(0, 64, stride)), 3600 / (tot / len(range(0, 64, stride)))), flush=True)
    if stride < 16:
        stride = 16

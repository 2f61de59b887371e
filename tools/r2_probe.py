"""Round-2 probes on one GPU: (1) the mat-vec at chi <= 512 with and without stream-K; (2) the hard scan point next to
s = 0 (configs[3], point 32 of 64) with the reference's restart and with thick restarts."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pydmrg_b200 import device as D, mpo as M, scan

which = sys.argv[1:] or ["matvec", "hard"]
if "matvec" in which:
    for mode in (0, 1):
        D.lib.pdmrg_set_streamk_mode(mode)
        r = bench.small_chi_matvec()
        print("streamk_mode", mode, json.dumps({k: (round(v["ms"] * 1e3, 1), [round(x * 1e3, 1) for x in v["stage_ms"]], round(v["executed_gemm_tflops_whole_matvec"], 1)) for k, v in r.items()}), flush=True)
    D.lib.pdmrg_set_streamk_mode(1)
if "hard" in which:
    s_all = np.linspace(-0.5, 0.5, 64)
    mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    scan.run_scan(lambda s: M.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_all[:1], 16, ramp=(4,), maxIter=0)
    for idx in (32, 31, 0):
        for opts in ({"restart_keep": 4}, {"restart_keep": 4, "max_space": 20}, {"restart_keep": 6, "max_space": 28}) + (({},) if idx == 0 else ()):
            st = {}
            np.random.seed(0)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = scan.run_scan(mpo_of, [s_all[idx]], 256, maxIter=3, tol=1e-8, stats=st, **opts)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print("point", idx, "s=%+.4f" % s_all[idx], opts, "%.2f s" % dt, "E", complex(r["E"][0]), "EE", float(np.real(r["EE"][0])), "matvecs", st.get("n_matvec"),
                  "restarts", st.get("n_thick_restarts"), flush=True)

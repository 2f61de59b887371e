"""Run the BASELINE.json configurations end to end on one GPU and print timings
(development / evidence tool; results are summarised in profiles/README.md)."""
import json, os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, ".")
from pydmrg_b200 import core, dmrg, env, mpo as M, mps as mps_tools

which = sys.argv[1:] or ["cfg0", "cfg1", "cfg3pt"]
out = {}

SVD = os.environ.get("PDMRG_SVD", "auto")
RECYCLE = os.environ.get("PDMRG_RECYCLE", "1") != "0"      # warm-started truncation after each bond's first visit


def timed_sweeps(mpo, chi, dtype, n_sweeps, seed=0, two_site=True, sched=None, res_tol=None):
    ctx = core.Context(dtype, svd_method=SVD)
    N = len(mpo[0])
    np.random.seed(seed)
    tmp = tempfile.mkdtemp()
    t0 = time.perf_counter()
    if sched:
        r = dmrg.run_dmrg(mpo, mbd=sched, tol=1e-9, maxIter=1, fname=os.path.join(tmp, "st"), twoSite=two_site, ctx=ctx)
        mpsL, g = mps_tools.load_mps(os.path.join(tmp, "st_mbd%d" % (len(sched) - 1)))
        if not two_site or os.environ.get("PDMRG_PAD"):          # two-site sweeps grow the bonds on their own
            mpsL = mps_tools.increase_all_mbd(mpsL, chi)
    else:
        mpsL = mps_tools.make_all_mps_right(mps_tools.create_all_mps(N, chi, 1))
    t_prep = time.perf_counter() - t0
    mpsL = [[ctx.dev(t) for t in mpsL[0]]]
    F = env.calc_env(mpsL[0], mpo, chi, gaugeSite=0, ctx=ctx)
    res = []
    cache = {} if RECYCLE else None
    for s in range(n_sweeps):
        stats = {"profile": bool(os.environ.get("PDMRG_PROFILE"))}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if two_site:
            E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, stats=stats, res_tol=res_tol, recycle=cache)
        else:
            E, mpsL, F, EE, EEs = dmrg.rightSweep(mpsL, mpo, F, s, ctx=ctx, stats=stats)
            E, mpsL, F, EE, EEs = dmrg.leftSweep(mpsL, mpo, F, s, ctx=ctx, stats=stats)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res.append({"sweep": s, "sec": dt, "E": complex(E).real, "EE": float(EE), "n_matvec": stats.get("n_matvec"), "cycles": stats.get("cycles"),
                    "t_eig": stats.get("t_eig"), "t_svd": stats.get("t_svd"), "t_env": stats.get("t_env"),
                    "n_recycled": stats.get("n_recycled", 0), "n_exact_trunc": stats.get("n_exact_trunc", 0),
                    "max_bond": int(max(t.shape[2] for t in mpsL[0]))})
        print("   sweep", res[-1], flush=True)
    return {"prep_sec": t_prep, "sweeps": res}

if "cfg0" in which:
    mpo = M.asep_mpo(10, (0.35, 0, 1, 0, 2 / 3, 0, -0.5))
    np.random.seed(0); t0 = time.perf_counter(); st = {}
    r = dmrg.run_dmrg(mpo, mbd=10, tol=1e-8, maxIter=10, stats=st)
    out["cfg0"] = {"sec": time.perf_counter() - t0, "E": complex(r[0]).real, "n_matvec": st["n_matvec"]}
    print("cfg0", out["cfg0"], flush=True)
if "cfg1" in which:
    print("cfg1: Heisenberg N=100 chi=256 f64 two-site", flush=True)
    out["cfg1"] = timed_sweeps(M.heisenberg_mpo(100), 256, "f64", 5, sched=[16, 64])
if "cfg3pt" in which:
    print("cfg3 point: ASEP N=50 chi=256 c128 one s value, two-site", flush=True)
    out["cfg3pt"] = timed_sweeps(M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)), 256, "c128", 5, sched=[16, 64])
if "cfg3pt1" in which:
    print("cfg3 point: ASEP N=50 chi=256 c128 one s value, ONE-site (the reference's finite algorithm)", flush=True)
    out["cfg3pt1"] = timed_sweeps(M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)), 256, "c128", 3, two_site=False, sched=[16, 64])
if "cfg2" in which:
    print("cfg2: TASEP N=100 chi=1024 c128 two-site, one sweep after a [16,64,256] ramp", flush=True)
    out["cfg2"] = timed_sweeps(M.asep_mpo(100, (0.35, 0, 1, 0, 2 / 3, 0, -0.5)), 1024, "c128", int(os.environ.get("PDMRG_CFG2_SWEEPS", "5")),
                               sched=[16, 64, 256])
if "onesite" in which:
    print("one-site (the reference's finite algorithm): ASEP N=50, mbd=[16,64,256], 2 sweeps per stage", flush=True)
    mpo = M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))
    out["onesite"] = {}
    for rec in (False, True):
        np.random.seed(0); st = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = dmrg.run_dmrg(mpo, mbd=[16, 64, 256], tol=1e-10, maxIter=2, recycle=rec, stats=st)
        torch.cuda.synchronize()
        out["onesite"]["qr_renorm" if rec else "rdm_renorm"] = {"sec": time.perf_counter() - t0, "E": [float(np.real(e)) for e in np.atleast_1d(r[0])],
                                                               "n_matvec": st.get("n_matvec")}
        print("   ", "QR renormalisation" if rec else "RDM eigen-decomposition", out["onesite"]["qr_renorm" if rec else "rdm_renorm"], flush=True)
if "cfg4" in which:
    print("cfg4: Heisenberg N=200 chi=2048 f64 two-site (HBM-sizing stress), sweeps after a [16,64,256] ramp", flush=True)
    torch.cuda.reset_peak_memory_stats()
    out["cfg4"] = timed_sweeps(M.heisenberg_mpo(200), 2048, "f64", int(os.environ.get("PDMRG_CFG4_SWEEPS", "2")), sched=[16, 64, 256])
    out["cfg4"]["peak_hbm_gb"] = torch.cuda.max_memory_allocated() / 1e9
    print("   peak HBM %.1f GB" % out["cfg4"]["peak_hbm_gb"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/run_configs.json", "w"), indent=1)

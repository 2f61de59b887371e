"""Run under torchrun with N >= 2 ranks (one per GPU): two-site sweeps whose local mat-vecs are sharded over the
ranks (run_dmrg(..., shard=ShardSpec)) against the same sweeps run unsharded on every rank.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_sharded_sweep_check.py

First GPU runs: round 2 (profiles/sharded_sweep_r02_n2.txt, _n8.txt).
Environment: PDMRG_CHECK_CHI (default 64), PDMRG_CHECK_N (default 16), PDMRG_CHECK_TIMING=1 adds a chi=1024
timing of one converged sweep sharded vs unsharded (TASEP N=100, BASELINE configs[2])."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    emulate = os.environ.get("PDMRG_CHECK_EMULATE") == "1"      # CPU dry run of THIS script: gloo + tests/emu_device.py
    dev = "cpu" if emulate else "cuda"
    if emulate:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        os.environ["PDMRG_PY_DAVIDSON"] = "1"
        import emu_device
        emu_device.install(setattr)
        from pydmrg_b200 import core
        core.default_context = lambda dtype="c128": emu_device.Context(dtype)
        import pydmrg_b200.dmrg as _d
        _d.default_context = core.default_context
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    from pydmrg_b200 import dmrg, mpo as M, parallel
    chi = int(os.environ.get("PDMRG_CHECK_CHI", str(max(64, 16 * world))))      # ShardSpec needs 16 ket columns per rank
    N = int(os.environ.get("PDMRG_CHECK_N", "16"))
    ok = True
    report = {"world": world, "chi": chi, "N": N}
    for model, dtype in (("asep", "c128"), ("heisenberg", "f64")):
        mpo = M.asep_mpo(N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)) if model == "asep" else M.heisenberg_mpo(N)
        res = {}
        for mode in (None, "ket", "bond"):
            # the CPU dry run has no peer memory: NCCL-style path only; on GPUs "ket" runs the fused C++ solve, "bond" the
            # Python loop with one all-reduce per mat-vec
            spec = None if mode is None else parallel.ShardSpec(rank, world, dist, mode=mode, min_dim=32, fused=not emulate)
            np.random.seed(2)                                   # identical start on every rank
            st = {}
            out = dmrg.run_dmrg(mpo, mbd=chi, tol=0.0, maxIter=3, twoSite=True, dtype=dtype, stats=st, res_tol=1e-11,
                                returnEntSpec=True, shard=spec)
            res[mode] = (complex(out[0]), float(np.real(out[1])), st.get("n_sharded_solves", 0), st.get("n_matvec", 0),
                         st.get("n_sharded_env", 0), st.get("n_fused_sharded_solves", 0))
        E0, EE0 = res[None][0], res[None][1]
        for mode in ("ket", "bond"):
            E, EE, ns, nmv, nenv, nfused = res[mode]
            # every rank must hold the same numbers as rank 0
            t = torch.tensor([E.real, E.imag, EE], dtype=torch.float64, device=dev)
            t0 = t.clone(); dist.broadcast(t0, 0)
            same = bool(torch.equal(t, t0))
            good = abs(E - E0) < 1e-10 * abs(E0) and abs(EE - EE0) < 1e-8 and ns > 0 and nenv > 0 and same
            ok = ok and good
            report["%s_%s" % (model, mode)] = {"dE": abs(E - E0), "dEE": abs(EE - EE0), "sharded_solves": ns, "fused_cpp_solves": nfused, "sharded_block_updates": nenv,
                                               "n_matvec": nmv, "n_matvec_unsharded": res[None][3], "same_on_all_ranks": same}
    if os.environ.get("PDMRG_CHECK_TIMING") == "1":
        # BASELINE configs[2] at the sweep level: entangled TASEP (s = +0.5) N=100; the state is grown ONCE through the
        # chi ramp (unsharded, identical on every rank: same seed), then swept at chi = big unsharded and sharded from
        # the same start -- sweep times, mat-vec counts, E and EE side by side
        from pydmrg_b200 import core, env as env_tools
        ctx = core.Context("c128")
        big = int(os.environ.get("PDMRG_CHECK_BIG_CHI", "1024"))
        nsweeps = int(os.environ.get("PDMRG_CHECK_SWEEPS", "4"))
        mpo = M.asep_mpo(100, (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, 0.5))
        np.random.seed(0)
        guess = None
        for c, rt, mi in [(c, rt, mi) for c, rt, mi in ((16, 1e-5, 0), (64, 1e-6, 0), (256, None, 1)) if c < big]:
            out = dmrg.run_dmrg(mpo, mbd=c, tol=1e-9, maxIter=mi, twoSite=True, ctx=ctx, returnState=True, initGuess=guess, res_tol=rt)
            guess = out[-1]
        start = guess[0]
        for mode in (None, "ket"):
            spec = None if mode is None else parallel.ShardSpec(rank, world, dist, mode=mode)
            mpsL = [[ctx.dev(t) for t in start]]
            F = env_tools.calc_env(mpsL[0], mpo, big, gaugeSite=0, ctx=ctx)
            cache, ts, mv, st = {}, [], [], {}
            for _ in range(nsweeps):
                st = {}
                torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
                E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, big, ctx=ctx, stats=st, recycle=cache, shard=spec)
                torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
                mv.append(st.get("n_matvec"))
            report["sweep_sec_%s" % (mode or "unsharded")] = {"sec": [round(t, 3) for t in ts], "n_matvec": mv, "E": [E.real, E.imag], "EE": float(EE),
                                                             "sharded_solves": st.get("n_sharded_solves", 0),
                                                             "fused_cpp_solves": st.get("n_fused_sharded_solves", 0),
                                                             "sharded_block_updates": st.get("n_sharded_env", 0)}
            del mpsL, F, cache
            torch.cuda.empty_cache()
        a, b = report["sweep_sec_unsharded"], report["sweep_sec_ket"]
        report["sweep_dE"] = abs(complex(*a["E"]) - complex(*b["E"]))
        report["sweep_dEE"] = abs(a["EE"] - b["EE"])
        ok = ok and report["sweep_dE"] < 1e-10 * abs(complex(*a["E"])) and report["sweep_dEE"] < 1e-8 and b["sharded_solves"] > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["ok"] = bool(flag.item())
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()

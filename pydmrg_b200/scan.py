"""s-field / parameter scans: independent DMRG points distributed over GPUs.

The reference's scan scripts are serial continuations -- every point warm-starts from the previous
point's saved MPS (pydmrg/scripts/asep/current/bondDimConv.py:19-27, multipleStates.py:114-129).
Here the points are dealt to the ranks (one rank per GPU): every ``world``-th point for the default
cold-started scan (the cost of a point varies 3x along the scan, so this balances the ranks), or
contiguous blocks with the continuation kept INSIDE a block (``continuation=True``).  Either way there
is no data-path communication at all; results are gathered once at the end.

    torchrun --nproc-per-node 8 -m pydmrg_b200.scan --N 50 --chi 256 --points 64
"""
import time

import numpy as np

from . import mpo as mpo_lib
from .parallel import WorkQueue, gather_scan, gather_scan_dynamic, scan_indices


def run_scan(mpo_of, values, mbd, rank=0, world=1, dist=None, ramp=(16, 64), tol=1e-8, maxIter=10, minIter=0,
             twoSite=True, dtype="c128", continuation=False, ramp_res_tol=1e-5, stats=None, schedule="static", order=None,
             workers=1, seed=None, **solver):
    """Run ``run_dmrg(mpo_of(v), mbd=...)`` for every v in ``values`` (this rank's block only).

    Every point cold-starts through the chi ramp by default: the ramp stages run one sweep with a
    loose Davidson residual (``ramp_res_tol``), the target chi runs to ``tol`` with the solver's
    normal stopping rule.  ``continuation=True`` warm-starts a point from the previous point of the
    block, as the reference's scripts do -- with two-site sweeps of a single (right) eigenvector the
    Ritz value can leave the Perron branch of these non-Hermitian generators after a warm start
    (observed at N=50, chi=256), so it is opt-in.

    The default ramp (16, 64) was the best of (16,64), (64,), (32,), (8,32,128), (16,64,128) at N=50,
    chi=256 (skipping the chi=16 stage doubles the mat-vec count at s=-0.5); capping the Davidson cycles
    per bond in the ramp was tried and rejected (the final stage then needs 15-40x more mat-vecs and can
    settle on the wrong branch of the non-Hermitian generator).

    ``schedule='dynamic'`` (cold starts only): the points are not dealt out in advance; a rank takes the next point of
    ``order`` (default: as given; pass the expected-most-expensive first) whenever it is free (parallel.WorkQueue: one
    atomic counter in the process group's store, nothing on the data path).

    ``workers`` > 1 (cold starts; implies the dynamic schedule): that many host threads of this rank each drive their own
    scan point on their own CUDA stream of the SAME GPU.  Most of a cold point is latency-bound (bond dimensions 2..64
    in the chi ramp and at the chain ends: a Davidson cycle is ~12 tiny kernels and a host round trip), so one point
    leaves the GPU idle most of the time; concurrent points fill it.  The C-ABI calls release the GIL, the library keeps
    its cuSOLVER / cuBLAS handles, pinned staging and stream-K scratch per thread / per stream.
    ``seed``: the random start of point i is drawn from np.random.seed(seed + i), or seed[i] when a sequence is given
    (reproducible whatever the schedule);
    default: the global NumPy state as the reference uses it (only reproducible with one worker and a static schedule).

    Returns dict(E, EE: full-length complex arrays (gathered), sec: per-point wall seconds of this
    rank's points, points: their indices)."""
    from .dmrg import run_dmrg
    n = len(values)
    interleave = not continuation          # cold starts: every world-th point (balanced); continuation: blocks
    workers = 1 if continuation else max(1, int(workers))      # a continuation is serial by definition
    solver = dict(solver)
    user_ctx = solver.pop("ctx", None)                           # one worker: the caller's context; several: one per worker
    if schedule not in ("static", "dynamic"):
        raise ValueError("schedule must be 'static' or 'dynamic'")
    dynamic = (schedule == "dynamic" or workers > 1) and not continuation
    if dynamic:
        todo = WorkQueue(list(range(n)) if order is None else [int(i) for i in order], rank, world, dist)
    else:
        todo = scan_indices(n, rank, world, interleave)
    done = []                                                    # (index, E, EE, seconds) of this rank
    import threading
    seed_lock = threading.Lock()

    def first_guess(idx, chi):
        """Random right-canonical start of point idx (dmrg.py:363-366), drawn under a lock from a per-point seed."""
        if seed is None:
            return None
        from . import mps as mps_tools
        with seed_lock:
            np.random.seed(int(seed[idx]) if np.ndim(seed) else int(seed) + int(idx))
            return mps_tools.make_all_mps_right(mps_tools.create_all_mps(len(mpo_of(values[idx])[0]), chi, int(solver.get("nStates", 1))))

    def solve_points(points, ctx, st):
        prev = None
        for idx in points:
            mpo = mpo_of(values[idx])
            t0 = time.perf_counter()
            if prev is None or not continuation:
                # cold start: chi ramp with in-memory continuation between stages
                stages = [c for c in ramp if c < mbd] + [mbd]
                guess = first_guess(idx, stages[0])
                for chi in stages:
                    last = chi == mbd
                    opts = dict(solver)
                    if not last and ramp_res_tol is not None:
                        opts["res_tol"] = ramp_res_tol
                    if not last:
                        # the stagnation stop is for the final stage only: cutting local solves short in the cold ramp
                        # sent 3 of 64 points of configs[3] to a metastable state (E off by 3-10 %), while the ramp's
                        # cycles are the cheap, latency-bound ones anyway
                        opts.pop("stagnation", None)
                    out = run_dmrg(mpo, mbd=chi, tol=tol, maxIter=maxIter if last else 0, minIter=minIter,
                                   initGuess=guess, twoSite=twoSite, dtype=dtype, returnState=True, stats=st, ctx=ctx, **opts)
                    # one-site runs leave the centre at gaugeSiteSave = N//2+1 (dmrg.py:322), two-site ones at site 0
                    guess = out[-1] if twoSite else (_grow(out[-1], min([c for c in list(ramp) + [mbd] if c > chi] or [mbd])),
                                                     len(mpo[0]) // 2 + 1)
            else:
                out = run_dmrg(mpo, mbd=mbd, tol=tol, maxIter=maxIter, minIter=minIter, initGuess=prev, twoSite=twoSite,
                               dtype=dtype, returnState=True, stats=st, ctx=ctx, **solver)
            done.append((idx, out[0], out[1], time.perf_counter() - t0))
            prev = out[-1] if twoSite else (out[-1], len(mpo[0]) // 2 + 1)

    if workers == 1:
        solve_points(todo, user_ctx, stats)
    else:
        import torch
        from .core import Context
        dev_index = torch.cuda.current_device()
        errors, all_stats = [], []

        def worker():
            try:
                torch.cuda.set_device(dev_index)
                with torch.cuda.stream(torch.cuda.Stream(device=dev_index)):
                    st = {} if stats is not None else None
                    all_stats.append(st)
                    solve_points(todo, Context(dtype), st)
                    torch.cuda.current_stream().synchronize()
            except BaseException as exc:                         # surface the first failure in the caller
                errors.append(exc)

        threads = [threading.Thread(target=worker, name="pdmrg-scan-%d" % k) for k in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        if stats is not None:
            for st in all_stats:
                for key, val in (st or {}).items():
                    if isinstance(val, (int, float)) and key != "last_residual":
                        stats[key] = stats.get(key, 0) + val
    done.sort(key=lambda r: r[0])
    mine = [r[0] for r in done]
    E_loc, EE_loc, secs = [r[1] for r in done], [r[2] for r in done], [r[3] for r in done]
    if dynamic:
        return {"E": gather_scan_dynamic(mine, E_loc, n, world, dist), "EE": gather_scan_dynamic(mine, EE_loc, n, world, dist),
                "sec": secs, "points": mine}
    return {"E": gather_scan(E_loc, n, rank, world, dist, interleave), "EE": gather_scan(EE_loc, n, rank, world, dist, interleave),
            "sec": secs, "points": mine}


def scgf_shape_ok(s_vals, E, slack=1e-6):
    """Self-check of an s-scan: the leading eigenvalue of the tilted generator is a scaled cumulant generating function,
    hence monotone and convex in s.  A point whose sweeps got stuck in a metastable state shows up as a kink (seen with a
    sweep cap of 4: 3 of 64 points of configs[3] were off by 3-10 %); False then."""
    s = np.asarray(s_vals, dtype=float)
    e = np.real(np.asarray(E))
    o = np.argsort(s)
    s, e = s[o], e[o]
    if len(s) < 3:
        return True
    d1 = np.diff(e) / np.diff(s)
    mono = np.all(np.diff(e) <= slack) or np.all(np.diff(e) >= -slack)
    return bool(mono and np.all(np.diff(d1) >= -1e-3 * (1.0 + np.abs(d1[:-1]))))


def _grow(mpsList, chi):
    from . import mps as mps_tools
    return mps_tools.increase_all_mbd([[np.array(t) for t in M] for M in mpsList], chi)


def main():
    import argparse, json, os
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=50)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--points", type=int, default=64)
    ap.add_argument("--smin", type=float, default=-0.5)
    ap.add_argument("--smax", type=float, default=0.5)
    ap.add_argument("--max-iter", type=int, default=10, help="sweep cap at the target chi (dmrg.py:309 default 10); tol ends a point earlier")
    ap.add_argument("--pin-cores", action="store_true",
                    help="give every local rank its own slice of the host cores (the scan is host-latency bound)")
    ap.add_argument("--precond", default=None, choices=[None, "diag"], help="Davidson preconditioner (default: the reference's identity)")
    ap.add_argument("--max-space", type=int, default=12, help="Davidson subspace before a restart (reference: 12)")
    ap.add_argument("--workers", type=int, default=3, help="concurrent scan points per GPU (host threads, one CUDA stream each)")
    ap.add_argument("--restart-keep", type=int, default=0, help="thick restart of the Davidson subspace (0: the reference's restart)")
    ap.add_argument("--stagnation", type=int, default=0, help="final chi stage: stop a local solve whose residual was not halved within this "
                                                              "many cycles (0 = the reference's behaviour, up to 1000 cycles per solve; 30 makes the "
                                                              "two points next to s=0 12x cheaper but sent 1 of 8 test points to a complex branch)")
    ap.add_argument("--schedule", default="dynamic", choices=["static", "dynamic"],
                    help="dynamic: a rank takes the next point when it is free, points nearest s = 0 (the expensive ones) first")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.pin_cores and hasattr(os, "sched_setaffinity"):
        cores = sorted(os.sched_getaffinity(0))
        lw, lr = int(os.environ.get("LOCAL_WORLD_SIZE", world)), int(os.environ.get("LOCAL_RANK", 0))
        per = max(1, len(cores) // max(1, lw))
        mine = cores[lr * per:(lr + 1) * per] or cores
        os.sched_setaffinity(0, set(mine))
    blas_threads = None
    if world > 1:
        # NumPy / SciPy BLAS pools default to ALL host cores in every rank and their idle workers spin before they
        # sleep: eight ranks of a latency-bound scan would oversubscribe the host.  Give each rank its share.
        try:
            import scipy.linalg  # noqa: F401  (load SciPy's BLAS too: only loaded pools can be limited)
            from threadpoolctl import threadpool_limits
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            blas_threads = max(1, (os.cpu_count() or 1) // max(1, lw))
            threadpool_limits(limits=blas_threads)
        except Exception:
            blas_threads = None
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    s_vals = np.linspace(args.smin, args.smax, args.points)        # scripts/asep/current/bondDimConv.py:10
    mpo_of = lambda s: mpo_lib.asep_mpo(args.N, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    # library start-up (cuSOLVER / cuBLAS handles, allocator, NCCL) is paid once per process: keep it out of the clock
    run_scan(lambda s: mpo_lib.asep_mpo(8, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s))), s_vals[:1], 16, ramp=(4,), maxIter=0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    stats = {}
    order = [int(i) for i in np.argsort(np.abs(s_vals), kind="stable")]        # the gap closes at s = 0: those points first
    res = run_scan(mpo_of, s_vals, args.chi, rank, world, dist if world > 1 else None, maxIter=args.max_iter, stats=stats,
                   precond=args.precond, max_space=args.max_space, schedule=args.schedule, order=order, workers=args.workers,
                   seed=0, restart_keep=args.restart_keep, stagnation=args.stagnation)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    per_rank = [{"rank": rank, "sec": round(sum(res["sec"]), 3), "points": len(res["points"]), "cycles": stats.get("cycles"),
                 "n_matvec": stats.get("n_matvec")}]
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
    if rank == 0:
        print(json.dumps({"metric": "s_points_per_hour", "value": args.points / dt * 3600.0, "n_gpus": world,
                          "host_cores": os.cpu_count(), "cores_per_rank": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None,
                          "points": args.points, "N": args.N, "chi": args.chi, "wall_sec": dt, "per_rank": per_rank,
                          "precond": args.precond, "max_space": args.max_space, "schedule": args.schedule, "workers": args.workers,
                          "restart_keep": args.restart_keep, "stagnation": args.stagnation, "blas_threads_per_rank": blas_threads,
                          "E_first": [res["E"][0].real, res["E"][0].imag], "E_last": [res["E"][-1].real, res["E"][-1].imag],
                          "E_real": [round(float(np.real(e)), 9) for e in res["E"]],
                          "E_monotone_and_convex_in_s": scgf_shape_ok(s_vals, res["E"])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

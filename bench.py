#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native DMRG sweep hot path.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one
GPU): TASEP open chain N=100, chi=1024, d=2, MPO bond w=6, complex128, two-site local problem
at a full-chi bond.  A *step* is one pass of the hot path's dominant operation over that
problem: one two-site H_eff mat-vec  y = -L.W.W.R.x  (the reference's ``Hfun`` closure,
pydmrg/tools/diag_tools.py:146-162) on seeded synthetic L, R, x.

    metric  heff_matvec_fp64_gflops : algorithmic dense-W FLOPs (4.17e11 per mat-vec at
            chi=1024) / time; ``value`` = device-resident, ``e2e`` = through the
            reference-shaped ``Hfun(ndarray) -> ndarray`` closure with pinned host buffers
            (H2D of x and D2H of y inside the timed region).
    extras  ``roofline`` (dominant kernel = the DMMA GEMM, both stages timed live with CUDA
            events), ``cpu_baseline`` (oracle port on the host cores), ``bond_step`` (Davidson +
            truncation + block update of one chi=1024 bond), ``sweep_sec`` (measured two-site
            sweep of the N=100 chain; ``--no-full-sweep`` skips it) and ``sweep_sec_derived``
            (the same from bond steps with a fixed Davidson cycle count).

N > 1 (torchrun, one rank per GPU): each rank owns an independent s-field scan point of the
same shape (weak scaling, no data-path collective); the MPO-bond-sharded mat-vec with one
NCCL all-reduce per mat-vec is measured in addition and reported under ``sharded``.

``--impl reference`` times the oracle port of the reference's CPU implementation of the same
closure (single-threaded un-optimised einsum, exactly the calls the reference makes) on a
bounded sample; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASEP = (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, -0.5)
FP64_PEAK_FALLBACK_TFLOPS = 36.9      # profiles/fp64_peaks_r01.json (cuBLAS zgemm 8192^3 burst, this pool)


def fp64_peak():
    """FP64 tensor (DMMA) denominator.  MEASURED_PEAKS.json (driver-written) carries only HBM and
    bf16 numbers, so the FP64 figure is the one measured on this pool's B200 by this repo
    (tools/fp64_peaks.py: cuBLAS Zgemm 8192^3; tools/dmma_peak.cu: raw DMMA issue rate)."""
    for name in ("MEASURED_PEAKS.json", os.path.join("profiles", "fp64_peaks_r01.json")):
        p = os.path.join(ROOT, name)
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                for k in ("fp64_tflops", "zgemm_8192_tflops_burst"):
                    if k in d:
                        return float(d[k]), "%s:%s" % (name, k)
            except Exception:
                pass
    return FP64_PEAK_FALLBACK_TFLOPS, "fallback constant (profiles/fp64_peaks_r01.json value)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def heff2_flops(chi, d=2, w=6, cplx=True):
    """Dense-W algorithmic FLOPs of one two-site mat-vec."""
    return (8.0 if cplx else 2.0) * (2 * d * d * w * chi ** 3 + 2 * d ** 3 * w * w * chi ** 2)


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """Oracle port of the reference's CPU path; bounded sample (chi_sample) of the same closure."""
    if rank != 0:
        return
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    chi = args.ref_chi
    W1, W2 = M.asep_mpo(6, TASEP)[0][2:4]
    rng = np.random.default_rng(0)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    L, R, x = cr(chi, 6, chi), cr(chi, 6, chi), cr(2, 2, chi, chi)
    f = lambda: orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "einsum")
    for _ in range(min(args.warmup, 1)):
        f()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f()
    dt = (time.perf_counter() - t0) / args.steps
    fl = heff2_flops(chi)
    val = fl / dt * 1e-9
    # BLAS-fair variant (same contractions through threaded zgemm), reported alongside
    tb = time.perf_counter(); orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas"); tb = time.perf_counter() - tb
    line = {"impl": "reference", "metric": "heff_matvec_fp64_gflops", "value": val, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": "TASEP N=100 chi=1024 two-site H_eff mat-vec (d=2,w=6,c128)", "chi": 1024,
                       "sample_chi": chi},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": 1, "kind": "port",
                             "sample": "two-site Hfun, un-optimised np.einsum exactly as diag_tools.py:153-156, chi=%d "
                                       "(c_einsum is single-threaded and not BLAS; GFLOP/s is size-insensitive), %d steps"
                                       % (chi, args.steps),
                             "blas_fair_gflops": fl / tb * 1e-9, "blas_fair_cores": os.cpu_count()},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--ref-chi", type=int, default=256)
    ap.add_argument("--davidson-cycles", type=int, default=8)
    ap.add_argument("--no-extras", action="store_true", help="skip bond-step / cpu-baseline extras")
    ap.add_argument("--no-full-sweep", action="store_true", help="skip the measured two-site sweeps of the N=100 chain (~1 min)")
    ap.add_argument("--svd", default="auto", choices=["auto", "gesvd", "gesvdj", "gram"])
    ap.add_argument("--gauge-shortcut", action="store_true",
                    help="N = 1: also time the converged sweeps with run_dmrg's gauge_shortcut option (sweep_sec_shortcut); off by "
                         "default until validated on a GPU")
    ap.add_argument("--sharded-sweep", action="store_true",
                    help="N > 1: also time two-site sweeps of the N=100 chain with the mat-vecs and block updates of the large "
                         "bonds sharded over the ranks (run_dmrg(shard=...)); off by default until validated on GPUs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pydmrg_b200 import core, device as D, mpo as M, parallel
    from pydmrg_b200._lib import launch_count

    chi, d, w = args.chi, 2, 6
    ctx = core.Context("c128", svd_method=args.svd)
    # scan point of this rank (weak scaling: same shape, different s)
    s_val = TASEP[6] + 0.05 * rank
    prm = TASEP[:6] + (s_val,)
    W1, W2 = M.asep_mpo(6, prm)[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.complex128)
    g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
    rnd = lambda *s: torch.randn(*s, dtype=torch.complex128, device="cuda", generator=g)
    L, R = rnd(chi, w, chi), rnd(chi, w, chi)
    x = rnd(d * d * chi * chi)
    y = torch.empty_like(x)
    plan = D.HeffPlan(D.C128, d * d, chi, chi, [(L, Wc, R)], ctx.ws)
    flops = heff2_flops(chi)
    assert abs(flops - plan.flops_dense) < 1e-6 * flops

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        plan.apply(x, y)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.apply(x, y)
    e1.record()
    barrier()
    launches = launch_count() - n0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * flops / (ms_step * 1e-3) * 1e-9

    # ---- per-stage timing of the same step (roofline of the dominant kernel) -----------------
    st = np.zeros(3)
    nrep = max(3, min(args.steps, 10))
    for _ in range(nrep):
        st += np.array(plan.apply_timed(x, y))
    st /= nrep
    # FLOPs the two GEMM launches actually execute: chi^3 terms of the non-empty MPO slabs (for TASEP
    # rows 3,4 of W are zero, so stage C runs 4 of 6 slabs); complex MAC = 8 FLOP
    cube = plan.flops_gemm
    gemm_ms = st[0] + st[2]
    peak, peak_src = fp64_peak()
    achieved = cube / (gemm_ms * 1e-3) * 1e-12
    roofline = {"bound": "tensor", "kernel": "gemm_nt_dmma_kernel<c128> (FP64 DMMA.8x8x4, TMA-fed)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src + " (FP64 tensor pipe; measured on this pool)",
                "launches_per_step": 2, "flops_per_launch_avg": cube / 2, "ms_per_launch_avg": gemm_ms / 2,
                "note": "achieved counts EXECUTED GEMM FLOPs (MPO block sparsity skipped); the headline value "
                        "counts the dense-W algorithmic FLOPs of SURVEY 8d",
                "dense_algorithmic_flops_per_step": flops, "executed_gemm_flops_per_step": cube,
                "stage_ms": {"gemm_stage_A": st[0], "mix": st[1], "gemm_stage_C": st[2]},
                "share_of_step": gemm_ms / ms_step}
    tr = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_avg")
        except Exception:
            pass

    # ---- e2e: the reference-shaped closure with HOST buffers ---------------------------------
    xh = torch.empty(x.shape, dtype=x.dtype).pin_memory()
    yh = torch.empty(x.shape, dtype=x.dtype).pin_memory()
    xh.copy_(x)
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        x.copy_(xh, non_blocking=True)
        plan.apply(x, y)
        yh.copy_(y, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    nbytes = x.numel() * 16
    e2e = {"value": world * flops / (e2e_ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world, "ms_per_step": e2e_ms,
           "api": "Hfun(x_host) -> y_host  (make_ham_func closure; pinned buffers; L,R,W bound at closure creation)"}

    line = {"metric": "heff_matvec_fp64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": "TASEP N=100 chi=%d two-site H_eff mat-vec (d=2,w=6,c128); BASELINE configs[2]" % chi,
                       "chi": chi, "N": 100, "d": d, "w": w, "flops_per_step": flops,
                       "l2": "operands 96+96+64 MiB and the 384 MiB intermediate exceed the 126 MB L2; no flush needed",
                       "parallelism": "1 scan point per GPU" if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    # ---- MPO-bond-sharded mat-vec with one NCCL all-reduce (strong scaling), N > 1 --------------
    if world > 1:
        torch.manual_seed(99)    # identical operands on every rank
        gs = torch.Generator(device="cuda"); gs.manual_seed(99)
        rs = lambda *s: torch.randn(*s, dtype=torch.complex128, device="cuda", generator=gs)
        Ls, Rs, xs = rs(chi, w, chi), rs(chi, w, chi), rs(d * d * chi * chi)
        W1s, W2s = M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5))[0][2:4]
        Wcs = D.combine_twosite(W1s, W2s, np.complex128)
        sh = parallel.ShardedHeff(D.C128, d * d, chi, chi, [(Ls, Wcs, Rs)], ctx.ws, rank, world, dist)
        ys = torch.empty_like(xs)
        for _ in range(3):
            sh.apply(xs, ys)
        barrier()
        e0.record()
        for _ in range(args.steps):
            sh.apply(xs, ys)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per, ideal = parallel.partition_cost(sh.partition[0])
        line["sharded"] = {"what": "one chi=%d two-site mat-vec (dense ASEP W) split over the MPO bond" % chi,
                           "nccl_allreduce": {"ms_per_matvec": float(t.item()), "gflops": flops / (t.item() * 1e-3) * 1e-9},
                           "scaling": "strong", "slab_gemms_per_rank": per, "ideal_speedup": ideal}
        # same decomposition, exchange fused into the kernels over peer memory (no NCCL on the data path)
        fz = parallel.FusedShardedHeff(D.C128, d * d, chi, chi, [(Ls, Wcs, Rs)], ctx.ws, rank, world, dist)
        for _ in range(3):
            fz.apply(xs, ys)
        barrier()
        e0.record()
        for _ in range(args.steps):
            fz.apply(xs, ys)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["sharded"]["fused_peer_memory"] = {"ms_per_matvec": float(t.item()), "gflops": flops / (t.item() * 1e-3) * 1e-9,
                                                "how": "GEMM scatter epilogue (peer st.global) + flag-guarded reduce/all-gather kernel"}
        fz.close()
        # ket-index sharding: every rank does 1/N of both GEMM stages, one NCCL all-reduce
        kz = parallel.ShardedHeff(D.C128, d * d, chi, chi, [(Ls, Wcs, Rs)], ctx.ws, rank, world, dist, mode="ket")
        for _ in range(3):
            kz.apply(xs, ys)
        barrier()
        e0.record()
        for _ in range(args.steps):
            kz.apply(xs, ys)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["sharded"]["ket_split_nccl"] = {"ms_per_matvec": float(t.item()), "gflops": flops / (t.item() * 1e-3) * 1e-9,
                                             "how": "left ket index split N ways (balanced), 1 NCCL all-reduce of y"}
        # the same split with the exchange fused into the kernels (long-K stage C, scatter epilogue, peer stores)
        fk = parallel.FusedShardedHeff(D.C128, d * d, chi, chi, [(Ls, Wcs, Rs)], ctx.ws, rank, world, dist, mode="ket")
        for _ in range(3):
            fk.apply(xs, ys)
        barrier()
        e0.record()
        for _ in range(args.steps):
            fk.apply(xs, ys)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["sharded"]["ket_split_fused"] = {"ms_per_matvec": float(t.item()), "gflops": flops / (t.item() * 1e-3) * 1e-9,
                                              "how": "ket split; stage C = one long-K GEMM whose scatter epilogue reduce-scatters the "
                                                     "partial y over NVLink, then a flag-guarded reduce + all-gather kernel (no NCCL)"}
        fk.close()

    if world > 1 and args.sharded_sweep:
        spec = parallel.ShardSpec(rank, world, dist, mode="ket")
        line["sweep_sec_sharded"] = full_sweep(args, ctx, shard=spec)
        line["sweep_sec_sharded"]["how"] += "; every rank runs the sweep, mat-vecs and block updates of bonds >= %d split %d ways" % (
            spec.min_dim, world)

    # ---- extras on rank 0 at N=1: bond step, sweep estimate, CPU baseline ----------------------
    if world == 1 and not args.no_extras:
        line["bond_step"] = bond_step(args, ctx, plan, L, R, W1, W2, chi)
        bs = line["bond_step"]
        nb = 2 * 99
        line["sweep_sec_derived"] = {"value": bs["ms_total_recycled"] * nb * 1e-3,
                                     "how": "198 bond steps x measured chi=%d bond step with %d Davidson cycles each and the "
                                            "warm-started truncation; upper bound (edge bonds are smaller, a converged sweep "
                                            "needs 1 cycle per bond)" % (chi, args.davidson_cycles)}
        line["krylov"] = krylov_bandwidth(chi)
        line["config4_matvec"] = config4_matvec()
        line["cpu_baseline"] = cpu_baseline(args)
        line["scan"] = scan_sample()
        if not args.no_full_sweep:
            line["sweep_sec"] = full_sweep(args, ctx)
            if args.gauge_shortcut:
                line["sweep_sec_shortcut"] = full_sweep(args, ctx, gauge_shortcut=True)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bond_step(args, ctx, plan, L, R, W1, W2, chi):
    """One two-site bond step at full chi: Davidson (fixed cycles) -> SVD truncation -> block update."""
    import torch
    from pydmrg_b200 import device as D
    from pydmrg_b200.davidson import davidson
    from pydmrg_b200.truncate import svd_truncate, truncate_twosite
    d = 2
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    x0 = torch.randn(d * d * chi * chi, dtype=torch.complex128, device="cuda", generator=g)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {}
    ctx_ref = None
    if args.svd in ("auto", "gram"):         # also time the LAPACK-grade back-end once, for the record
        from pydmrg_b200 import core
        ctx_ref = core.Context("c128", svd_method="gesvd")
    for rep in range(2):                     # first repetition warms cuSOLVER / workspaces
        stats = {}
        a, b, c, e = ev(), ev(), ev(), ev()
        a.record()
        _, vecs = davidson(lambda xx, yy: plan.apply(xx, yy, -1.0), [x0], plan.n, D.C128, ctx.device, nroots=1,
                           fixed_cycles=args.davidson_cycles, stats=stats)
        b.record()
        cache = {}
        A, B, S, EE, EEs = truncate_twosite(vecs[0], (d, d, chi, chi), chi, ctx, "right", recycle=cache, site=0,
                                            warm=None, exact=True)                          # a bond's first visit
        c.record()
        newL = D.env_update_right(A, W1.astype(np.complex128), L, None, ctx.ws)
        e.record()
        # a bond is warm-started only after it was decomposed exactly twice at its current shape (or once with
        # nothing discarded); the random test vector discards a lot, so give it its second exact visit (untimed)
        truncate_twosite(vecs[0], (d, d, chi, chi), chi, ctx, "right", recycle=cache, site=0, warm=None, exact=True)
        f, h = ev(), ev()
        st2 = {}
        f.record()
        truncate_twosite(vecs[0], (d, d, chi, chi), chi, ctx, "left", recycle=cache, site=0, warm=(A, B), stats=st2)   # a later visit
        h.record()
        torch.cuda.synchronize()
        out = {"davidson_ms": a.elapsed_time(b), "n_matvec": stats["n_matvec"], "svd_truncate_ms": b.elapsed_time(c),
               "recycled_truncate_ms": f.elapsed_time(h) if st2.get("n_recycled") else None,
               "env_update_ms": c.elapsed_time(e), "ms_total": a.elapsed_time(e),
               "ms_total_recycled": a.elapsed_time(b) + f.elapsed_time(h) + c.elapsed_time(e),
               "truncation": "first visit of a bond: projective (pdmrg_left_basis: Gram GEMM + heevd, gated tail refinement); "
                             "later visits: warm-started subspace step (pdmrg_recycle_basis)" if args.svd == "auto" else args.svd}
    if ctx_ref is not None:
        a, b = ev(), ev()
        svd_truncate(vecs[0], (d, d, chi, chi), chi, ctx_ref, absorb="right")
        a.record()
        svd_truncate(vecs[0], (d, d, chi, chi), chi, ctx_ref, absorb="right")
        b.record()
        torch.cuda.synchronize()
        out["svd_truncate_ms_gesvd"] = a.elapsed_time(b)
    return out


def krylov_bandwidth(chi, m=8, reps=20):
    """HBM roofline of the Davidson vector kernels (csrc/krylov.cu) at the headline size: vectors of
    n = d^2 chi^2 complex128 elements, a subspace of m basis vectors + m images (2 x m x 64 MiB at chi=1024,
    far beyond the 126 MB L2, so every pass streams from HBM).  achieved = ALGORITHMIC bytes (each vector
    touched once per pass) / CUDA-event time; peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth)."""
    import torch
    from pydmrg_b200 import device as D
    n = 4 * chi * chi
    B = 16 * n
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json:hbm_gbs"
    except Exception:
        peak, src = 6500.0, "fallback (B200_PROFILING.md copy bandwidth)"
    ops = D.VecOps(D.C128, n, torch.device("cuda"))
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    XS = torch.randn(m, n, dtype=torch.complex128, device="cuda", generator=g)
    AX = torch.randn(m, n, dtype=torch.complex128, device="cuda", generator=g)
    y = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
    r = torch.empty_like(y)
    coef = (np.arange(1, m + 1) / m).astype(np.complex128)
    cdev = ops.dots(XS, m, y).clone()
    cases = {
        "dots (m inner products <X_i|y>)": (lambda: ops.dots(XS, m, y), (m + 1) * B),
        "combine (Ritz vector = sum c_i X_i)": (lambda: ops.combine(coef, XS, m, r), (m + 1) * B),
        "residual (r = sum v_i (AX_i - e X_i), fused norm)": (lambda: ops.residual(coef, 0.5, XS, AX, m, r), (2 * m + 1) * B),
        "project (r -= sum c_i X_i, fused norm)": (lambda: ops.project(XS, m, cdev, r), (m + 2) * B),
        "normalize": (lambda: ops.normalize(y, ops.norm2, r), 2 * B),
    }
    out = {"n": n, "m": m, "peak_gbs": peak, "peak_source": src, "kernels": {}}
    for name, (f, nbytes) in cases.items():
        for _ in range(3):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = nbytes / (ms * 1e-3) * 1e-9
        out["kernels"][name] = {"ms": ms, "bytes": nbytes, "achieved_gbs": gbs, "frac": gbs / peak}
    return out


def config4_matvec(chi=2048, reps=10):
    """BASELINE configs[4] probe: the Heisenberg (w=5, real f64) two-site mat-vec at chi=2048 on ONE GPU
    (operands 2 x 168 MB blocks + 2 x 671 MB intermediates: the N=200 chain's 34 GB of blocks fit one B200)."""
    import torch
    from pydmrg_b200 import core, device as D, mpo as M
    ctx = core.Context("f64")
    W1, W2 = M.heisenberg_mpo(6)[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.float64)
    w = W1.shape[0]
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
    L, R, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(4 * chi * chi)
    y = torch.empty_like(x)
    plan = D.HeffPlan(D.F64, 4, chi, chi, [(L, Wc, R)], ctx.ws)
    for _ in range(3):
        plan.apply(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        plan.apply(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * (2 * 4 * w * chi ** 3 + 2 * 8 * w * w * chi ** 2)
    out = {"workload": "Heisenberg w=%d f64 two-site H_eff mat-vec, chi=%d" % (w, chi), "ms": ms,
           "dense_flops": flops, "tflops_dense_count": flops / (ms * 1e-3) * 1e-12,
           "executed_gemm_tflops": plan.flops_gemm / (ms * 1e-3) * 1e-12}
    plan.close()
    return out


def cpu_baseline(args):
    """Oracle port on this box's host cores: (A) the reference's own un-optimised einsum calls
    (single-threaded), (B) the BLAS-fair restatement on all cores.  Bounded: ~10-20 s."""
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    W1, W2 = M.asep_mpo(6, TASEP)[0][2:4]
    rng = np.random.default_rng(0)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    chi_a = args.ref_chi
    L, R, x = cr(chi_a, 6, chi_a), cr(chi_a, 6, chi_a), cr(2, 2, chi_a, chi_a)
    t = time.perf_counter(); n = 0
    while n < 3 or (time.perf_counter() - t < 5 and n < 10):
        orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "einsum"); n += 1
    ta = (time.perf_counter() - t) / n
    chi_b = min(args.chi, 1024)
    L, R, x = cr(chi_b, 6, chi_b), cr(chi_b, 6, chi_b), cr(2, 2, chi_b, chi_b)
    orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas")
    t = time.perf_counter()
    orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas")
    tb = time.perf_counter() - t
    return {"value": heff2_flops(chi_a) / ta * 1e-9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "sample": "reference-faithful two-site Hfun (np.einsum, no optimize: single-threaded c_einsum) at chi=%d, %d calls"
                      % (chi_a, n),
            "blas_fair": {"value": heff2_flops(chi_b) / tb * 1e-9, "unit": "GFLOP/s", "cores": os.cpu_count(),
                          "sample": "same contractions via np.tensordot (threaded zgemm) at chi=%d, 1 call" % chi_b}}


def scan_sample():
    """s-points/hour on one GPU from a bounded sample of BASELINE configs[3] (ASEP N=50, chi=256 two-site,
    64 s-values in [-0.5, 0.5]): four points spread over the range (the cost of a point varies 3x with s: the
    s = -0.5 end needs ~22k mat-vecs, the middle ~4.5k), each cold-started through a [16, 64] chi ramp."""
    import torch
    from pydmrg_b200 import mpo as M, scan
    idx = [0, 21, 42, 63]
    s_vals = np.linspace(-0.5, 0.5, 64)[idx]
    mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
    np.random.seed(0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_vals, 256, maxIter=3, tol=1e-8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    per = float(np.mean(res["sec"]))
    return {"config": "ASEP N=50 chi=256 two-site, points %s of 64 s-values, each cold through a [16,64] chi ramp, tol 1e-8" % idx,
            "s": [float(v) for v in s_vals], "sec_per_point": res["sec"], "s_points_per_hour_per_gpu": 3600.0 / per,
            "E": [float(np.real(e)) for e in res["E"]], "E_imag_max": float(np.max(np.abs(np.imag(res["E"]))))}


def full_sweep(args, ctx, shard=None, gauge_shortcut=False):
    """Measured two-site sweeps (right + left over all 99 bonds) of the TASEP N=100 chain at chi: the state is
    grown through a [16, 64, 256] chi ramp (so it is physical) and then swept five times with mbd = chi.  Two-site
    sweeps are not zero-padded: every bond grows by up to a factor d per visit, so the first two sweeps run on
    the growing tensors (256 -> 512 -> 1024), the third is the first at full size (every bond decomposed from
    scratch) and the later ones are converged sweeps with the warm-started truncation (centre bond exact).
    ``value`` is the last sweep."""
    import tempfile
    import torch
    from pydmrg_b200 import dmrg, env, mpo as M, mps as mps_tools
    N, chi = 100, args.chi
    mpo = M.asep_mpo(N, TASEP)
    np.random.seed(0)
    tmp = tempfile.mkdtemp()
    sched = [c for c in (16, 64, 256) if c < chi]
    t0 = time.perf_counter()
    extra = {} if shard is None else {"shard": shard}
    if gauge_shortcut:
        extra["gauge_shortcut"] = True
    dmrg.run_dmrg(mpo, mbd=sched, tol=1e-9, maxIter=1, fname=os.path.join(tmp, "st"), twoSite=True, ctx=ctx, **extra)
    mpsL, _ = mps_tools.load_mps(os.path.join(tmp, "st_mbd%d" % (len(sched) - 1)))
    mpsL = [[ctx.dev(t) for t in mpsL[0]]]
    F = env.calc_env(mpsL[0], mpo, chi, gaugeSite=0, ctx=ctx)
    torch.cuda.synchronize()
    ramp = time.perf_counter() - t0
    cache, secs, bonds, E, EE, stats = {}, [], [], None, None, {}
    for _ in range(5):
        stats = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, stats=stats, recycle=cache, **extra)
        torch.cuda.synchronize(); secs.append(time.perf_counter() - t0)
        bonds.append(int(max(t.shape[2] for t in mpsL[0])))
    return {"value": secs[-1], "unit": "s", "sweeps_sec": secs, "max_bond_after_sweep": bonds, "ramp_sec": ramp,
            "E": [float(np.real(E)), float(np.imag(E))], "EE": float(EE), "n_matvec": stats.get("n_matvec"),
            "bonds": 2 * (N - 1), "n_recycled": stats.get("n_recycled", 0), "n_exact_trunc": stats.get("n_exact_trunc", 0),
            "n_sharded_solves": stats.get("n_sharded_solves", 0), "n_sharded_env": stats.get("n_sharded_env", 0),
            "n_gauge_moves": stats.get("n_gauge_moves", 0),
            "how": "measured: TASEP N=%d chi=%d c128 two-site sweeps after a %s ramp" % (N, chi, sched)}


if __name__ == "__main__":
    main()

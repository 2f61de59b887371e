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

N > 1 (torchrun, one rank per GPU): the headline is the SAME mat-vec strong-scaled over the ranks (ket split, exchange
fused into the kernels over NVLink; ``scaling: strong``); all four sharded variants are timed and checked against the
unsharded result under ``sharded``; ``scan`` = 8 points per GPU of the configs[3] s-scan (weak scaling, dynamic queue,
3 concurrent points per GPU); ``weak_replicas`` = N independent mat-vecs for the record.

``--impl reference`` times the oracle port of the reference's CPU implementation of the same chi=1024 closure
(single-threaded un-optimised einsum, exactly the calls the reference makes) on bounded slab samples of it; rank 0 only.
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

# TASEP s-ensemble point of the benchmark: s = +0.5 is the current-ENHANCED side (hops carry e^{+s}), an entangled state
# (EE = 0.83 at N=100); round 1 used s = -0.5, whose dominant state is a product state (EE = 1e-15: nothing to solve)
TASEP = (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, 0.5)
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
def seeded_operands(chi, seed=0):
    """The bench operands as HOST arrays, seeded: identical in both arms and in the parity check."""
    rng = np.random.default_rng(seed)
    cr = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    return cr(chi, 6, chi), cr(chi, 6, chi), cr(2, 2, chi, chi)


def run_reference(args, rank, world):
    """The reference's CPU implementation of the step (oracle port: the un-optimised np.einsum calls of
    diag_tools.py:153-156, single-threaded c_einsum) on the SAME configuration as the GPU arm, chi=1024.  A full call
    takes ~60 s, so a step is a bounded sample of it: every stage of the reference's contraction carries the bra index
    r of the right block, so restricting R to a slab of chi/16 rows computes exactly y[..., slab] with exactly 1/16 of
    the FLOPs of every stage -- step k takes slab k mod 16.  ONE full call is timed as well (size / slab insensitivity)."""
    if rank != 0:
        return
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import mpo as M
    W1, W2 = M.asep_mpo(6, TASEP)[0][2:4]
    chi, nslab = args.chi, 16
    L, R, x = seeded_operands(chi)
    rows = chi // nslab
    slab = lambda k: _slab_call(orc, x, L, W1, W2, R, k % nslab, rows)
    for k in range(args.warmup):
        slab(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        slab(k)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    fl = heff2_flops(chi)
    val = (fl / nslab) / dt * 1e-9
    full = None
    if not args.no_extras:
        t0 = time.perf_counter()
        yfull = orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "einsum")
        full = time.perf_counter() - t0
        ys = _slab_call(orc, x, L, W1, W2, R, 3, rows)
        slab_err = float(np.linalg.norm(ys - yfull.reshape(x.shape)[..., 3 * rows:4 * rows].reshape(-1)) / np.linalg.norm(ys))
    # BLAS-fair variant (same contractions through threaded zgemm), reported alongside
    tb = time.perf_counter(); orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas"); tb = time.perf_counter() - tb
    line = {"impl": "reference", "metric": "heff_matvec_fp64_gflops", "value": val, "unit": "GFLOP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": workload_config(chi),
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": 1, "kind": "port",
                             "sample": "per step: the chi=%d two-site Hfun of the GPU arm restricted to a slab of %d of the %d right-bond "
                                       "rows (1/%d of the FLOPs of every einsum stage of diag_tools.py:153-156; un-optimised np.einsum = "
                                       "single-threaded c_einsum); %d steps" % (chi, rows, chi, nslab, args.steps),
                             "full_call_sec": full, "full_call_gflops": (fl / full * 1e-9) if full else None,
                             "slab_vs_full_rel_err": slab_err if full else None,
                             "blas_fair_gflops": fl / tb * 1e-9, "blas_fair_cores": os.cpu_count()},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _slab_call(orc, x, L, W1, W2, R, k, rows):
    """y[..., k*rows:(k+1)*rows] of the two-site Hfun through the reference's four einsum stages (r restricted)."""
    Rk = R[k * rows:(k + 1) * rows]
    xr = x
    s1 = np.einsum("ros,nqks->ronqk", Rk, xr)                    # diag_tools.py:153
    s2 = np.einsum("lopq,ronqk->lprnk", W2, s1)                  # :154
    s3 = np.einsum("jlmn,lprnk->jmprk", W1, s2)                  # :155
    return -np.einsum("ijk,jmprk->mpir", L, s3).reshape(-1)      # :156, :162


def workload_config(chi):
    return {"workload": "TASEP (alpha=0.35, beta=2/3, s=+0.5) N=100 chi=%d two-site H_eff mat-vec (d=2,w=6,c128); BASELINE configs[2]" % chi,
            "chi": chi, "N": 100, "d": 2, "w": 6, "flops_per_step": heff2_flops(chi),
            "l2": "operands 96+96+64 MiB and the 384 MiB intermediate exceed the 126 MB L2; no flush needed"}


def measure_fp64_peak(reps=3, n=8192):
    """FP64 tensor-pipe denominator measured NOW on this box: cuBLAS Zgemm n^3 through torch.matmul (a library call used
    only as the yardstick; MEASURED_PEAKS.json has no FP64 entry).  Best of ``reps``."""
    import torch
    try:
        a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
        b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
        torch.matmul(a, b)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            torch.matmul(a, b)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 8.0 * n ** 3 / (best * 1e-3) * 1e-12
    except Exception:
        return None


def timed_loop(fn, steps, barrier, dist, world):
    """K calls bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks -> ms per step."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


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
    ap.add_argument("--no-full-sweep", action="store_true", help="skip the measured two-site sweeps of the N=100 chain (~2 min)")
    ap.add_argument("--svd", default="auto", choices=["auto", "gesvd", "gesvdj", "gram"])
    ap.add_argument("--gauge-shortcut", action="store_true", help="N = 1: also time the sweeps with run_dmrg's gauge_shortcut option")
    ap.add_argument("--sharded-sweep", action="store_true",
                    help="N > 1: also time two-site sweeps of the N=100 chain with the mat-vecs and block updates of the large "
                         "bonds sharded over the ranks (run_dmrg(shard=...))")
    ap.add_argument("--budget-sec", type=int, default=300, help="N = 1: wall-clock budget of the extras; what does not fit is skipped "
                                                                "and says so in the line")
    ap.add_argument("--scan-points", type=int, default=-1, help="points of the configs[3] scan to run (dynamic scheduling over the "
                                                                   "ranks, 3 workers per GPU); -1 = 8 per GPU, 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # rank 0 must print ONE JSON line on stdout, but NCCL writes its version banner there from C (whatever NCCL_DEBUG says
    # on this image): everything that goes to fd 1 during the run is sent to stderr, the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pydmrg_b200 import core, device as D, eigs, mpo as M, parallel
    from pydmrg_b200._lib import launch_count

    chi, d, w = args.chi, 2, 6
    ctx = core.Context("c128", svd_method=args.svd)
    W1, W2 = M.asep_mpo(6, TASEP)[0][2:4]
    Wc = D.combine_twosite(W1, W2, np.complex128)
    # identical seeded operands on every rank (the N > 1 headline is ONE mat-vec strong-scaled over the ranks)
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
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

    # ---- the step: N = 1 the plain mat-vec; N > 1 the SAME mat-vec split over the ranks (strong scaling) ------------
    sharded_variants = {}
    parity = {}
    if world == 1:
        step, how = (lambda: plan.apply(x, y)), "single GPU"
    else:
        plan.apply(x, y)
        y_ref = y.clone()

        def check(tag, obj, yy):
            obj.apply(x, yy)
            torch.cuda.synchronize()
            err = float((torch.linalg.vector_norm(yy - y_ref) / torch.linalg.vector_norm(y_ref)).item())
            y0 = yy.clone()
            dist.broadcast(torch.view_as_real(y0), src=0)
            same = torch.tensor([1 if torch.equal(y0, yy) else 0], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            parity[tag] = {"rel_err_vs_unsharded": err, "bit_identical_on_all_ranks": bool(same.item())}

        ops1 = [(L, Wc, R)]
        ys = torch.empty_like(x)
        objs = {"bond_nccl": parallel.ShardedHeff(D.C128, d * d, chi, chi, ops1, ctx.ws, rank, world, dist),
                "bond_fused": parallel.FusedShardedHeff(D.C128, d * d, chi, chi, ops1, ctx.ws, rank, world, dist),
                "ket_nccl": parallel.ShardedHeff(D.C128, d * d, chi, chi, ops1, ctx.ws, rank, world, dist, mode="ket"),
                "ket_fused": parallel.FusedShardedHeff(D.C128, d * d, chi, chi, ops1, ctx.ws, rank, world, dist, mode="ket")}
        for tag, obj in objs.items():
            check(tag, obj, ys)
        head = objs["ket_fused"]
        step, how = (lambda: head.apply(x, ys)), ("ONE mat-vec split %d ways over the left ket index; stage C = one long-K GEMM "
                                                  "whose scatter epilogue reduce-scatters the partial y over NVLink, then a "
                                                  "flag-guarded reduce + all-gather kernel (no NCCL on the data path)" % world)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = launch_count()
    ms_step = timed_loop(step, args.steps, barrier, dist, world)
    launches = launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    value = flops / (ms_step * 1e-3) * 1e-9

    # ---- per-stage timing of the single-GPU step (roofline of the dominant kernel) --------------------------------
    st = np.zeros(3)
    nrep = max(3, min(args.steps, 10))
    for _ in range(nrep):
        st += np.array(plan.apply_timed(x, y))
    st /= nrep
    ms_single = timed_loop(lambda: plan.apply(x, y), max(3, min(args.steps, 10)), barrier, dist, world) if world > 1 else ms_step
    # FLOPs the two GEMM launches actually execute: chi^3 terms of the non-empty MPO slabs (for TASEP
    # rows 3,4 of W are zero, so stage C runs 4 of 6 slabs); complex MAC = 8 FLOP
    cube = plan.flops_gemm
    gemm_ms = st[0] + st[2]
    # (skipped with --no-extras: that is the command the ncu launch list is taken from, and the cuBLAS kernels are not part of a step)
    peak_now = measure_fp64_peak() if ((rank == 0 or world == 1) and not args.no_extras) else None
    peak_file, peak_src = fp64_peak()
    peak = peak_now if peak_now else peak_file
    achieved = cube / (gemm_ms * 1e-3) * 1e-12
    roofline = {"bound": "tensor", "kernel": "gemm_nt_dmma_kernel<c128> (FP64 DMMA.8x8x4, TMA-fed)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None,
                "peak_source": ("cuBLAS Zgemm 8192^3 measured in this run (fp64_peak_measured_now)" if peak_now else peak_src)
                               + "; FP64 tensor pipe, MEASURED_PEAKS.json has no FP64 entry",
                "fp64_peak_measured_now": peak_now, "fp64_peak_committed": peak_file,
                "launches_per_step": 2, "flops_per_launch_avg": cube / 2, "ms_per_launch_avg": gemm_ms / 2,
                "note": "achieved counts EXECUTED GEMM FLOPs (MPO block sparsity skipped); the headline value "
                        "counts the dense-W algorithmic FLOPs of SURVEY 8d; measured on the single-GPU step",
                "dense_algorithmic_flops_per_step": flops, "executed_gemm_flops_per_step": cube,
                "stage_ms": {"gemm_stage_A": st[0], "mix": st[1], "gemm_stage_C": st[2]},
                "share_of_step": gemm_ms / ms_single}
    tr = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch_avg")
        except Exception:
            pass

    # ---- e2e: the product's reference-shaped closure, HOST vector in, HOST vector out ------------------------------
    nbytes = x.numel() * 16
    xh = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    xh.copy_(x)
    xh_np = xh.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    if world == 1:
        # eigs.make_ham_func(M, W, F, site, oneSite=False) -- the two-slot [L, R] addressing of the reference's closure
        dummyA, dummyB = torch.empty((d, chi, 1), dtype=x.dtype, device="cuda"), torch.empty((d, 1, chi), dtype=x.dtype, device="cuda")
        Hfun = eigs.make_ham_func([dummyA, dummyB], [[W1, W2]], [[L, R]], 0, oneSite=False, ctx=ctx)[0]
        holder = {}

        def e2e_step():
            holder["y"] = Hfun(xh_np)                            # H2D of x, mat-vec, D2H of y, synchronised

        api = "pydmrg_b200.eigs.make_ham_func(M, W, F, site, oneSite=False)[0](x_host: ndarray) -> y_host: ndarray"
    else:
        yh = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        xd2 = torch.empty_like(x)

        def e2e_step():
            xd2.copy_(xh, non_blocking=True)                     # every rank needs x; every rank ends with y
            head.apply(xd2, ys)
            yh.copy_(ys, non_blocking=True)
            torch.cuda.synchronize()

        api = "FusedShardedHeff(mode='ket').apply with pinned host x / y on every rank"
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
    e2e = {"value": flops / (e2e_ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world, "ms_per_step": e2e_ms, "api": api}

    cfg = workload_config(chi)
    cfg["parallelism"] = how
    line = {"metric": "heff_matvec_fp64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    if world > 1:
        line["ms_per_step_unsharded_one_gpu"] = ms_single
        line["speedup_vs_one_gpu"] = ms_single / ms_step
        for tag, obj in objs.items():
            for _ in range(3):
                obj.apply(x, ys)
            ms = timed_loop(lambda: obj.apply(x, ys), args.steps, barrier, dist, world)
            sharded_variants[tag] = {"ms_per_matvec": ms, "gflops": flops / (ms * 1e-3) * 1e-9, "speedup": ms_single / ms, **parity[tag]}
        per, ideal = parallel.partition_cost(objs["bond_nccl"].partition[0])
        line["sharded"] = {"what": "one chi=%d two-site mat-vec split over the ranks; every variant checked against the unsharded "
                                   "result on the same operands (rel_err) and for bit-identical y on all ranks" % chi,
                           "variants": sharded_variants, "bond_split_slab_gemms_per_rank": per, "bond_split_ideal_speedup": ideal}
        # weak scaling for the record (N independent mat-vecs, no collective: cannot fail)
        ms_w = timed_loop(lambda: plan.apply(x, y), args.steps, barrier, dist, world)
        line["weak_replicas"] = {"ms_per_step": ms_w, "aggregate_gflops": world * flops / (ms_w * 1e-3) * 1e-9}
        for obj in (objs["bond_fused"], objs["ket_fused"]):
            obj.close()
        npts = 8 * world if args.scan_points < 0 else args.scan_points
        if npts:
            line["scan"] = scan_multi(npts, rank, world, dist)

    if world > 1 and args.sharded_sweep:
        spec = parallel.ShardSpec(rank, world, dist, mode="ket")
        line["sweep_sec_sharded"] = full_sweep(args, ctx, shard=spec)
        line["sweep_sec_sharded"]["how"] += "; every rank runs the sweep, mat-vecs and block updates of bonds >= %d split %d ways" % (
            spec.min_dim, world)

    # ---- extras on rank 0 at N=1, most important first, inside a wall-clock budget ------------------------
    if world == 1 and not args.no_extras:
        t_start = time.perf_counter()

        def extra(key, fn, need=0.0):
            """Run one extra unless the budget is (nearly) spent; log its duration to stderr."""
            used = time.perf_counter() - t_start
            if used + need > args.budget_sec:
                line[key] = {"skipped": "time budget (%.0f s of %d s used)" % (used, args.budget_sec)}
                return None
            t0 = time.perf_counter()
            line[key] = fn()
            print("[bench] %-20s %6.1f s" % (key, time.perf_counter() - t0), file=sys.stderr, flush=True)
            return line[key]

        cb = extra("cpu_baseline", lambda: cpu_baseline(args, ctx))
        if cb:
            line["parity_rel_err"] = cb["parity_rel_err"]
        bs = extra("bond_step", lambda: bond_step(args, ctx, plan, L, R, W1, W2, chi))
        if bs:
            line["sweep_sec_derived"] = {"value": bs["ms_total_recycled"] * 2 * 99 * 1e-3,
                                         "how": "198 bond steps x measured chi=%d bond step with %d Davidson cycles each and the "
                                                "warm-started truncation (edge bonds are smaller)" % (chi, args.davidson_cycles)}
        if not args.no_full_sweep:
            extra("sweep_sec", lambda: full_sweep(args, ctx), need=90.0)
            if args.gauge_shortcut:
                extra("sweep_sec_shortcut", lambda: full_sweep(args, ctx, gauge_shortcut=True), need=90.0)
        extra("small_chi_matvec", small_chi_matvec)
        extra("krylov", lambda: krylov_bandwidth(chi))
        extra("config4_matvec", config4_matvec)
        npts = 8 if args.scan_points < 0 else args.scan_points
        if npts:
            extra("scan", lambda: scan_multi(npts, 0, 1, None), need=45.0)
        line["extras_sec"] = time.perf_counter() - t_start
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    os.close(real_stdout)


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


def cpu_baseline(args, ctx):
    """Oracle port on this box's host cores, on a bounded sample of the SAME chi=1024 workload: (A) the reference's own
    un-optimised einsum calls (single-threaded c_einsum) on r-slabs of the mat-vec (see run_reference), (B) the BLAS-fair
    restatement on all cores, one full call -- whose output is also the parity check of the GPU result on the same
    seeded operands (``parity_rel_err``, the oracle as the checker)."""
    import torch
    from oracle import dmrg_oracle as orc
    from pydmrg_b200 import device as D, mpo as M
    W1, W2 = M.asep_mpo(6, TASEP)[0][2:4]
    chi, nslab = args.chi, 16
    rows = chi // nslab
    L, R, x = seeded_operands(chi)
    t = time.perf_counter(); n = 0
    while n < 2 or (time.perf_counter() - t < 12 and n < nslab):
        _slab_call(orc, x, L, W1, W2, R, n % nslab, rows); n += 1
    ta = (time.perf_counter() - t) / n
    t = time.perf_counter()
    ref = orc.heff_twosite(x.reshape(-1), [L], [W1], [W2], [R], x.shape, "blas")
    tb = time.perf_counter() - t
    plan = D.HeffPlan(D.C128, 4, chi, chi, [(ctx.dev(L), D.combine_twosite(W1, W2, np.complex128), ctx.dev(R))], ctx.ws)
    xd = ctx.dev(x.reshape(-1)); yd = torch.empty_like(xd)
    plan.apply(xd, yd, -1.0)
    err = float(np.linalg.norm(yd.cpu().numpy() - ref) / np.linalg.norm(ref))
    plan.close()
    return {"value": heff2_flops(chi) / nslab / ta * 1e-9, "unit": "GFLOP/s", "cores": 1, "kind": "port",
            "sample": "reference-faithful two-site Hfun (np.einsum, no optimize: single-threaded c_einsum) at chi=%d on %d slabs of "
                      "%d right-bond rows (1/%d of every stage each)" % (chi, n, rows, nslab),
            "parity_rel_err": err, "parity_what": "GPU mat-vec vs oracle (BLAS mode) on the same seeded chi=%d operands" % chi,
            "blas_fair": {"value": heff2_flops(chi) / tb * 1e-9, "unit": "GFLOP/s", "cores": os.cpu_count(),
                          "sample": "same contractions via np.tensordot (threaded zgemm) at chi=%d, 1 full call" % chi}}


def small_chi_matvec(reps=30):
    """The mat-vec where configs[1] and [3] live (chi = 256, 512; Heisenberg f64 w=5 and ASEP c128 w=6): ms, executed GEMM
    TFLOP/s of the two GEMM launches (CUDA events, pdmrg_heff_apply_timed) and of the whole mat-vec."""
    import torch
    from pydmrg_b200 import core, device as D, mpo as M
    out = {}
    for tag, code, dt, mk in (("heisenberg_f64", D.F64, torch.float64, lambda: M.heisenberg_mpo(6)),
                              ("asep_c128", D.C128, torch.complex128, lambda: M.asep_mpo(6, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, -0.5)))):
        ctx = core.Context("f64" if code == D.F64 else "c128")
        W1, W2 = mk()[0][2:4]
        w = W1.shape[0]
        for chi in (128, 256, 512):
            Wc = D.combine_twosite(W1, W2, D.np_dtype(code))
            g = torch.Generator(device="cuda"); g.manual_seed(5)
            rnd = lambda *s: torch.randn(*s, dtype=dt, device="cuda", generator=g)
            L, R, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(4 * chi * chi)
            y = torch.empty_like(x)
            plan = D.HeffPlan(code, 4, chi, chi, [(L, Wc, R)], ctx.ws)
            for _ in range(5):
                plan.apply(x, y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps):
                plan.apply(x, y)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            st = np.zeros(3)
            for _ in range(5):
                st += np.array(plan.apply_timed(x, y))
            st /= 5
            out["%s_chi%d" % (tag, chi)] = {"ms": ms, "stage_ms": [float(v) for v in st],
                                            "executed_gemm_tflops_whole_matvec": plan.flops_gemm / (ms * 1e-3) * 1e-12,
                                            "executed_gemm_tflops_gemm_launches": plan.flops_gemm / ((st[0] + st[2]) * 1e-3) * 1e-12}
            plan.close()
    return out


SCAN_PRM = (0.5, 0.5, 0.1, 0.9, 0.5, 0.5)


def scgf_shape_ok(s_vals, E):
    from pydmrg_b200.scan import scgf_shape_ok as ok
    return ok(s_vals, E)


def scan_multi(points, rank, world, dist, max_iter=10):
    """configs[3] over the ranks: ``points`` of the 64 s-values (evenly spread over the 62 that are not adjacent to s=0;
    8 per GPU by default, i.e. weak scaling), cold starts, dynamic queue (a worker takes the next point when it is free,
    points nearest s=0 first), 3 concurrent points per GPU; no data-path collective."""
    import torch
    from pydmrg_b200 import mpo as M, scan
    # the two points next to s = 0 (indices 31, 32; s = -/+0.008) are left out: their leading pair is nearly defective, single-
    # state solves run into the reference's 1000-cycle cap at every bond (12 s and 95 s per point, profiles/scan_diag_r02.txt)
    # and the point is ill-posed without nStates=2; ONE such point would be the whole wall time of an 8-GPU scan
    pool = [i for i in range(64) if i not in (31, 32)]
    points = min(points, len(pool))
    idx = [pool[int(round(v))] for v in np.linspace(0, len(pool) - 1, points)]
    s_vals = np.linspace(-0.5, 0.5, 64)[idx]
    mpo_of = lambda s: M.asep_mpo(50, SCAN_PRM + (float(s),))
    scan.run_scan(lambda s: M.asep_mpo(8, SCAN_PRM + (float(s),)), s_vals[:1], 16, ramp=(4,), maxIter=0)      # start-up
    order = [int(i) for i in np.argsort(np.abs(s_vals), kind="stable")]          # the gap closes at s = 0: those points first
    st = {}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = scan.run_scan(mpo_of, s_vals, 256, rank, world, dist if world > 1 else None, maxIter=max_iter, tol=1e-8, schedule="dynamic",
                        order=order, workers=3, seed=[pool.index(i) for i in idx], stats=st)     # one seed per GRID point, whatever the subset
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"points": points, "point_indices_of_64": idx, "wall_sec": dt, "s_points_per_hour": points / dt * 3600.0, "scaling": "weak",
            "points_done_by_this_rank": len(res["points"]), "n_matvec_this_rank": st.get("n_matvec"),
            "E_real": [float(np.real(e)) for e in res["E"]], "E_imag_max": float(np.max(np.abs(np.imag(res["E"])))),
            "E_monotone_and_convex_in_s": scgf_shape_ok(s_vals, res["E"]),
            "how": "configs[3]: %d of the 64 s-values (ASEP N=50 chi=256 two-site), cold starts through a [16,64] ramp, tol 1e-8 / at "
                   "most %d sweeps at chi=256; the reference's solver settings; the two grid points adjacent to s=0 (nearly defective "
                   "leading pair: 12 s and 95 s per point at the reference's 1000-cycle cap, ill-posed without nStates=2) are excluded; "
                   "dynamic queue (points nearest s=0 first), 3 concurrent points per GPU" % (points, max_iter + 1)}


def full_sweep(args, ctx, shard=None, gauge_shortcut=False):
    """Measured two-site sweeps (right + left over all 99 bonds) of the TASEP N=100 chain at chi: the state is
    grown through a [16, 64, 256] chi ramp (so it is physical) and then swept four times with mbd = chi.  Two-site
    sweeps are not zero-padded: every bond grows by up to a factor d per visit, so the first two sweeps run on
    the growing tensors (256 -> 512 -> 1024), the third is the first at full size (every bond decomposed from
    scratch) and the later ones are converged sweeps with the warm-started truncation (centre bond exact).
    ``value`` is the last sweep."""
    import torch
    from pydmrg_b200 import dmrg, env, mpo as M
    N, chi = 100, args.chi
    mpo = M.asep_mpo(N, TASEP)
    np.random.seed(0)
    sched = [c for c in (16, 64, 256) if c < chi]
    t0 = time.perf_counter()
    extra = {} if shard is None else {"shard": shard}
    if gauge_shortcut:
        extra["gauge_shortcut"] = True
    # chi ramp with in-memory continuation: the cold chi=16 and chi=64 stages run ONE sweep with a loose Davidson residual
    # (their job is to get near the state cheaply: 176k latency-bound mat-vecs at the reference's 1e-7 otherwise), the
    # chi=256 stage two sweeps at the normal stopping rule
    ramp_stats, guess, E_st, EE_st = {}, None, [], []
    for c, rt, mi in [(c, rt, mi) for c, rt, mi in ((16, 1e-5, 0), (64, 1e-6, 0), (256, None, 1)) if c < chi]:
        o = dmrg.run_dmrg(mpo, mbd=c, tol=1e-9, maxIter=mi, twoSite=True, ctx=ctx, stats=ramp_stats, returnState=True, initGuess=guess,
                          res_tol=rt, **extra)
        guess = o[-1]
        E_st.append(o[0]); EE_st.append(o[1])
    ramp_out = [E_st, EE_st]
    mpsL = [[ctx.dev(t) for t in guess[0]]]
    F = env.calc_env(mpsL[0], mpo, chi, gaugeSite=0, ctx=ctx)
    torch.cuda.synchronize()
    ramp = time.perf_counter() - t0
    cache, secs, bonds, mvs, prof, E, EE, stats = {}, [], [], [], [], None, None, {}
    for _ in range(4):
        stats = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, chi, ctx=ctx, stats=stats, recycle=cache, **extra)
        torch.cuda.synchronize(); secs.append(time.perf_counter() - t0)
        bonds.append(int(max(t.shape[2] for t in mpsL[0])))
        mvs.append(int(stats.get("n_matvec", 0)))
    # bond steps with real solver work at full chi: one continuation step of an s-scan (s: 0.5 -> 0.45) from the converged
    # state, as the reference's scan scripts do (scripts/asep/current/bondDimConv.py:19-27) -- the centre is moved to site 46,
    # the blocks are rebuilt for the new operator and the 8 central bonds are solved to the reference's stopping rule
    # (capped at 200 cycles): seconds and mat-vecs per bond.  (A whole sweep of such steps is minutes; this is its unit.)
    step = {}
    if shard is None and not gauge_shortcut:
        from pydmrg_b200 import gauge
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mpo2 = M.asep_mpo(N, TASEP[:6] + (TASEP[6] - 0.05,))
        lo = N // 2 - 4
        mps2 = [gauge.move_gauge([t for t in mpsL[0]], 0, lo, ctx=ctx)]
        F2 = env.calc_env(mps2[0], mpo2, chi, gaugeSite=lo, ctx=ctx)
        torch.cuda.synchronize(); t_prep = time.perf_counter() - t0
        cache2, per_bond = {}, []
        for site in range(lo, lo + 8):
            st2 = {}
            torch.cuda.synchronize(); t0 = time.perf_counter()
            E2, mps2, F2, EE2, _, _ = dmrg.twoSiteStep(mps2, mpo2, F2, site, "right", chi, ctx=ctx, stats=st2, recycle=cache2, max_cycle=200)
            torch.cuda.synchronize()
            per_bond.append({"site": site, "sec": time.perf_counter() - t0, "n_matvec": int(st2.get("n_matvec", 0)),
                             "last_residual": st2.get("last_residual")})
        step = {"what": "scan continuation at chi=%d: parameter step s = 0.5 -> 0.45 from the converged state, the 8 central bonds solved to "
                        "the reference's stopping rule (at most 200 cycles), each with its first-visit truncation and block update" % chi,
                "prep_sec_gauge_move_and_blocks": t_prep, "bonds": per_bond,
                "mean_sec_per_bond": float(np.mean([b["sec"] for b in per_bond])),
                "mean_matvecs_per_bond": float(np.mean([b["n_matvec"] for b in per_bond])),
                "E_local": [float(np.real(E2)), float(np.imag(E2))]}
        del F2, cache2, mps2
    full = [i for i, b in enumerate(bonds) if b >= chi]
    first_full = full[1] if len(full) > 1 else (full[0] if full else len(secs) - 1)   # first sweep that STARTS at full size
    nb = 2 * (N - 1)
    return {"value": secs[-1], "unit": "s", "sweeps_sec": secs, "max_bond_after_sweep": bonds, "ramp_sec": ramp,
            "first_full_size_sweep_sec": secs[first_full], "converged_sweep_sec": secs[-1],
            "time_to_solution_sec": ramp + sum(secs[:first_full + 2]),
            "after_parameter_step": step,
            "n_matvec_per_sweep": mvs, "matvecs_per_bond_per_sweep": [round(m / nb, 2) for m in mvs],
            "E": [float(np.real(E)), float(np.imag(E))], "EE": float(EE),
            "E_ramp_stages": [float(np.real(e)) for e in np.atleast_1d(ramp_out[0])], "EE_ramp_stages": [float(np.real(e)) for e in np.atleast_1d(ramp_out[1])],
            "ramp_n_matvec": ramp_stats.get("n_matvec"),
            "n_matvec": stats.get("n_matvec"), "bonds": nb, "n_recycled": stats.get("n_recycled", 0),
            "n_exact_trunc": stats.get("n_exact_trunc", 0), "n_recycle_rejected": stats.get("n_recycle_rejected", 0),
            "disc_max": stats.get("disc_max"),
            "n_sharded_solves": stats.get("n_sharded_solves", 0), "n_sharded_env": stats.get("n_sharded_env", 0),
            "n_gauge_moves": stats.get("n_gauge_moves", 0),
            "how": "measured: TASEP s=+0.5 (entangled, EE ~ 0.83) N=%d chi=%d c128 two-site sweeps after a %s ramp; value = last "
                   "(converged) sweep: the chi=256 stage already holds the state to 5e-12, so a sweep at chi=%d costs ONE mat-vec per "
                   "bond -- that is what a converged sweep is; time_to_solution = ramp + growth sweeps + first full-size sweep + one "
                   "converged sweep; after_parameter_step = sweeps that have real solver work to do at full chi"
                   % (N, chi, sched, chi)}


if __name__ == "__main__":
    main()

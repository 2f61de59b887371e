#!/bin/bash
# First GPU session of the next round: everything that was written after the round-1 GPU budget was spent,
# in the order of what it unblocks.  Each item is one gpurun call (run from the repo root, here):
#
#   1 GPU :  gpurun --timeout 1500 -- 'bash tools/round2_validate.sh one'
#   2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_validate.sh two'
#   8 GPUs:  gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_validate.sh eight'
#
# Outputs land in gpurun_out/ (merged back); copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
case "${1:-one}" in
one)
  # (a) the validated suite + the pending tests (two-site nStates=2, state-averaged truncation)
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
  PDMRG_RUN_PENDING=1 timeout 300 python -m pytest tests/test_pending_gpu.py -m gpu -q > gpurun_out/r2_pending.log 2>&1; echo "pending rc=$?" | tee -a gpurun_out/r2_pending.log
  # (b) bench line with the start policy of the warm-started truncation in place (profiles/README note)
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_err.log; echo "bench rc=$?"
  # (b2) cold scan point with fewer Davidson cycles: diagonal preconditioner and a larger subspace
  timeout 300 python - > gpurun_out/r2_scan_precond.txt 2>&1 <<'PY'
import time, numpy as np, torch
from pydmrg_b200 import mpo as M, scan
mpo_of = lambda s: M.asep_mpo(50, (0.5, 0.5, 0.1, 0.9, 0.5, 0.5, float(s)))
scan.run_scan(mpo_of, [-0.5], 64, maxIter=0)                                     # start-up
for opts in ({}, {"precond": "diag"}, {"max_space": 24}, {"precond": "diag", "max_space": 24}):
    st = {}
    np.random.seed(0); torch.cuda.synchronize(); t0 = time.perf_counter()
    r = scan.run_scan(mpo_of, [-0.5, 0.17], 256, maxIter=3, stats=st, **opts)
    torch.cuda.synchronize()
    print(opts, "sec/point", r["sec"], "E", [complex(e) for e in r["E"]], "matvecs", st.get("n_matvec"), "cycles", st.get("cycles"), flush=True)
PY
  # (b3) converged chi=1024 sweeps with and without the gauge shortcut
  timeout 400 python - > gpurun_out/r2_gauge_shortcut.txt 2>&1 <<'PY'
import time, numpy as np, torch
from pydmrg_b200 import core, dmrg, env, mpo as M
ctx = core.Context("c128")
mpo = M.asep_mpo(100, (0.35, 0.0, 1.0, 0.0, 2.0 / 3.0, 0.0, -0.5))
np.random.seed(0)
out = dmrg.run_dmrg(mpo, mbd=[16, 64, 256], tol=1e-9, maxIter=1, twoSite=True, ctx=ctx, returnState=True)
for gs in (False, True):
    mpsL = [[ctx.dev(t) for t in out[-1][0]]]
    F = env.calc_env(mpsL[0], mpo, 1024, gaugeSite=0, ctx=ctx)
    cache = {}
    for sw in range(6):
        st = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        E, mpsL, F, EE, EEs, S = dmrg.twoSiteSweep(mpsL, mpo, F, 1024, ctx=ctx, stats=st, recycle=cache, gauge_shortcut=gs)
        torch.cuda.synchronize()
        print("gauge_shortcut", gs, "sweep", sw, "%.3f s" % (time.perf_counter() - t0), "E", complex(E), "EE", float(EE),
              {k: st.get(k) for k in ("n_matvec", "n_gauge_moves", "n_recycled", "n_exact_trunc")}, flush=True)
    del mpsL, F, cache
    torch.cuda.empty_cache()
PY
  # (c) where does a latency-bound Davidson cycle go (host vs device), one process
  timeout 200 python tools/scan_contention.py > gpurun_out/r2_contention_n1.txt 2>&1
  ;;
two)
  # sharded sweeps (run_dmrg(shard=...)) against unsharded ones, then the chi=1024 sweep timing
  timeout 400 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tests/mgpu_sharded_sweep_check.py > gpurun_out/r2_sharded_sweep_n2.txt 2>&1; echo "sharded sweep check rc=$?"
  PDMRG_CHECK_TIMING=1 timeout 400 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
      tests/mgpu_sharded_sweep_check.py > gpurun_out/r2_sharded_sweep_timing_n2.txt 2>&1; echo "timing rc=$?"
  # the same through bench.py (adds sweep_sec_sharded to the JSON line)
  timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
      bench.py --gpus 2 --steps 5 --warmup 3 --sharded-sweep > gpurun_out/r2_bench_n2_sharded_sweep.json 2> gpurun_out/r2_bench_n2_err.log; echo "bench sharded sweep rc=$?"
  ;;
eight)
  # host contention of the latency-bound scan: cycles/s per rank with 1, 2, 4, 8 ranks, unpinned and pinned
  for n in 1 2 4 8; do
    for opt in "" "--pin" "--blas1"; do
      timeout 200 python -m torch.distributed.run --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n \
          tools/scan_contention.py $opt > gpurun_out/r2_contention_n${n}${opt:+_${opt#--}}.txt 2>&1
    done
  done
  nproc > gpurun_out/r2_host.txt; lscpu >> gpurun_out/r2_host.txt 2>&1; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_host.txt 2>&1
  nvidia-smi topo -m >> gpurun_out/r2_host.txt 2>&1
  timeout 300 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 \
      -m pydmrg_b200.scan --N 50 --chi 256 --points 64 --max-iter 3 --pin-cores > gpurun_out/r2_scan_n8_pinned.json 2> gpurun_out/r2_scan_err.log
  ;;
esac

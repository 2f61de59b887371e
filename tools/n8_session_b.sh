#!/bin/bash
# 8-GPU session B (round 2): sharded sweeps (correctness at chi=128, timing at chi=1024, fused C++ solves), the strong-scaling
# bench line with parity flags, and the configs[3] scan over 8 GPUs x 3 workers with the dynamic queue.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
PDMRG_CHECK_TIMING=1 PDMRG_CHECK_SWEEPS=3 timeout 500 $TR --master-port 29611 tests/mgpu_sharded_sweep_check.py > gpurun_out/r2_sharded_sweep_n8.txt 2>&1; echo "sharded sweep check rc=$?"
grep "^{" gpurun_out/r2_sharded_sweep_n8.txt | cut -c1-3000
timeout 500 $TR --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 --scan-points 64 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8_err.log; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n8_err.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "speedup_vs_one_gpu", "ms_per_step_unsharded_one_gpu")})
print(json.dumps(d["sharded"]["variants"]))
print(json.dumps(d["e2e"])); print(json.dumps(d["scan"])[:600])
PY

#!/bin/bash
# Round-2 validation on one B200: the whole GPU suite, the bench line, the reference arm, configs[2] for real.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/r2_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -22 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1_err.log; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1_err.log
timeout 600 python tools/run_config2.py > gpurun_out/r2_config2.json 2> gpurun_out/r2_config2_err.log; echo "config2 rc=$?"; tail -3 gpurun_out/r2_config2_err.log; cat gpurun_out/r2_config2.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref_err.log; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref.json | cut -c1-1500

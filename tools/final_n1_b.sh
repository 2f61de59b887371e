#!/bin/bash
# Round-2 evidence on one B200: bench line (with the extras budget), launch list + full ncu capture of the GEMM, small-chi launch list.
set -u
mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1_err.log; echo "bench rc=$?"; grep "\[bench\]" gpurun_out/r2_bench_n1_err.log
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu1.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_dmma -s 6 -c 2 -o gpurun_out/r2_gemm_prof python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu2.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_small_matvec.py > gpurun_out/r2_plain2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_small_matvec_launches.csv python tools/ncu_small_matvec.py > gpurun_out/r2_ncu3.log 2>&1
echo "ncu small rc=$?"
timeout 300 python tools/run_configs.py cfg0 cfg1 > gpurun_out/r2_run_configs.json 2> gpurun_out/r2_run_configs_err.log; echo "configs rc=$?"; tail -c 1500 gpurun_out/r2_run_configs.json

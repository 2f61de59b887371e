#!/bin/bash
# 8-GPU session A (round 2): correctness of the sharded sweeps / mat-vecs at N=8 and the host-contention probe of the scan.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/r2_host_n8.txt; lscpu >> gpurun_out/r2_host_n8.txt 2>&1; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_host_n8.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2_host_n8.txt 2>&1
TR="python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tests/mgpu_sharded_sweep_check.py > gpurun_out/r2_sharded_sweep_n8.txt 2>&1; echo "sharded sweep check rc=$?"
tail -3 gpurun_out/r2_sharded_sweep_n8.txt
PDMRG_CHECK_CHI=1024 timeout 200 $TR --master-port 29612 tests/mgpu_sharded_check.py > gpurun_out/r2_sharded_n8.txt 2>&1; echo "sharded check rc=$?"
tail -3 gpurun_out/r2_sharded_n8.txt
for opt in "" "--pin" "--blas1"; do
  timeout 120 $TR --master-port 29613 tools/scan_contention.py $opt > gpurun_out/r2_contention_n8${opt:+_${opt#--}}.txt 2>&1
  grep "rank 0/" gpurun_out/r2_contention_n8${opt:+_${opt#--}}.txt
done

#!/bin/bash
# round-2 profile captures (one GPU): launch list of the bench command, full captures of the two variants of the
# speculative kernel, the memcheck subset
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs > gpurun_out/r02_bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 2 -c 1 -o gpurun_out/r02_fixed_full -f \
   python tools/prof_one.py 16.0 1 1 150 3 > gpurun_out/r02_ncu_fixed.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 3 -c 1 -o gpurun_out/r02_var_full -f \
   python tools/prof_var.py 8.0 1 1 3 > gpurun_out/r02_ncu_var.log 2>&1
tail -2 gpurun_out/r02_ncu_fixed.log; tail -2 gpurun_out/r02_ncu_var.log

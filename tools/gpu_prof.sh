#!/bin/bash
# timing of the scan kernel + one ncu full capture (no parity tests)
mkdir -p gpurun_out
timeout 200 python tools/prof_one.py 4.0 1 1 150 4 > gpurun_out/prof_one.log 2>&1; tail -2 gpurun_out/prof_one.log
timeout 200 python tools/prof_one.py 4.0 0 0 150 3 > gpurun_out/prof_one_count.log 2>&1; tail -1 gpurun_out/prof_one_count.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 1 -c 1 -o gpurun_out/scan_full -f \
   python tools/prof_one.py 4.0 1 1 150 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log

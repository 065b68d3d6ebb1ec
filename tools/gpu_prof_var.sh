#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 1 -c 1 -o gpurun_out/var_full -f \
   python tools/prof_var.py 2.0 1 1 3 > gpurun_out/ncu_var.log 2>&1
tail -3 gpurun_out/ncu_var.log

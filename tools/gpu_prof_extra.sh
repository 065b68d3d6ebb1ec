#!/bin/bash
# ncu --set full captures of the secondary kernels: the record filter (validate_dnan, 8 GiB) and the stream kernel on
# variable-length reads (config 4, 2 GiB) and on Illumina-like ids of varying length (2 GiB)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_filter_kernel -s 14 -c 1 -o gpurun_out/filter_full -f \
   python tools/prof_filter.py 8 3 > gpurun_out/ncu_filter.log 2>&1
tail -2 gpurun_out/ncu_filter.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 1 -c 1 -o gpurun_out/var_full -f \
   python tools/prof_var.py 2.0 1 1 3 > gpurun_out/ncu_var.log 2>&1
tail -2 gpurun_out/ncu_var.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fq_stream_kernel -s 1 -c 1 -o gpurun_out/real_full -f \
   python tools/prof_real.py 2.0 1 1 > gpurun_out/ncu_real.log 2>&1
tail -2 gpurun_out/ncu_real.log

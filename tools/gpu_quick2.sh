#!/bin/bash
# quick timing of the two headline shapes (kernel time): fixed 150 bp fused, variable 50-300 bp fused / index only
mkdir -p gpurun_out
{
echo "fused fixed 150 (8 GiB):"; timeout 300 python tools/prof_one.py 8.0 1 1 150 3 2>&1 | tail -1
echo "count fixed 150 (8 GiB):"; timeout 300 python tools/prof_one.py 8.0 0 0 150 3 2>&1 | tail -1
echo "fused var 50-300 (8 GiB):"; timeout 300 python tools/prof_var.py 8.0 1 1 3 2>&1 | tail -1
echo "index var 50-300 (8 GiB):"; timeout 300 python tools/prof_var.py 8.0 0 1 3 2>&1 | tail -1
echo "fused fixed 300 (8 GiB):"; timeout 300 python tools/prof_one.py 8.0 1 1 300 3 2>&1 | tail -1
} > gpurun_out/quick2.txt 2>&1
sed -e 's/Outcome(.*line_phase=0)//' gpurun_out/quick2.txt

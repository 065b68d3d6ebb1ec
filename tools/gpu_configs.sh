#!/bin/bash
# the BASELINE configs 2-4 through prof_one / prof_var (kernel time of the dominant kernel, 16 GiB shapes scaled to 8 GiB)
mkdir -p gpurun_out
{
echo "config 2 (delimit + record count, no index, no histograms), 8 GiB fixed 150 bp:"; timeout 300 python tools/prof_one.py 8.0 0 0 150 3 2>&1 | tail -1
echo "config 2' (delimit + line-end index), 8 GiB fixed 150 bp:"; timeout 300 python tools/prof_one.py 8.0 0 1 150 3 2>&1 | tail -1
echo "config 3 (per-position ACGTN + quality histograms, no index), 8 GiB fixed 150 bp:"; timeout 300 python tools/prof_one.py 8.0 1 0 150 3 2>&1 | tail -1
echo "configs 2+3 fused (delimit + index + histograms), 8 GiB fixed 150 bp:"; timeout 300 python tools/prof_one.py 8.0 1 1 150 3 2>&1 | tail -1
echo "config 4 (variable 50-300 bp, delimit + index + histograms), 8 GiB:"; timeout 300 python tools/prof_var.py 8.0 1 1 3 2>&1 | tail -1
echo "config 4' (variable 50-300 bp, delimit + index only), 8 GiB:"; timeout 300 python tools/prof_var.py 8.0 0 1 3 2>&1 | tail -1
echo "fixed 300 bp (delimit + index + histograms), 8 GiB:"; timeout 300 python tools/prof_one.py 8.0 1 1 300 3 2>&1 | tail -1
} > gpurun_out/configs.txt 2>&1
cat gpurun_out/configs.txt

#!/bin/bash
# iteration loop on the GPU box: parity tests (fail fast), then timing of the scan kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ] && [ "$1" != "force" ]; then exit $rc; fi
timeout 300 python tools/prof_one.py 4.0 1 1 150 4 > gpurun_out/prof_one.log 2>&1; tail -4 gpurun_out/prof_one.log
timeout 300 python tools/prof_one.py 4.0 0 1 150 3 > gpurun_out/prof_one_nohist.log 2>&1; tail -3 gpurun_out/prof_one_nohist.log
timeout 300 python tools/prof_one.py 4.0 0 0 150 3 > gpurun_out/prof_one_count.log 2>&1; tail -3 gpurun_out/prof_one_count.log

#!/bin/bash
# parity tests, a slice of the fuzz sweep, then kernel timings of the main shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 600 python tools/fuzz_sweep.py 0 60 2>&1 | tail -3
for i in 1 2; do timeout 200 python tools/prof_one.py 16.0 1 1 150 4 2>&1 | tail -2 | sed 's/Outcome.*scan ms/scan ms/'; done
timeout 200 python tools/prof_one.py 8.0 0 1 150 3 2>&1 | tail -1 | sed 's/Outcome.*scan ms/noHist scan ms/'
timeout 300 python tools/prof_real.py 4.0 1 1 2>&1 | tail -1
timeout 200 python tools/prof_one.py 4.0 1 1 300 3 2>&1 | tail -1 | sed 's/Outcome.*scan ms/300bp scan ms/'
timeout 200 python tools/prof_one.py 4.0 1 1 100 3 2>&1 | tail -1 | sed 's/Outcome.*scan ms/100bp scan ms/'

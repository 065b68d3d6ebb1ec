#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; tail -5 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 900 python tools/fuzz_sweep.py 0 100 2>&1 | tail -4
FQB_DEBUG=1 timeout 200 python tools/prof_one.py 8.0 1 1 150 4 2>&1 | tail -2
timeout 200 python tools/prof_one.py 8.0 0 1 150 3 2>&1 | tail -1
timeout 200 python tools/prof_one.py 8.0 0 0 150 3 2>&1 | tail -1
FQB_DEBUG=1 timeout 300 python tools/prof_real.py 4.0 1 1 2>&1 | tail -2
FQB_DEBUG=1 timeout 300 python tools/prof_real.py 4.0 0 1 2>&1 | tail -2
timeout 300 python tools/prof_real.py 4.0 0 0 2>&1 | tail -1
timeout 200 python tools/prof_var.py 4.0 1 1 3 2>&1 | tail -1
timeout 200 python tools/prof_var.py 4.0 0 1 3 2>&1 | tail -1

#!/bin/bash
# iteration loop on the GPU box: parity tests (fail fast), then timing of the scan kernel + timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ] && [ "$1" != "force" ]; then exit $rc; fi
FQB_DEBUG=1 timeout 200 python tools/prof_one.py 4.0 1 1 150 4 > gpurun_out/prof_one.log 2>&1; tail -3 gpurun_out/prof_one.log
timeout 200 python tools/prof_one.py 4.0 0 1 150 3 > gpurun_out/prof_one_nohist.log 2>&1; tail -1 gpurun_out/prof_one_nohist.log
timeout 200 python tools/prof_one.py 4.0 0 0 150 3 > gpurun_out/prof_one_count.log 2>&1; tail -1 gpurun_out/prof_one_count.log
FQB_TRACE=gpurun_out/trace.bin timeout 200 python tools/prof_one.py 4.0 1 1 150 2 > gpurun_out/trace.log 2>&1
python tools/trace_view.py gpurun_out/trace.bin 70 800 8 > gpurun_out/trace.txt 2>&1
rm -f gpurun_out/trace.bin
cat gpurun_out/trace.txt
timeout 200 python tools/prof_var.py 4.0 1 1 3 > gpurun_out/prof_var.log 2>&1; tail -1 gpurun_out/prof_var.log
timeout 200 python tools/prof_one.py 4.0 1 1 300 3 > gpurun_out/prof_300.log 2>&1; tail -1 gpurun_out/prof_300.log

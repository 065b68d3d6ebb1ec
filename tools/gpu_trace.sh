#!/bin/bash
mkdir -p gpurun_out
FQB_TRACE=gpurun_out/trace.bin timeout 200 python tools/prof_one.py 4.0 1 1 150 2 > gpurun_out/trace.log 2>&1; tail -1 gpurun_out/trace.log
python tools/trace_view.py gpurun_out/trace.bin 70 800 12 > gpurun_out/trace.txt 2>&1
python tools/trace_view.py gpurun_out/trace.bin 3 800 6 >> gpurun_out/trace.txt 2>&1
rm -f gpurun_out/trace.bin
cat gpurun_out/trace.txt

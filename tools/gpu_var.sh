#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  FQB_SCFG10=$v timeout 200 python tools/prof_var.py 4.0 1 1 3 2>&1 | tail -1 | sed "s/.*scan ms/variant $v var-length: scan ms/"
  FQB_SCFG10=$v timeout 200 python tools/prof_one.py 4.0 1 1 300 3 2>&1 | tail -1 | sed "s/.*scan ms/variant $v fixed-300: scan ms/"
done

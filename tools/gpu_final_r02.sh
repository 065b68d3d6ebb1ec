#!/bin/bash
# final round-2 captures on one GPU: the bench line, the reference arm, then the profiles
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.log 2>&1; tail -1 gpurun_out/r02_bench_1gpu.log > gpurun_out/r02_bench_1gpu.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.log 2>&1; tail -1 gpurun_out/r02_bench_ref.log > gpurun_out/r02_bench_reference_arm.json
bash tools/gpu_profiles_r02.sh
bash tools/gpu_configs.sh > gpurun_out/r02_configs.txt 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d.get("e2e", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"))
for c in d.get("configs", []):
    print(f'{c["value"]:8.1f} GB/s  scan {c["kernel_ms"]:.3f} idx {c["index_kernel_ms"]:.3f} frac {c["roofline_frac"]:.3f}  {c["workload"]}')
PY

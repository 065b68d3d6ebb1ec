"""One parse of N GiB synthetic input for ncu captures (python tools/prof_one.py [gib] [hist] [index])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_rs_b200 as fq
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
hist = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
want_index = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
read_len = int(sys.argv[4]) if len(sys.argv) > 4 else 150
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
rb = 17 + 2 * (read_len + 1) + 2
n = int(gib * (1 << 30)) // rb * rb
eng = fq.Engine(max_len=read_len)
t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
eng.synth_fixed(t, n, read_len=read_len)
idx = torch.empty(4 * (n // rb) + 8, dtype=torch.int32, device="cuda") if want_index else None
for _ in range(reps):
    eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
    out, st = eng.fetch()
    print(out, "scan ms", eng.last_scan_ms(), "GB/s", n / eng.last_scan_ms() / 1e6)

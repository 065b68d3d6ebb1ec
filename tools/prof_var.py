"""One parse of ~N GiB synthetic VARIABLE-length (50..300 bp) input (BASELINE config 4):
   python tools/prof_var.py [gib] [hist] [index] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_rs_b200 as fq
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
hist = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
want_index = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n_rec = int(gib * (1 << 30) / 371)
eng = fq.Engine(max_len=300)
t, n = eng.synth_var(n_rec, pad=64)
idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda") if want_index else None
for _ in range(reps):
    eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
    out, st = eng.fetch()
    print(out, "bytes", n, "scan ms", eng.last_scan_ms(), "GB/s", n / eng.last_scan_ms() / 1e6)

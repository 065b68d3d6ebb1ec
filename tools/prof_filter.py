"""Timing of the record filter (validate_dna / validate_dnan + compaction) on N GiB of synthetic
input resident in HBM: python tools/prof_filter.py [gib] [reps].  CUDA events on the launching stream."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_rs_b200 as fq
from fastq_rs_b200 import _lib

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rb = 321
n = int(gib * (1 << 30)) // rb * rb
eng = fq.Engine(max_len=150)
t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
eng.synth_fixed(t, n)
idx = torch.empty(4 * (n // rb) + 8, dtype=torch.int32, device="cuda")
eng.parse_device(t, n_own=n, n_avail=n, hist=False, index=idx)
out, _ = eng.fetch(want_stats=False)
assert out.status == 0 and out.n_records == n // rb
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for mode, name in ((_lib.KEEP_ALL, "keep_all"), (_lib.KEEP_DNA, "validate_dna"), (_lib.KEEP_DNAN, "validate_dnan")):
    ms = []
    for r in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.filter_device(t, idx, out.n_records, mode, dst)
        e1.record()
        nk, nb = eng.fetch_filter()
        torch.cuda.synchronize()
        if r >= 2:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    # algorithmic bytes: every input byte examined once + index read + kept bytes written
    alg = n + 16 * out.n_records + nb
    print(f"{name:14s} {gib:g} GiB  kept {nk}/{out.n_records} records ({nb / n:.3f} of the bytes)  "
          f"{med:.3f} ms  input {n / med / 1e6:.0f} GB/s  algorithmic {alg / med / 1e6:.0f} GB/s")

"""One-off parity check at BASELINE's full single-GPU size: the 16 GiB synthetic stream is parsed on the GPU
(delimit + index + histograms), copied to the host, and parsed again by the CPU oracle (parallel_each + stats
closure, all host threads); every counter of the statistics block must be equal, and the GPU's line-end index
must equal the closed form of the fixed-length format.   python tools/verify_full_size.py [gib]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fastq_rs_b200 as fq
from oracle import oracle

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 16.0
rb = 321
n = int(gib * (1 << 30)) // rb * rb
n_rec = n // rb
eng = fq.Engine(max_len=150)
t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
eng.synth_fixed(t, n)
idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda")
eng.parse_device(t, n_own=n, n_avail=n, hist=True, index=idx)
out, st = eng.fetch()
assert out.status == 0 and out.finished and out.n_records == n_rec and out.n_lines == 4 * n_rec, out
# index: record k has its line ends at 321 k + {16, 167, 169, 320} (low 32 bits)
k = torch.arange(n_rec, dtype=torch.int64, device="cuda") * rb
for j, off in enumerate((16, 167, 169, 320)):
    want = ((k + off) & 0xFFFFFFFF).to(torch.int64)
    got = idx[j:4 * n_rec:4].to(torch.int64) & 0xFFFFFFFF
    assert bool((got == want).all()), f"index column {j}"
del k, want, got, idx
host = np.empty(n, dtype=np.uint8)
torch.from_numpy(host).copy_(t[:n])
torch.cuda.synchronize()
del t
cores = os.cpu_count() or 1
t0 = time.perf_counter()
rc, ost, _sets = oracle.parallel_each_stats(host, 150, max(1, cores - 1))
dt = time.perf_counter() - t0
assert rc == 0
assert st.n_records == ost.n_records == n_rec and st.n_bases == ost.n_bases == 150 * n_rec
np.testing.assert_array_equal(st.len_hist, ost.len_hist)
np.testing.assert_array_equal(st.base_hist, ost.base_hist)
np.testing.assert_array_equal(st.qual_hist, ost.qual_hist)
print(f"OK: {gib:g} GiB, {n_rec} records: statistics block ({st.words.size} u64 words) bit-exact vs the CPU oracle "
      f"(parallel_each on {cores} threads, {n / dt / 1e9:.2f} GB/s), {4 * n_rec} index entries equal the closed form; "
      f"sum of all quality counters {int(st.qual_hist.sum())}")

"""Illumina-like input: fixed 150 bp reads, headers of VARYING length (tile / x / y coordinates):
   python tools/prof_real.py [n_records] [hist] [index]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastq_rs_b200 as fq
n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 1500000
hist = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
want_index = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
rng = np.random.default_rng(7)
seqs = rng.integers(0, 4, size=(n_rec, 150), dtype=np.uint8)
seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[seqs]
quals = rng.integers(35, 75, size=(n_rec, 150), dtype=np.uint8)
xs = rng.integers(1000, 30000, size=n_rec); ys = rng.integers(1000, 100000, size=n_rec); tiles = rng.integers(1101, 2678, size=n_rec)
parts = []
for i in range(n_rec):
    parts.append(b"@A00123:45:HXXXXDSXX:1:%d:%d:%d 1:N:0:ATCACGTT\n" % (tiles[i], xs[i], ys[i]))
    parts.append(seqs[i].tobytes()); parts.append(b"\n+\n"); parts.append(quals[i].tobytes()); parts.append(b"\n")
data = b"".join(parts)
n = len(data)
eng = fq.Engine(max_len=150)
t = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
t[:n] = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy())
idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda") if want_index else None
for _ in range(3):
    eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
    out, st = eng.fetch()
    print(out.status, out.n_records, "bytes", n, "scan ms", eng.last_scan_ms(), "GB/s", n / eng.last_scan_ms() / 1e6)

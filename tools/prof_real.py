"""Illumina-like input: fixed 150 bp reads, headers of VARYING length (tile / x / y coordinates); a block of
records built on the host is tiled on the device:  python tools/prof_real.py [gib] [hist] [index]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastq_rs_b200 as fq
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
hist = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
want_index = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
n_blk = 200000
rng = np.random.default_rng(7)
seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n_blk, 150), dtype=np.uint8)]
quals = rng.integers(35, 75, size=(n_blk, 150), dtype=np.uint8)
xs = rng.integers(1000, 30000, size=n_blk); ys = rng.integers(1000, 100000, size=n_blk); tiles = rng.integers(1101, 2678, size=n_blk)
parts = []
for i in range(n_blk):
    parts.append(b"@A00123:45:HXXXXDSXX:1:%d:%d:%d 1:N:0:ATCACGTT\n" % (tiles[i], xs[i], ys[i]))
    parts.append(seqs[i].tobytes()); parts.append(b"\n+\n"); parts.append(quals[i].tobytes()); parts.append(b"\n")
block = np.frombuffer(b"".join(parts), dtype=np.uint8)
reps = max(1, int(gib * (1 << 30)) // block.size)
n, n_rec = block.size * reps, n_blk * reps
eng = fq.Engine(max_len=150)
t = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
t[:n] = torch.from_numpy(block.copy()).cuda().repeat(reps)
idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda") if want_index else None
for _ in range(3):
    eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
    out, st = eng.fetch()
    assert out.status == 0 and out.n_records == n_rec
    print(out.status, out.n_records, "bytes", n, "scan ms", eng.last_scan_ms(), "GB/s", n / eng.last_scan_ms() / 1e6)

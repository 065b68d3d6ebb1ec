"""The u16 counter halves of the speculative kernel on a 64 GiB shard with ONE quality symbol (binned-quality-like input:
every record bumps the same 150 counters of a CTA -- 1.45 M records per CTA, 22 x the u16 range): the statistics must be
exactly 64 x the statistics of the 1 GiB block the shard is made of (which stays below 65 535 records per CTA).
   python tools/verify_drain_64GiB.py [n_blocks]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fastq_rs_b200 as fq
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = fq.Engine(max_len=150)
n_rec = (1 << 30) // 321
n = n_rec * 321
blk = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.synth_fixed(blk, n)
blk.view(n_rec, 321)[:, 170:320] = ord("F")            # the quality line: one symbol
pad = torch.zeros(64, dtype=torch.uint8, device="cuda")
one = torch.cat([blk, pad])
eng.parse_device(one, n_own=n, n_avail=n, hist=True)
o1, s1 = eng.fetch()
assert o1.status == 0 and o1.n_records == n_rec and not eng.last_path()["exact"]
big = torch.cat([blk.repeat(reps), pad])
del blk, one
N = n * reps
eng.parse_device(big, n_own=N, n_avail=N, hist=True)
o, s = eng.fetch()
p = eng.last_path()
print(f"{N / (1 << 30):.1f} GiB, {o.n_records} records ({o.n_records // 148} per CTA), scan kernel {eng.last_scan_ms():.2f} ms "
      f"= {N / eng.last_scan_ms() / 1e6:.0f} GB/s, path {p}")
assert o.status == 0 and o.n_records == n_rec * reps and not p["exact"]
assert np.array_equal(s.words[8:], s1.words[8:] * np.uint64(reps)) and s.n_bases == s1.n_bases * reps
print(f"all {s.words.size} statistics words == {reps} x the block's; qual_hist['F'] per position = {int(s.qual_hist[0, ord('F')])}")

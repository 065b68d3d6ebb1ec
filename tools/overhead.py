"""Step overhead: parse_device + fetch on tiny inputs (everything but the scan kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastq_rs_b200 as fq
eng = fq.Engine(max_len=150)
for n_rec in (3000, 300000):
    n = n_rec * 321
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n)
    idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda")
    for hist, index in ((True, idx), (False, None)):
        for _ in range(5):
            eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=index); eng.fetch()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=index); eng.fetch()
        dt = (time.perf_counter() - t0) / 200
        print(f"{n/1e6:8.2f} MB hist={hist} index={index is not None}: {dt*1e6:8.1f} us per parse+fetch, kernel {eng.last_scan_ms()*1e3:.1f} us")

"""One-off wider sweep of tests/test_gpu_parity.py::test_prediction_fuzz (seeds from argv range)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import fastq_rs_b200 as fq
from oracle import oracle
import test_gpu_parity as T
eng = fq.Engine(max_len=150, slot_bytes=1 << 20)
eng300 = fq.Engine(max_len=300, slot_bytes=1 << 20)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    try:
        T.test_prediction_fuzz.__wrapped__(seed, torch, oracle, eng, eng300) if hasattr(T.test_prediction_fuzz, "__wrapped__") else \
            T.test_prediction_fuzz(seed, torch, oracle, eng, eng300)
    except Exception as e:
        bad += 1
        print("FAIL seed", seed, repr(e)[:300])
print("done", hi - lo, "seeds,", bad, "failures")

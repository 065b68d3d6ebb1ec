"""Summarise a kernel timeline written with FQB_TRACE=<file> (see fq_scan.cu trace_ev):
   python tools/trace_view.py gpurun_out/trace.bin [cta] [k0] [n]"""
import sys
import numpy as np
TRACE_K = 2048
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, TRACE_K, 16).astype(np.int64)
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 70
k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 800
n = int(sys.argv[4]) if len(sys.argv) > 4 else 12
names = ["issue", "full", "pass1", "scanned0", "lb_start", "based", "ctl_free", "rec_start", "rec0_free", "rec_free_max", "scanned_max"]
t = a[cta]
ref = t[k0, 0]
print("cta", cta, "clock cycles relative to issue(k0); columns:", names)
for k in range(k0, k0 + n):
    print(k, " ".join(f"{int(t[k, e] - ref):8d}" for e in range(11)))
# steady-state averages over all CTAs
ks = slice(k0, k0 + 400)
v = a[:, ks, :]
ok = (v[:, :, 0] > 0) & (v[:, :, 9] > 0)
def d(x, y):
    return float(np.mean((v[:, :, x] - v[:, :, y])[ok]))
print("mean cycles: issue->full %.0f | full->pass1 %.0f | pass1->scanned(max) %.0f | scanned(max)->rec_start %.0f | "
      "rec_start->rec_free(max) %.0f | issue->rec_free(max) %.0f" % (d(1, 0), d(2, 1), d(10, 2), d(7, 10), d(9, 7), d(9, 0)))
per_tile = np.diff(a[:, k0:k0 + 400, 0], axis=1)
print("mean cycles between consecutive issues: %.0f" % float(np.mean(per_tile[per_tile > 0])))

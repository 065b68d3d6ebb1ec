"""fqb_parse_host from pinned memory with / without the line-end index coming back to the host
(the generic-closure path of a Rust shim):  python tools/prof_host_index.py [gib]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastq_rs_b200 as fq
from fastq_rs_b200 import _lib
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30)) // 321 * 321
eng = fq.Engine(max_len=150, slot_bytes=64 << 20)
L = _lib.lib()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
eng.synth_fixed(d, n)
p = ctypes.c_void_p(); assert L.fqb_host_alloc(n, ctypes.byref(p)) == 0
host = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
torch.from_numpy(host).copy_(d); torch.cuda.synchronize()
n_idx = n // 321 * 4
pi = ctypes.c_void_p(); assert L.fqb_host_alloc(n_idx * 4, ctypes.byref(pi)) == 0     # pinned index buffer
res = _lib.Result(); got = ctypes.c_uint64(0)
for flags, name in ((_lib.F_HIST, "hist, no index"), (_lib.F_INDEX, "index only"), (_lib.F_HIST | _lib.F_INDEX, "hist + index")):
    for r in range(3):
        t0 = time.perf_counter()
        rc = L.fqb_parse_host(eng.ctx, p, n, 0, flags, ctypes.byref(res), None, pi if flags & _lib.F_INDEX else None,
                              n_idx if flags & _lib.F_INDEX else 0, ctypes.byref(got))
        dt = time.perf_counter() - t0
        assert rc == 0 and res.status == 0 and res.n_records == n // 321, (rc, res.status)
    print(f"{name:16s} {gib:g} GiB  {dt * 1e3:8.2f} ms  {n / dt / 1e9:6.1f} GB/s  index entries {got.value}")
idx = np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_uint32)), shape=(n_idx,))
exp = (np.arange(8, dtype=np.uint64) // 4 * 321 + np.array([16, 167, 169, 320] * 2, dtype=np.uint64)).astype(np.uint32)
assert (idx[:8] == exp).all() and int(idx[-1]) == (n - 1) & 0xFFFFFFFF

#!/usr/bin/env python
"""bench.py -- FASTQ GB/s (delimit + per-position base/quality histograms) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over one batch of synthetic 150 bp FASTQ (SURVEY.md 8(d),
input A): newline / record-boundary scan with '@'/'+' validation, the line-end index, and the
per-position ACGTN + quality-byte histograms, fused in one sm_100a kernel (fq_stream.cu).

  value     whole-job GB/s with the bytes already resident in HBM (CUDA events, max over ranks).  Two engine
            contexts take the steps in turn: step k+1 is enqueued behind step k's kernels before the host reads
            step k's result, so the read-back overlaps the next step's kernels (--no-pipeline: one context)
  configs   (N = 1) every BASELINE configuration through the same API, one context
  e2e       same metric through the host API (fqb_parse_host): pinned host bytes -> H2D -> kernels
            -> D2H of the outcome + stats block, every step
  roofline  the scan kernel against the measured HBM bandwidth (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the CPU oracle's parallel_each (C restatement of
            src/lib.rs:509-566; the Rust crate cannot be built here) on the box's host cores
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
REC_BYTES = 17 + 2 * (READ_LEN + 1) + 2  # 321
GIB = 1 << 30


def profiled_traffic(gib: float):
    """(dram read + write bytes per launch of the dominant kernel, source) from the committed ncu --set full
    capture (profiles/), valid for the workload it was captured on (16 GiB)."""
    for name in ("r02_stream_kernel_ncu_full_16GiB.txt", "r01_stream_kernel_ncu_full_16GiB.txt"):
        path = os.path.join(ROOT, "profiles", name)
        if abs(gib - 16.0) > 1e-9 or not os.path.exists(path):
            continue
        for line in open(path):
            if line.startswith("traffic = dram read + write per launch"):
                return float(line.split()[-1]), "ncu --set full, profiles/" + name
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  The region is tens of milliseconds long,
    so the samples come from NVML directly (one every ~2 ms); `nvidia-smi` (the B200_PROFILING.md
    recipe; ~100 ms per call) is the fallback when pynvml is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.source = "nvidia-smi"
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nv = (pynvml, h, float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self.source = "nvml"
        except Exception:
            self._nv = None

    def _sample_nvml(self):
        nv, h, mx = self._nv
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.rows.append([sm, mx, None] + ["Active" if r & bits[k] else "Not Active" for k in
                                           ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nv is not None:
                    self._sample_nvml()
                    self._stop.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, TypeError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if str(v).lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def gen_host_sample(n_bytes: int, threads: int) -> np.ndarray:
    """Synthetic input A on the host with the oracle's generator (threads: ctypes drops the GIL)."""
    from oracle import oracle
    n_bytes = n_bytes // REC_BYTES * REC_BYTES
    out = np.empty(n_bytes, dtype=np.uint8)
    L = oracle.lib()
    nrec = n_bytes // REC_BYTES
    per = (nrec + threads - 1) // threads

    def work(k):
        a, b = k * per * REC_BYTES, min(nrec, (k + 1) * per) * REC_BYTES
        if b > a:
            L.fqo_synth_fixed(oracle.SEED, READ_LEN, a, b - a, out.ctypes.data + a)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def cpu_parallel_each(sample: np.ndarray, workers: int, reps: int = 1):
    """The reference's CPU path (parallel_each + stats closure) via the oracle; GB/s."""
    from oracle import oracle
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, st, _ = oracle.parallel_each_stats(sample, READ_LEN, workers)
        dt = time.perf_counter() - t0
        assert rc == 0 and st.n_records == sample.size // REC_BYTES
        best = dt if best is None else min(best, dt)
    return sample.size / best / 1e9


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    workers = max(1, cores - 1)          # + the serial producer thread (src/lib.rs:535)
    sample_bytes = int(args.cpu_sample_gib * GIB)
    sample = gen_host_sample(sample_bytes, cores)
    for _ in range(args.warmup):
        cpu_parallel_each(sample[: min(sample.size, 64 << 20) // REC_BYTES * REC_BYTES], workers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_parallel_each(sample, workers)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample.size / dt / 1e9
    desc = f"{sample.size / GIB:.2f} GiB prefix of the synthetic 150 bp stream per step, in host RAM"
    line = {
        "impl": "reference", "metric": "fastq_delimit_hist_GBps", "value": v, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": workers + 1, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle = C restatement of fastq-rs parallel_each (serial delimiting + N stats workers); "
                "the Rust crate cannot be built in this image",
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {"workload": f"{args.gib:g} GiB/GPU in-HBM synthetic 150 bp fixed-length FASTQ: delimit + line-end index "
                        "+ per-position ACGTN and quality histograms (BASELINE configs[1]+[2], one fused pass)",
            "read_len": READ_LEN, "record_bytes": REC_BYTES, "bytes_per_gpu": int(args.gib * GIB),
            "sharding": f"byte-chunk x{n_gpus}" if n_gpus > 1 else "none",
            "l2": "inputs (GiBs) far larger than the 126 MB L2; no flush needed",
            "pipeline": "1 context: every step is read back before the next is enqueued" if args.no_pipeline else
                        "2 engine contexts take the steps in turn: step k+1 is enqueued behind step k's kernels before the "
                        "host reads step k's result (the GPU still runs the steps one after the other)"}


def sharded_bit_exact_check(fq, eng, sp, dist, world, rank, dev):
    """Before anything is timed at N > 1: a small stream (cuts mid-record, a shard size that is no multiple of the
    record size) parsed by the N ranks together -- NCCL all-reduce of [block | outcome slots] -- must give, bit
    for bit, the statistics block and outcome of ONE parse of the whole stream on rank 0's GPU."""
    import torch
    from fastq_rs_b200.sharded import ShardSpec
    shard = (6 << 20) + 4096 + 16 * 7                      # bytes per rank
    total = world * shard // REC_BYTES * REC_BYTES
    a, b = rank * shard, min(total, (rank + 1) * shard)
    halo = min(total - b, fq._lib.MAX_RECORD_BYTES)
    front = 16 if a > 0 else 0
    buf = torch.empty(16 + (b - a) + halo + 64, dtype=torch.uint8, device=dev)
    eng.synth_fixed(buf.data_ptr() + 16 - front, (b - a) + halo + front, byte_off=a - front, read_len=READ_LEN)
    idx = torch.empty(4 * ((b - a) // REC_BYTES + 4), dtype=torch.int32, device=dev)
    c0 = sp.collectives
    out, st = sp.parse(ShardSpec(buf[16 - front:], a, b, halo, front, is_last=(b + halo == total)), hist=True, index=idx)
    assert sp.collectives - c0 == 1 and sp.reparsed == 0, "the common case is one collective"
    ok = torch.ones(1, dtype=torch.int64, device=dev)
    if rank == 0:
        whole = torch.empty(total + 64, dtype=torch.uint8, device=dev)
        eng.synth_fixed(whole, total, read_len=READ_LEN)
        eng.parse_device(whole, n_own=total, n_avail=total, hist=True)
        out1, st1 = eng.fetch()
        same = (out.status, out.n_records, out.n_lines, out.finished) == (out1.status, out1.n_records, out1.n_lines, out1.finished)
        same = same and bool(np.array_equal(st.words, st1.words)) and out1.n_records == total // REC_BYTES
        ok[0] = 1 if same else 0
    dist.broadcast(ok, src=0)
    assert int(ok.item()) == 1, "N-rank result differs from the single-GPU parse of the same stream"
    return {"bytes": total, "records": total // REC_BYTES, "words_compared": int(st.words.size), "collectives": 1}


def time_config(eng, torch, steps, warmup, parse, n_bytes, n_rec, peak, index_entries):
    """One BASELINE configuration on one GPU: (GB/s over whole steps, kernel ms, roofline fraction)."""
    for _ in range(max(1, warmup)):
        parse()
        out, _ = eng.fetch()
    assert out.status == 0 and out.n_records == n_rec, out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k, ki = [], []
    e0.record()
    for _ in range(steps):
        parse()
        eng.fetch()
        k.append(eng.last_scan_ms())
        ki.append(eng.last_index_ms())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    k_ms, i_ms = float(np.mean(k)), float(np.mean(ki))
    alg = n_bytes + 4 * index_entries
    # the scan kernel reads every byte; the dense index (16 B / record) is written by the index kernel behind it
    # from the scan kernel's window descriptors: the algorithmic bytes are charged to the two together
    return {"value": n_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms, "kernel_ms": k_ms, "index_kernel_ms": i_ms,
            "algorithmic_bytes": alg, "roofline_frac": alg / ((k_ms + i_ms) * 1e-3) / 1e9 / peak,
            "scan_kernel_frac": n_bytes / (k_ms * 1e-3) / 1e9 / peak, "records": n_rec}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import fastq_rs_b200 as fq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    eng = fq.Engine(max_len=READ_LEN, device=local, slot_bytes=args.slot_mib << 20, n_slots=3)
    peak, peak_src = peaks()

    # ---- the stream: world x gib GiB of synthetic input A, sharded by byte chunk ------------
    shard = int(args.gib * GIB) // 16 * 16
    total = world * shard // REC_BYTES * REC_BYTES          # whole records overall
    a, b = rank * shard, min(total, (rank + 1) * shard)
    halo = min(total - b, fq._lib.MAX_RECORD_BYTES)
    n_own, n_avail = b - a, b - a + halo
    buf = torch.empty(16 + n_avail + 64, dtype=torch.uint8, device=dev)
    front = 16 if a > 0 else 0
    eng.synth_fixed(buf.data_ptr() + 16 - front, n_avail + front, byte_off=a - front, read_len=READ_LEN)
    data = buf[16:]
    n_rec_upper = n_avail // REC_BYTES + 2
    index = torch.empty(4 * n_rec_upper, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    # N > 1: every rank parses its shard at once with an inferred start; ONE all-reduce of [statistics block |
    # outcome slots] (the library's own NCCL call, fqb_allreduce) both sums the statistics and gathers the
    # outcomes that confirm the inferred line numbers; one device-to-host copy.  No byte is read twice.
    from fastq_rs_b200.sharded import ShardedParser, ShardSpec
    sp = ShardedParser(eng, dist=dist if world > 1 else None, device=dev)
    spec = ShardSpec(buf[16 - front:], a, b, halo, front, is_last=(b + halo == total))
    check = sharded_bit_exact_check(fq, eng, sp, dist, world, rank, dev) if world > 1 else None

    # Two engine contexts take the steps in turn: step k + 1 is enqueued (on its own stream, behind step k's kernels)
    # before the host waits for step k's result, so the copy of the result and the host-side bookkeeping of one step
    # overlap the kernels of the next.  On the GPU the steps still run one after the other.
    depth = 1 if args.no_pipeline else 2
    engs, sps, idxs = [eng], [sp], [index]
    if depth == 2:
        engs.append(fq.Engine(max_len=READ_LEN, device=local, slot_bytes=args.slot_mib << 20, n_slots=3))
        sps.append(ShardedParser(engs[1], dist=dist if world > 1 else None, device=dev))
        idxs.append(torch.empty_like(index))
    streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
    for s_ in streams:
        s_.wait_stream(torch.cuda.current_stream())
    scan_ms, index_ms = [], []

    def begin(k):
        """Enqueue one pass of the hot path over this rank's shard (+ the one collective when sharded)."""
        j = k % depth
        if depth == 2:
            streams[j].wait_stream(streams[j ^ 1])
        with torch.cuda.stream(streams[j]):
            if world > 1:
                sps[j].begin(spec, hist=True, index=idxs[j])
            else:
                engs[j].parse_device(data, n_own=n_own, n_avail=n_avail, hist=True, index=idxs[j], line_base=0,
                                     stream_offset=a, line_start=True, front16=False, eof=True)

    def end(k):
        """Wait for step k: its global (Outcome, Stats) on the host."""
        j = k % depth
        with torch.cuda.stream(streams[j]):
            r = sps[j].finish() if world > 1 else engs[j].fetch()
        scan_ms.append(engs[j].last_scan_ms())
        index_ms.append(engs[j].last_index_ms())
        return r

    def run_steps(n):
        r = None
        begin(0)
        for k in range(1, n):
            begin(k)
            r = end(k - 1)
            assert r[0].status == 0, r[0]
        return end(n - 1)

    out, st = run_steps(max(args.warmup, depth))
    assert out.status == 0, out
    assert all(x.reparsed == 0 for x in sps), "an inferred shard start was not confirmed"
    assert st.n_records == total // REC_BYTES, (st.n_records, total // REC_BYTES)
    assert int(st.qual_hist.sum()) == READ_LEN * st.n_records

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: HBM-resident ---------------------------------------------------------
    launches0 = sum(e.launch_count() for e in engs)
    coll0 = sum(x.collectives for x in sps)
    scan_ms.clear()
    index_ms.clear()
    with ClockSampler(local) as clk:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(args.steps)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(e.launch_count() for e in engs) - launches0
    colls = sum(x.collectives for x in sps) - coll0
    for e in engs[1:]:
        e.close()
    del idxs[1:], sps[1:], engs[1:]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = total / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel, live CUDA events on its stream ----------------------
    k_ms, i_ms = float(np.mean(scan_ms)), float(np.mean(index_ms))
    alg_bytes = n_own + 16 * (n_own // REC_BYTES)            # 1 B read per input byte + 16 B index per record
    # The scan kernel (fq_stream_kernel) reads every input byte and leaves 16-byte window descriptors; the dense
    # 16 B / record index is written by the index kernel right behind it (fq_stream_compact_kernel).  The
    # algorithmic bytes of the step are charged to the two kernels TOGETHER; `scan_kernel` below has the scan
    # kernel alone against the bytes it moves itself (the input).
    achieved = alg_bytes / ((k_ms + i_ms) * 1e-3) / 1e9
    traffic, traffic_src = profiled_traffic(args.gib)
    roofline = {"bound": "hbm", "kernel": "fq_stream_kernel<SCfg<5,28,4096>, HIST, predicting variant> + fq_stream_compact_kernel (index)",
                "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": args.traffic_bytes if args.traffic_bytes is not None else traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms": k_ms + i_ms, "algorithmic_bytes": alg_bytes,
                "kernels": {"fq_stream_kernel": {"ms": k_ms, "bytes": n_own, "GBps": n_own / (k_ms * 1e-3) / 1e9,
                                                 "frac": n_own / (k_ms * 1e-3) / 1e9 / peak},
                            "fq_stream_compact_kernel": {"ms": i_ms, "bytes": 16 * (n_own // REC_BYTES),
                                                         "GBps": 16 * (n_own // REC_BYTES) / (max(i_ms, 1e-9) * 1e-3) / 1e9}}}

    # ---- every BASELINE configuration, one GPU each (N = 1 line only) --------------------------
    configs = None
    if world == 1 and not args.no_configs:
        n_rec = n_own // REC_BYTES
        cs, cw = max(3, min(args.steps, 5)), 2

        def dev_parse(hist, idx):
            return lambda: eng.parse_device(data, n_own=n_own, n_avail=n_avail, hist=hist, index=idx)
        configs = []
        for name, hist, idx in (("configs[1]: delimit + record count", False, None),
                                ("configs[1]': delimit + line-end index", False, index),
                                ("configs[2]: per-position ACGTN + quality histograms", True, None),
                                ("configs[1]+[2] fused: delimit + index + histograms (the headline)", True, index)):
            c = time_config(eng, torch, cs, cw, dev_parse(hist, idx), n_own, n_rec, peak, 4 * n_rec if idx is not None else 0)
            c["workload"] = f"{args.gib:g} GiB fixed 150 bp, {name}"
            configs.append(c)
        # configs[3]: variable 50-300 bp (P = 300), same bytes per GPU
        del index
        eng3 = fq.Engine(max_len=300, device=local)
        n_var = int(args.gib * GIB / 371.3)
        vbuf, vbytes = eng3.synth_var(n_var, pad=64)
        vidx = torch.empty(4 * n_var + 8, dtype=torch.int32, device=dev)
        for name, hist, idx in (("configs[3]: delimit + index + histograms", True, vidx),
                                ("configs[3]': delimit + index", False, vidx)):
            c = time_config(eng3, torch, cs, cw, (lambda h=hist, i=idx: eng3.parse_device(vbuf, n_own=vbytes, n_avail=vbytes, hist=h, index=i)),
                            vbytes, n_var, peak, 4 * n_var)
            c["workload"] = f"{vbytes / GIB:.2f} GiB variable-length 50-300 bp, {name}"
            configs.append(c)
        eng3.close()
        del vbuf, vidx
        # not a BASELINE configuration, but what sequencers write: fixed 150 bp reads whose id lines vary in length
        # (tile / x / y coordinates), so no two records start a fixed distance apart
        rng = np.random.default_rng(7)
        n_blk = 100000
        seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n_blk, READ_LEN), dtype=np.uint8)]
        quals = rng.integers(35, 75, size=(n_blk, READ_LEN), dtype=np.uint8)
        xs, ys, tiles = rng.integers(1000, 30000, size=n_blk), rng.integers(1000, 100000, size=n_blk), rng.integers(1101, 2678, size=n_blk)
        parts = []
        for i in range(n_blk):
            parts += [b"@A00123:45:HXXXXDSXX:1:%d:%d:%d 1:N:0:ATCACGTT\n" % (tiles[i], xs[i], ys[i]), seqs[i].tobytes(), b"\n+\n",
                      quals[i].tobytes(), b"\n"]
        block = np.frombuffer(b"".join(parts), dtype=np.uint8)
        reps = max(1, int(min(args.gib, 8.0) * GIB) // block.size)
        n_ill, rec_ill = block.size * reps, n_blk * reps
        ibuf = torch.zeros(n_ill + 64, dtype=torch.uint8, device=dev)
        ibuf[:n_ill] = torch.from_numpy(block.copy()).to(dev).repeat(reps)
        iidx = torch.empty(4 * rec_ill + 8, dtype=torch.int32, device=dev)
        c = time_config(eng, torch, cs, cw, lambda: eng.parse_device(ibuf, n_own=n_ill, n_avail=n_ill, hist=True, index=iidx),
                        n_ill, rec_ill, peak, 4 * rec_ill)
        c["workload"] = f"{n_ill / GIB:.2f} GiB fixed 150 bp with Illumina-like ids of varying length (not a BASELINE configuration), delimit + index + histograms"
        configs.append(c)
        del ibuf, iidx
        index = torch.empty(4 * n_rec_upper, dtype=torch.int32, device=dev)

    # ---- end to end through the host API (rank-local; pinned host bytes) ---------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, eng, data, n_own if world == 1 else n_own // REC_BYTES * REC_BYTES, world, dev, a)

    line = None
    if rank == 0:
        line = {
            "metric": "fastq_delimit_hist_GBps", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "clocks": clk.summary(),
            "gpu_launches": int(launches), "records": total // REC_BYTES,
        }
        if configs is not None:
            line["configs"] = configs
        if world > 1:
            line["collectives_per_step"] = colls / args.steps
            line["sharded_check"] = check
    if e2e is not None:
        ev = torch.tensor([e2e["bytes"], e2e["seconds"]], dtype=torch.float64, device=dev)
        if world > 1:
            tot_b = ev[0:1].clone()
            mx_s = ev[1:2].clone()
            dist.all_reduce(tot_b, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx_s, op=dist.ReduceOp.MAX)
            ev = torch.cat([tot_b, mx_s])
        if rank == 0:
            line["e2e"] = {"value": float(ev[0].item()) / float(ev[1].item()) / 1e9, "unit": "GB/s",
                           "h2d_bytes_per_step": int(e2e["bytes"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "bytes_per_gpu": int(e2e["bytes"]), "api": "fqb_parse_host(FQB_F_HIST | FQB_F_INDEX): pinned ring of 3 slots; outcome, statistics block and the "
                                  "line-end index (pinned buffer) all land on the host",
                           "sample": f"{e2e['bytes'] / GIB:.2f} GiB of each rank's shard, pinned host memory"}
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = max(1, cores - 1)
        sample = gen_host_sample(int(args.cpu_sample_gib * GIB), cores)
        v = cpu_parallel_each(sample, workers, reps=2)
        line["cpu_baseline"] = {
            "value": v, "unit": "GB/s", "cores": workers + 1, "kind": "port",
            "sample": f"{sample.size / GIB:.2f} GiB prefix of the same synthetic stream, host RAM, best of 2; "
                      f"oracle parallel_each({workers}) + stats closure"}
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_e2e(args, eng, data, n, world, dev, stream_off):
    """fqb_parse_host on pinned host memory holding this rank's bytes (whole records)."""
    import torch
    from fastq_rs_b200 import _lib
    L = _lib.lib()
    n = min(n, int(args.e2e_gib * GIB)) // REC_BYTES * REC_BYTES
    # this rank's shard may start mid-record: skip to its first record start so the host stream is a valid file
    skip = (-stream_off) % REC_BYTES
    n = min(n, (data.numel() - skip)) // REC_BYTES * REC_BYTES
    p = ctypes.c_void_p()
    while n > 0 and L.fqb_host_alloc(n, ctypes.byref(p)) != 0:
        n = n // 2 // REC_BYTES * REC_BYTES
    if n <= 0:
        return None
    host = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
    torch.from_numpy(host).copy_(data[skip:skip + n])
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, args.e2e_steps))
    # the whole result comes back to the host: outcome, statistics block and the line-end index (4 x u32 per
    # record, written into a pinned buffer by the device chunk by chunk)
    n_idx = n // REC_BYTES * 4
    pi = ctypes.c_void_p()
    if L.fqb_host_alloc(n_idx * 4, ctypes.byref(pi)) != 0:
        L.fqb_host_free(p)
        return None
    words = np.zeros(eng.n_words, dtype=np.uint64)
    res, got = _lib.Result(), ctypes.c_uint64(0)

    def step():
        rc = L.fqb_parse_host(eng.ctx, p, n, 0, _lib.F_HIST | _lib.F_INDEX, ctypes.byref(res), words.ctypes.data,
                              pi, n_idx, ctypes.byref(got))
        assert rc == 0 and res.status == 0 and res.n_records == n // REC_BYTES and got.value == n_idx, (rc, res.status)

    step()                                                # warm-up (allocates the ring)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    idx = np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_uint32)), shape=(n_idx,))
    assert int(words[0]) == n // REC_BYTES and int(idx[3]) == REC_BYTES - 1 and int(idx[-1]) == (n - 1) & 0xFFFFFFFF
    L.fqb_host_free(pi)
    L.fqb_host_free(p)
    return {"bytes": n, "seconds": dt, "d2h": eng.n_words * 8 + ctypes.sizeof(_lib.Result) + n_idx * 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=16.0, help="GiB of FASTQ per GPU")
    ap.add_argument("--e2e-gib", type=float, default=None,
                    help="GiB per GPU streamed from pinned host memory in the e2e leg; default: the whole shard of every "
                         "rank (halved until the pinned allocation succeeds, and reported)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--slot-mib", type=int, default=64)
    ap.add_argument("--cpu-sample-gib", type=float, default=2.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="one engine context: read every step back before the next")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration list (N = 1 line)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes/launch from the committed ncu --set full capture (profiles/)")
    args = ap.parse_args()
    if args.e2e_gib is None:
        args.e2e_gib = args.gib
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

"""The N-rank driver (fastq_rs_b200/sharded.py) on CPU: world_size-2 gloo, oracle-backed stand-in
engine.  Checks the protocol -- inferred shard starts confirmed by the exact prefix of the newline
counts, re-parse when an inference is not confirmed, first error in stream order wins, one
all_reduce of the statistics block -- against the single-stream oracle result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, data_bytes, mode, q):
    from fake_engine import FakeEngine, stats_words
    from fastq_rs_b200.sharded import ShardedParser, ShardSpec, shard_bounds, MAX_RECORD_BYTES
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    data = np.frombuffer(data_bytes, dtype=np.uint8)
    total = data.size
    a, b = shard_bounds(total, world)[rank]
    halo = min(total - b, MAX_RECORD_BYTES)
    front = 16 if a > 0 else 0
    t = torch.from_numpy(data[a - front:b + halo].copy())
    eng = FakeEngine(150, lie_phase=(mode == "lie" and rank == 1), fail_infer=(mode == "nophase"))
    sp = ShardedParser(eng, dist=dist)
    spec = ShardSpec(t, a, b, halo, front, is_last=(b + halo == total))
    if mode == "two_in_flight":
        # begin / finish: a second parser (its own engine context) takes the next step before the first is read
        eng2 = FakeEngine(150)
        sp2 = ShardedParser(eng2, dist=dist)
        sp.begin(spec)
        sp2.begin(spec)
        out, st = sp.finish()
        out2, st2 = sp2.finish()
        assert out2 == out and np.array_equal(st2.words, st.words)
        assert (sp2.reparsed, sp2.collectives) == (0, 1)
    else:
        out, st = sp.parse(spec)
    q.put((rank, out, st.words.copy(), eng.n_parses, (sp.reparsed, sp.collectives)))
    dist.barrier()
    dist.destroy_process_group()


def run_world(data: bytes, mode: str, tmp_path, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    init_file = str(tmp_path / f"init_{mode}")
    procs = [ctx.Process(target=_worker, args=(r, world, init_file, data, mode, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def synth(n_rec):
    from oracle import oracle
    return oracle.synth_fixed(n_rec * 321).tobytes()


def expect(data: bytes):
    from fake_engine import stats_words
    from oracle import oracle
    res, st = oracle.each_stats(data, 150)
    return res, stats_words(150, st, res.n_records)


@pytest.mark.parametrize("mode", ["plain", "lie", "nophase", "two_in_flight"])
def test_two_ranks_match_single_stream(mode, tmp_path):
    data = synth(257)                      # the cut falls inside a record
    res, words = expect(data)
    out = run_world(data, mode, tmp_path)
    for rank, o, w, n_parses, reparsed in out:
        assert (o.status, o.n_records, o.n_lines) == (0, res.n_records, data.count(b"\n"))
        assert np.array_equal(w, words)
        if rank == 1:
            assert n_parses == (1 if mode in ("plain", "two_in_flight") else 2), (mode, n_parses)
        else:
            assert n_parses == 1
        # the common case is ONE collective: the all-reduce of [block | outcome slots] is also the gather
        assert reparsed[1] == 1 if mode in ("plain", "two_in_flight") else reparsed[1] > 1


def test_error_in_first_shard_silences_the_second(tmp_path):
    data = bytearray(synth(300))
    data[321 * 40] = ord("X")              # record 40 loses its '@'
    res, words = expect(bytes(data))
    assert res.status == 1 and res.n_records == 40
    out = run_world(bytes(data), "plain", tmp_path)
    for rank, o, w, _, _ in out:
        assert (o.status, o.n_records, o.err_offset) == (1, 40, 321 * 40)
        assert np.array_equal(w, words)


def test_error_in_second_shard(tmp_path):
    data = bytearray(synth(300))
    data[321 * 250 + 17 + 151] = ord("-")  # record 250: separator line does not start with '+'
    res, words = expect(bytes(data))
    assert res.status == 2 and res.n_records == 250
    out = run_world(bytes(data), "plain", tmp_path)
    for rank, o, w, _, _ in out:
        assert (o.status, o.n_records, o.err_offset) == (2, 250, 321 * 250)
        assert np.array_equal(w, words)


def test_shard_bounds():
    from fastq_rs_b200.sharded import shard_bounds
    for total in (0, 1, 15, 16, 17, 1000, 12345):
        for world in (1, 2, 3, 8):
            sh = shard_bounds(total, world)
            assert sh[0][0] == 0 and sh[-1][1] == total
            assert all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
            assert all(a % 16 == 0 or a == total for a, _ in sh)

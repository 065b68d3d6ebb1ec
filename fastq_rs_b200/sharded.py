"""N-rank driver of the delimit + statistics path: one process per GPU, one byte shard per rank.

The reference has nothing to shard (its `parallel_each` delimits on ONE thread, src/lib.rs:535);
this is the B200-side answer to "the path shards by byte chunk" (SURVEY.md 8(e)):

  * rank g holds bytes [a_g, b_g) of the stream in HBM, plus a halo of <= 68 KiB of its successor;
    a record belongs to the shard it STARTS in (the device-side form of Buffer's carry-over,
    src/buffer.rs:30-72);
  * which lines of a shard are headers depends on the number of '\\n' in front of it.  Every rank
    parses at once with FQB_F_INFER_START (the kernel infers where its first record starts from the
    grammar and reports the line phase that implies), then ONE all-gather of a few words per rank
    (newline count, inferred phase, status) gives every rank its exact line number; an inference
    that does not agree with it (or FQB_E_PHASE) is redone with the exact line_base -- so the result
    never depends on the inference, and in the common case no byte is read twice;
  * errors: the first bad record in stream order wins (Parser::each delivers everything before it,
    src/lib.rs:226-237): shards behind it contribute nothing;
  * ONE all_reduce(sum) of the u64 statistics block.

`torch.distributed` is plumbing only (NCCL on the GPUs, gloo in the CPU tests).  The engine object
is duck-typed (`parse_device`, `fetch`, `device_stats`) so that the protocol can be tested on CPU
with a stand-in; the product engine is `fastq_rs_b200.Engine`.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import E_PHASE, OK
from .engine import Outcome, Stats

MAX_RECORD_BYTES = 68 * 1024


def shard_bounds(total: int, world: int, align: int = 16) -> list[tuple[int, int]]:
    """Contiguous byte shards [a, b) of a stream of `total` bytes, cut at multiples of `align`."""
    per = -(-total // world)
    per = -(-per // align) * align
    out = []
    for g in range(world):
        a = min(total, g * per)
        b = min(total, (g + 1) * per)
        out.append((a, b))
    return out


@dataclass
class ShardSpec:
    """One rank's piece of the stream: `data` holds bytes [a - front, b + halo) where front is 16
    for every shard but the first (the byte in front of a shard tells whether it starts a line)."""
    data: object            # uint8 tensor (CUDA for the product engine)
    a: int                  # stream offset of the first owned byte
    b: int                  # stream offset behind the last owned byte
    halo: int               # readable bytes behind b (<= MAX_RECORD_BYTES)
    front: int              # 16 if data starts 16 bytes in front of a, else 0
    is_last: bool           # nothing of the stream exists behind b + halo


class ShardedParser:
    def __init__(self, engine, dist=None, group=None, device=None):
        self.eng = engine
        self.dist = dist
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.reparsed = 0   # shards parsed twice so far (inference not confirmed)

    # -- collectives (a few words per rank; the stats block) ---------------------------------
    def _all_gather_words(self, words: list[int]) -> np.ndarray:
        import torch
        mine = torch.tensor(words, dtype=torch.int64, device=self.device)
        if self.dist is None or self.world == 1:
            return mine.cpu().numpy().reshape(1, -1)
        out = torch.empty(self.world * len(words), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return out.cpu().numpy().reshape(self.world, -1)

    def _parse(self, sh: ShardSpec, hist, index, line_base, infer):
        view = sh.data[sh.front:] if sh.front else sh.data
        if sh.b == sh.a:   # an empty shard (more ranks than 16-byte pieces): nothing to own, nothing to read
            self.eng.parse_device(view, n_own=0, n_avail=0, hist=hist, index=None, line_base=line_base,
                                  stream_offset=sh.a, line_start=False, front16=False, eof=sh.is_last)
            return self.eng.fetch(want_stats=False)[0]
        self.eng.parse_device(view, n_own=sh.b - sh.a, n_avail=sh.b - sh.a + sh.halo, hist=hist, index=index,
                              line_base=line_base, stream_offset=sh.a, line_start=(sh.a == 0),
                              front16=(sh.front == 16), eof=sh.is_last, infer_start=infer)
        return self.eng.fetch(want_stats=False)[0]

    def _enqueue(self, sh: ShardSpec, hist, index, line_base, infer):
        """parse_device without waiting for it"""
        view = sh.data[sh.front:] if sh.front else sh.data
        if sh.b == sh.a:
            self.eng.parse_device(view, n_own=0, n_avail=0, hist=hist, index=None, line_base=line_base,
                                  stream_offset=sh.a, line_start=False, front16=False, eof=sh.is_last)
        else:
            self.eng.parse_device(view, n_own=sh.b - sh.a, n_avail=sh.b - sh.a + sh.halo, hist=hist, index=index,
                                  line_base=line_base, stream_offset=sh.a, line_start=(sh.a == 0),
                                  front16=(sh.front == 16), eof=sh.is_last, infer_start=infer)

    def parse(self, sh: ShardSpec, hist: bool = True, index=None):
        """Delimit (+histograms) the whole stream; every rank returns the GLOBAL (Outcome, Stats).
        `index` (optional int32 tensor) receives this rank's line ends (low 32 bits of the stream
        offsets of the '\n' in [a, b)).

        Common case = ONE host synchronisation: the parse, an all-gather of the 8-word device-resident
        outcomes, and an all-reduce of a copy of the statistics block are enqueued back to back; the
        host then reads both.  Only if some outcome shows an error, an unconfirmed inference or
        FQB_E_PHASE is the careful path below taken (the local statistics block is still intact)."""
        import torch
        infer = sh.a != 0 and sh.b > sh.a
        fast = hasattr(self.eng, "device_result") and self.dist is not None and self.world > 1
        if fast:
            self._enqueue(sh, hist, index, 0, infer)
            res_all = torch.empty(self.world * 8, dtype=torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(res_all, self.eng.device_result(), group=self.group)
            summed = self.eng.device_stats().clone()
            self.dist.all_reduce(summed, op=self.dist.ReduceOp.SUM, group=self.group)
            g8 = res_all.cpu().numpy().reshape(self.world, 8)          # the one synchronisation
            lb = np.concatenate([[0], np.cumsum(g8[:-1, 3])])          # exact line numbers: prefix of n_lines
            if bool((g8[:, 0] == OK).all()) and self._phases_ok(g8, lb):
                words = summed.cpu().numpy().view(np.uint64).copy()
                total = Outcome(status=OK, finished=bool(g8[-1, 1]), n_records=int(g8[:, 2].sum()),
                                n_lines=int(g8[:, 3].sum()), err_offset=0, tail_offset=None, line_phase=0)
                return total, Stats(self.eng.max_len, words)
            out = self.eng.fetch(want_stats=False)[0]
        else:
            out = self._parse(sh, hist, index, 0, infer)
        return self._careful(sh, hist, index, infer, out)

    def _phases_ok(self, g8, lb) -> bool:
        """every inferring shard (all but the first, unless it owns nothing) reports the phase its exact
        line number has; shards that own nothing (n_lines == 0 and n_records == 0) are exempt"""
        for r in range(1, self.world):
            empty = g8[r, 3] == 0 and g8[r, 2] == 0
            if not empty and (int(g8[r, 6]) & 3) != (int(lb[r]) & 3):
                return False
        return True

    def _careful(self, sh, hist, index, infer, out):
        """The general protocol: confirm or redo the inference, first error in stream order wins."""
        def words(o):
            return [o.status, o.err_offset, o.n_records, o.n_lines, int(o.finished), o.line_phase]
        g = self._all_gather_words(words(out))
        line_base = int(g[:self.rank, 3].sum())        # exact line number of this shard: prefix of n_lines
        if bool((g[:, 0] == E_PHASE).any()):
            # some shard could not even deliver its newline count: every rank counts (one cheap pass)
            view = sh.data[sh.front:] if sh.front else sh.data
            n = self.eng.count_lines(view, sh.b - sh.a) if sh.b > sh.a else 0
            g2 = self._all_gather_words([n])
            line_base = int(g2[:self.rank, 0].sum())
        unconfirmed = infer and (out.status == E_PHASE or out.line_phase != (line_base & 3))
        if unconfirmed:
            # the inference is not confirmed: parse again with the exact line number
            self.reparsed += 1
            out = self._parse(sh, hist, index, line_base, False)
        # did anybody parse again?  (every rank can tell from the gathered phases and counts)
        lb = np.concatenate([[0], np.cumsum(g[:-1, 3])]) if not bool((g[:, 0] == E_PHASE).any()) else None
        anybody = bool((g[:, 0] == E_PHASE).any()) or bool(((g[1:, 5] & 3) != (lb[1:] & 3)).any())
        if anybody:
            g = self._all_gather_words(words(out))
        # first error in stream order wins; shards behind it contribute nothing
        bad = np.nonzero(g[:, 0] != OK)[0]
        first_bad = int(bad[0]) if bad.size else None
        stats_dev = self.eng.device_stats()
        if first_bad is not None and self.rank > first_bad:
            stats_dev.zero_()
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(stats_dev, op=self.dist.ReduceOp.SUM, group=self.group)
        stats_words = stats_dev.cpu().numpy().view(np.uint64).copy()
        upto = self.world if first_bad is None else first_bad + 1
        total = Outcome(status=int(g[first_bad, 0]) if first_bad is not None else OK,
                        finished=first_bad is None and bool(g[-1, 4]),
                        n_records=int(g[:upto, 2].sum()),
                        n_lines=int(g[:, 3].sum()),
                        err_offset=int(g[first_bad, 1]) if first_bad is not None else 0,
                        tail_offset=None, line_phase=0)
        return total, Stats(self.eng.max_len, stats_words)

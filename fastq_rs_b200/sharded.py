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
  * ONE all_reduce(sum) of [u64 statistics block | world x 8 outcome words]: every rank writes its outcome into
    its own slot, so the reduction is also the gather that confirms the inferred line phases.

With the product engine (`fastq_rs_b200.Engine`) the collective is the library's own: NCCL bound behind the
C ABI (fqb_comm_init / fqb_allreduce / fqb_fetch_reduced), `torch.distributed` only carries the ncclUniqueId to
the other ranks.  The engine object is duck-typed so that the protocol can be tested on CPU with a stand-in
whose send buffer is reduced over gloo.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._lib import E_PHASE, E_RETRY, OK
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
    """Glue around the engine's N-rank entry points.  Per step, in the common case: parse (enqueued), ONE
    all-reduce of [statistics block | world x 8 outcome words] (each rank's outcome sits in its own slot, so the
    sum is also the gather), ONE device-to-host copy.  With the product engine the collective is the library's
    own (fqb_allreduce: NCCL bound behind the C ABI); a duck-typed engine without it (the CPU tests' stand-in over
    gloo) has the same buffer reduced through torch.distributed."""

    def __init__(self, engine, dist=None, group=None, device=None):
        self.eng = engine
        self.dist = dist
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.reparsed = 0   # shards parsed twice so far (inference not confirmed)
        self.collectives = 0
        self.native = hasattr(engine, "allreduce")
        if self.native and engine.comm_world() != self.world:
            self._comm_init()
        elif hasattr(engine, "set_rank"):
            engine.set_rank(self.rank, self.world)

    def _comm_init(self):
        """hand rank 0's ncclUniqueId to the other ranks (torch.distributed is only the messenger)"""
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device=self.device)
        if self.world > 1:
            if self.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(self.eng.comm_unique_id()), dtype=torch.uint8))
            self.dist.broadcast(uid, src=0, group=self.group)
        self.eng.comm_init(self.rank, self.world, bytes(uid.cpu().numpy().tobytes()) if self.world > 1 else None)

    # -- the one collective -----------------------------------------------------------------------
    def _reduce(self, want_stats=True):
        """(global stats words | None, outcomes[world, 8]) of the parse enqueued last"""
        self.collectives += 1
        if self.native:
            self.eng.allreduce()
            return self.eng.fetch_reduced(want_stats=want_stats)
        x = self.eng.device_exchange().clone()
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(x, op=self.dist.ReduceOp.SUM, group=self.group)
        h = x.cpu().numpy()
        nw = h.size - 8 * self.world
        return (h[:nw].view(np.uint64).copy() if want_stats else None), h[nw:].reshape(self.world, 8)

    def _my_slot(self):
        x = self.eng.device_exchange()
        nw = x.numel() - 8 * self.world
        return x[nw + 8 * self.rank: nw + 8 * self.rank + 8]

    def _enqueue(self, sh: ShardSpec, hist, index, line_base, infer):
        """parse_device without waiting for it"""
        view = sh.data[sh.front:] if sh.front else sh.data
        if sh.b == sh.a:   # an empty shard (more ranks than 16-byte pieces): nothing to own, nothing to read
            self.eng.parse_device(view, n_own=0, n_avail=0, hist=hist, index=None, line_base=line_base,
                                  stream_offset=sh.a, line_start=False, front16=False, eof=sh.is_last)
        else:
            self.eng.parse_device(view, n_own=sh.b - sh.a, n_avail=sh.b - sh.a + sh.halo, hist=hist, index=index,
                                  line_base=line_base, stream_offset=sh.a, line_start=(sh.a == 0),
                                  front16=(sh.front == 16), eof=sh.is_last, infer_start=infer)

    def parse(self, sh: ShardSpec, hist: bool = True, index=None):
        """Delimit (+histograms) the whole stream; every rank returns the GLOBAL (Outcome, Stats).
        `index` (optional int32 tensor) receives this rank's line ends (low 32 bits of the stream
        offsets of the '\n' in [a, b))."""
        self.begin(sh, hist, index)
        return self.finish()

    def begin(self, sh: ShardSpec, hist: bool = True, index=None):
        """The asynchronous half of parse(): the shard's parse and the one collective are enqueued on the current
        stream, nothing is waited for.  A caller with two parsers (two engine contexts) can enqueue the next
        stream's shard before it reads this one's result with finish()."""
        infer = sh.a != 0 and sh.b > sh.a
        self._pending = (sh, hist, index, infer)
        self._enqueue(sh, hist, index, 0, infer)
        self.collectives += 1
        if self.native:
            self.eng.allreduce()

    def finish(self):
        """Wait for begin()'s collective; the GLOBAL (Outcome, Stats) on every rank."""
        sh, hist, index, infer = self._pending
        self._pending = None
        if self.native:
            words, g = self.eng.fetch_reduced()
        else:
            self.collectives -= 1
            words, g = self._reduce()
        lb = np.concatenate([[0], np.cumsum(g[:-1, 3])])            # exact line numbers: prefix of the n_lines
        if bool((g[:, 0] == OK).all()) and self._phases_ok(g, lb):
            total = Outcome(status=OK, finished=bool(g[-1, 1]), n_records=int(g[:, 2].sum()),
                            n_lines=int(g[:, 3].sum()), err_offset=0, tail_offset=None, line_phase=0)
            return total, Stats(self.eng.max_len, words)
        return self._careful(sh, hist, index, infer, g, lb)

    def _phases_ok(self, g, lb) -> bool:
        """every inferring shard (all but the first, unless it owns nothing) reports the phase its exact
        line number has; shards that own nothing (n_lines == 0 and n_records == 0) are exempt"""
        for r in range(1, self.world):
            empty = g[r, 3] == 0 and g[r, 2] == 0
            if not empty and (int(g[r, 6]) & 3) != (int(lb[r]) & 3):
                return False
        return True

    def _careful(self, sh, hist, index, infer, g, lb):
        """Some outcome shows an error, an unconfirmed inference or FQB_E_PHASE: confirm or redo the
        inferences, first error in stream order wins.  Every rank holds all outcomes, so every rank takes the
        same decisions without further messages; one more reduction at the end."""
        if bool((g[:, 0] == E_PHASE).any()):
            # some shard could not even deliver its newline count: every rank counts (one cheap pass)
            view = sh.data[sh.front:] if sh.front else sh.data
            n = self.eng.count_lines(view, sh.b - sh.a) if sh.b > sh.a else 0
            self._my_slot()[3] = n          # (my outcome's n_lines word in the send buffer)
            _, g2 = self._reduce(want_stats=False)
            lb = np.concatenate([[0], np.cumsum(g2[:-1, 3])])
        line_base = int(lb[self.rank])
        mine = g[self.rank]
        if infer and (int(mine[0]) == E_PHASE or (int(mine[6]) & 3) != (line_base & 3)):
            self.reparsed += 1      # the inference is not confirmed: parse again with the exact line number
            self._enqueue(sh, hist, index, line_base, False)
        # (a rank that does not parse again still has its block and outcome in the send buffer)
        if self.native:
            # FQB_E_RETRY in my slot (a bad record in the middle of my shard): fetch completes that parse and puts the
            # final block and outcome into the send buffer; a no-op otherwise
            self.eng.fetch(want_stats=False)
        _, g = self._reduce(want_stats=False)
        # first error in stream order wins; shards behind it contribute nothing
        bad = np.nonzero(g[:, 0] != OK)[0]
        first_bad = int(bad[0]) if bad.size else None
        if first_bad is not None and self.rank > first_bad:
            self.eng.device_stats().zero_()
        words, _ = self._reduce()
        upto = self.world if first_bad is None else first_bad + 1
        total = Outcome(status=int(g[first_bad, 0]) if first_bad is not None else OK,
                        finished=first_bad is None and bool(g[-1, 1]),
                        n_records=int(g[:upto, 2].sum()),
                        n_lines=int(g[:, 3].sum()),
                        err_offset=int(g[first_bad, 4]) if first_bad is not None else 0,
                        tail_offset=None, line_phase=0)
        return total, Stats(self.eng.max_len, words)

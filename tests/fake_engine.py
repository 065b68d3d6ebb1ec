"""CPU stand-in for fastq_rs_b200.Engine, for testing the N-rank driver (sharded.py) without a GPU.
TEST INFRASTRUCTURE: it answers from the CPU oracle; it is never used by the product."""
import numpy as np
import torch

from fastq_rs_b200.engine import Outcome
from oracle import oracle

E_PHASE = 7


def stats_words(max_len, st, n_records):
    """oracle Stats -> flat u64 block in the layout of include/fastq_b200.h"""
    P = max_len
    w = np.zeros(8 + P + 2 + 6 * P + 256 * P, dtype=np.uint64)
    w[0], w[1] = n_records, st.n_bases
    w[8:8 + P + 2] = st.len_hist
    w[8 + P + 2:8 + P + 2 + 6 * P] = st.base_hist.reshape(-1)
    w[8 + P + 2 + 6 * P:] = st.qual_hist.reshape(-1)
    return w


class FakeEngine:
    def __init__(self, max_len=150, lie_phase=False, fail_infer=False):
        self.max_len = max_len
        self.lie_phase = lie_phase      # report a wrong inferred phase (the driver must parse again)
        self.fail_infer = fail_infer    # answer E_PHASE to every inference
        self.n_parses = 0
        self._nw = 8 + max_len + 2 + 6 * max_len + 256 * max_len
        self._out = None
        self._res = torch.zeros(8, dtype=torch.int64)
        self.set_rank(0, 1)

    def set_rank(self, rank, world):
        """the send buffer of the one collective: [statistics block | world x 8 outcome words] (fqb_device_exchange)"""
        self.rank, self.world = rank, world
        self._x = torch.zeros(self._nw + 8 * world, dtype=torch.int64)
        self._stats = self._x[:self._nw]

    def device_exchange(self):
        return self._x

    def device_result(self):
        """the 8-word outcome block of fqb_device_result"""
        return self._res

    def _publish(self, o):
        self._out = o
        self._res[:] = torch.tensor([o.status, int(o.finished), o.n_records, o.n_lines, o.err_offset,
                                     -1 if o.tail_offset is None else o.tail_offset, o.line_phase, 0], dtype=torch.int64)
        self._x[self._nw:] = 0
        self._x[self._nw + 8 * self.rank: self._nw + 8 * self.rank + 8] = self._res

    def count_lines(self, view, n):
        return int((view[:n].numpy() == 10).sum())

    def device_stats(self):
        return self._stats

    def fetch(self, want_stats=False):
        return self._out, None

    def parse_device(self, view, n_own, n_avail, hist=True, index=None, line_base=0, stream_offset=0,
                     line_start=True, eof=True, front16=False, infer_start=False, stream=None):
        self.n_parses += 1
        self._stats.zero_()
        buf = view[:n_avail].numpy()
        n_lines = int((buf[:n_own] == 10).sum())
        at_ls = line_start or (front16 and int(view.storage_offset()) >= 1 and
                               int(view._base[view.storage_offset() - 1] if view._base is not None else 0) == 10)
        nl = np.nonzero(buf == 10)[0]

        def start_for(K):   # first record start when K owned newlines lie in front of it
            if K == 0:
                return 0
            return int(nl[K - 1]) + 1 if K <= nl.size else None

        def try_parse(start):
            """records starting in [start, n_own): (status, err_offset, n_rec, end) via the oracle"""
            res, recs = oracle.each(buf[start:].tobytes())
            n, end = 0, start
            for r in recs:
                if start + r.offset >= n_own:
                    break
                n += 1
                end = start + r.offset + len(r.raw)
            status = 0
            err = 0
            if res.status != 0 and (len(recs) == n):   # the bad record is the next one
                nxt = end
                if nxt < n_own and not (not eof and res.status == 5 and False):
                    status, err = res.status, stream_offset + nxt
            return status, err, n, end

        if infer_start:
            if self.fail_infer:
                self._publish(Outcome(E_PHASE, False, 0, 0, 0, None, 0))
                return
            good = []
            for K in ([0, 1, 2, 3] if at_ls else [1, 2, 3, 4]):
                s = start_for(K)
                if s is None:
                    continue
                st, _, n, _ = try_parse(s)
                if st == 0 and n >= 2:
                    good.append(K)
            if len(good) != 1:
                self._publish(Outcome(E_PHASE, False, 0, 0, 0, None, 0))
                return
            K = good[0]
            phase = (4 - K) & 3
            if self.lie_phase:
                phase = (phase + 1) & 3
        else:
            ph = line_base & 3
            K = 0 if (at_ls and ph == 0) else 4 - ph
            phase = ph
        s = start_for(K)
        if s is None:
            s = n_avail
        status, err, n_rec, end = try_parse(s)
        if n_rec:
            _, st = oracle.each_stats(buf[s:end].tobytes(), self.max_len)
            self._stats[:] = torch.from_numpy(stats_words(self.max_len, st, n_rec).view(np.int64))
        self._publish(Outcome(status, status == 0, n_rec, n_lines, err, None, phase))


class FakeHostEngine:
    """Oracle-backed stand-in for the HOST entry points of fastq_rs_b200.Engine (parse_host with
    FQB_F_PARTIAL refills), so that the host mirror of the crate's drivers (fastq_rs_b200/parser.py: each,
    record_sets, parallel_each, each_zipped) can be tested without a GPU.  TEST INFRASTRUCTURE."""

    def __init__(self, max_len=150):
        self.max_len = max_len
        self.n_calls = 0

    def parse_host(self, data, *, hist=True, want_index=False, want_stats=True, partial=False, stream_offset=0):
        self.n_calls += 1
        buf = bytes(np.ascontiguousarray(data).view(np.uint8).tobytes()) if isinstance(data, np.ndarray) else bytes(data)
        # the too-long rule depends on the stream offset mod 16: put a record of that length (mod 16) in front
        lead = 16 + (stream_offset & 15)
        pre = b"@" + b"a" * (lead - 8) + b"\nA\n+\nB\n"
        res, idx = oracle.each_index(pre + buf)
        n = res.n_records - 1
        status, err, tail = res.status, res.err_offset - lead + stream_offset, None
        if partial and status == oracle.E_TRUNCATED:
            status, tail, err = 0, err, 0          # the refill ends inside a record: carried over, not an error
        index = (idx[1:, 1:5].reshape(-1).astype(np.int64) - lead + stream_offset).astype(np.uint64)
        out = Outcome(status, status == 0 and tail is None, n, buf.count(b"\n"), err if status else 0, tail, 0)
        st = None
        if want_stats:
            _, st = oracle.each_stats(buf, self.max_len)
        return out, st, (index if want_index else None)

"""Independent pure-Python model of Parser::each (src/lib.rs:221-303 + src/buffer.rs +
src/records.rs:201-247), written separately from oracle/fastq_oracle.c so that the two
restatements can be fuzzed against each other (tests/test_oracle_golden.py).  Small inputs
only.  TEST INFRASTRUCTURE."""

OK, E_HEADER, E_SEP, E_LENGTH, E_TOO_LONG, E_TRUNCATED = range(6)


def from_buffer(view: bytes):
    """-> ('empty'|'incomplete'|'record'|error-int, idx)"""
    if not view:
        return "empty", None
    if view[0:1] != b"@":
        return E_HEADER, None
    h = view.find(b"\n")
    if h < 0:
        return "incomplete", None
    s = view.find(b"\n", h + 1)
    if s < 0:
        return "incomplete", None
    if s + 1 >= len(view):
        return "incomplete", None
    if view[s + 1:s + 2] != b"+":
        return E_SEP, None
    p = view.find(b"\n", s + 1)
    if p < 0:
        return "incomplete", None
    q = view.find(b"\n", p + 1)
    if q < 0:
        return "incomplete", None
    if q - p != s - h:
        return E_LENGTH, None
    return "record", (h, s, p, q)


def trim(b: bytes) -> bytes:
    return b[:-1] if b.endswith(b"\r") else b


class Buf:
    def __init__(self, cap):
        self.d = bytearray(cap)
        self.cap = cap
        self.start = 0
        self.end = 0

    def view(self):
        return bytes(self.d[self.start:self.end])

    def clean(self):
        if self.start == 0:
            return
        n = self.end - self.start
        ne = (n + 15) & ~15
        ns = ne - n
        if ns >= self.start:
            return
        self.d[ns:ne] = self.d[self.start:self.end]
        self.start, self.end = ns, ne

    def read_into(self, src, pos, max_read):
        free = self.cap - self.end
        want = free if free < 4096 else free - free % 4096
        n = min(want, len(src) - pos)
        if max_read:
            n = min(n, max_read)
        self.d[self.end:self.end + n] = src[pos:pos + n]
        self.end += n
        return n


def each(data: bytes, bufsize=68 * 1024, max_read=0):
    """-> (status, records[(head, seq, qual, raw, offset)])"""
    b = Buf(bufsize)
    pos = 0
    consumed = 0
    out = []
    while True:
        kind, idx = from_buffer(b.view())
        if kind == "record":
            h, s, p, q = idx
            v = b.view()
            out.append((trim(v[1:h]), trim(v[h + 1:s]), trim(v[p + 1:q]), v[:q + 1], consumed))
            b.start += q + 1
            consumed += q + 1
            continue
        if kind == "empty":
            b.clean()
            n = b.read_into(data, pos, max_read)
            pos += n
            if n == 0:
                return OK, out
            continue
        if kind == "incomplete":
            b.clean()
            if b.cap - b.end == 0:
                return E_TOO_LONG, out
            n = b.read_into(data, pos, max_read)
            pos += n
            if n == 0:
                return E_TRUNCATED, out
            continue
        return kind, out


def stats(records, P):
    import numpy as np
    base = np.zeros((P, 6), dtype=np.uint64)
    qual = np.zeros((P, 256), dtype=np.uint64)
    lens = np.zeros(P + 2, dtype=np.uint64)
    cls = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3, ord("N"): 4}
    nb = cs = cq = 0
    for (_h, s, q, _raw, _off) in records:
        nb += len(s)
        lens[min(len(s), P + 1)] += 1
        for i, c in enumerate(s):
            if i < P:
                base[i, cls.get(c, 5)] += 1
            else:
                cs += 1
        for i, c in enumerate(q):
            if i < P:
                qual[i, c] += 1
            else:
                cq += 1
    return dict(n_records=len(records), n_bases=nb, clip_seq=cs, clip_qual=cq, base=base,
                qual=qual, lens=lens)

#!/usr/bin/env python
"""The two GPU-side closures of this package, from the command line:

    python examples/fastq_stats_filter.py stats  FILE [MAX_LEN]    # per-position mean quality + base composition
    python examples/fastq_stats_filter.py filter FILE OUT [dna|dnan]  # keep records whose seq() is pure ACGT(N)

`stats` streams the file through the pinned ring (nothing but the counters comes back); `filter` runs
Record::validate_dna / validate_dnan (src/records.rs:19-33) and the compaction of the survivors on the GPU
and writes them verbatim (Record::write).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastq_rs_b200 as fq  # noqa: E402


def main(argv):
    if len(argv) >= 2 and argv[0] == "stats":
        max_len = int(argv[2]) if len(argv) > 2 else 150
        outcome, st = fq.parse_path(argv[1], lambda parser: parser.stats(), max_len=max_len)
        print(f"records {st.n_records}  bases {st.n_bases}  status {outcome.status}")
        q = np.arange(256, dtype=np.float64) - 33.0
        for pos in range(max_len):
            n = int(st.qual_hist[pos].sum())
            if n == 0:
                break
            a, c, g, t, nn, other = (int(x) for x in st.base_hist[pos])
            print(f"{pos + 1:4d}  meanQ {float((st.qual_hist[pos] * q).sum()) / n:6.2f}  A {a} C {c} G {g} T {t} N {nn} other {other}")
        outcome.raise_for_status()
    elif len(argv) >= 3 and argv[0] == "filter":
        keep = argv[3] if len(argv) > 3 else "dnan"
        with open(argv[2], "wb") as w:
            n = fq.parse_path(argv[1], lambda parser: parser.filter_to(w, keep))
        print(n)
    else:
        sys.exit(__doc__)


if __name__ == "__main__":
    main(sys.argv[1:])

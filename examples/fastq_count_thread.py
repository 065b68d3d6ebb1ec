#!/usr/bin/env python
"""examples/fastq-count-thread.rs of the reference, on the B200 path: parallel_each with worker
threads that count the records of the RecordSets dealt to them (the GPU delimits, the batches are
dealt round-robin over bounded queues as in src/lib.rs:521-549).

    python examples/fastq_count_thread.py [FILE|-] [N_THREADS]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastq_rs_b200 as fq  # noqa: E402


def main(argv):
    path = None if not argv or argv[0] == "-" else argv[0]
    n_threads = int(argv[1]) if len(argv) > 1 else 1
    results = fq.parse_path(path, lambda parser: parser.parallel_each(
        n_threads, lambda record_sets: sum(s.len() for s in record_sets)))
    print(sum(results))


if __name__ == "__main__":
    main(sys.argv[1:])

#!/usr/bin/env python
"""examples/multiple-files.rs of the reference, on the B200 path: two FASTQ files (e.g. R1 / R2 of a
paired-end run) walked in lock step with each_zipped; both are delimited on the GPU refill by refill.

    python examples/multiple_files.py READS_1.fastq READS_2.fastq
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastq_rs_b200 as fq  # noqa: E402


def main(argv):
    if len(argv) != 2:
        sys.exit("Need two input files.")
    counts = [0, 0]

    def pair(rec1, rec2):
        counts[0] += rec1 is not None
        counts[1] += rec2 is not None
        return True, True

    fq.parse_path(argv[0], lambda p1: fq.parse_path(argv[1], lambda p2: fq.each_zipped(p1, p2, pair)))
    print(f"Number of reads: ({counts[0]}, {counts[1]})")


if __name__ == "__main__":
    main(sys.argv[1:])

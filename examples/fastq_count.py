#!/usr/bin/env python
"""examples/fastq-count.rs of the reference, on the B200 path:

    python examples/fastq_count.py [FILE|-]            # number of records
    python examples/fastq_count.py --each [FILE|-]     # the same through Parser.each and a closure

parse_path opens the file (or stdin), Parser.count() streams it through the pinned ring and the
delimiting kernels; with --each the line-end index comes back and the closure runs here, once per
record, exactly like `parser.each(|_| { total += 1; true })`.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastq_rs_b200 as fq  # noqa: E402


def main(argv):
    use_each = "--each" in argv
    args = [a for a in argv if a != "--each"]
    path = None if not args or args[0] == "-" else args[0]
    if use_each:
        total = 0

        def run(parser):
            nonlocal total

            def on_record(_rec):
                nonlocal total
                total += 1
                return True
            parser.each(on_record)            # raises FastqError("...") on an invalid file
        fq.parse_path(path, run)
    else:
        total = fq.parse_path(path, lambda parser: parser.count())
    print(total)


if __name__ == "__main__":
    main(sys.argv[1:])

"""fastq_rs_b200 -- B200-native FASTQ record delimiting + per-position statistics.

Host-side mirror of the `fastq` crate's Parser / Record / RefRecord / each / parallel_each
surface (aseyboldt/fastq-rs 0.6.0) over hand-written sm_100a kernels behind a C ABI
(include/fastq_b200.h).  See DESIGN.md.
"""
from ._lib import (E_HEADER, E_LENGTH, E_SEP, E_TOO_LONG, E_TRUNCATED, OK, SO_PATH, build)  # noqa: F401
from .engine import Engine, FastqError, Outcome, Stats, expand_index  # noqa: F401
from .parser import (BUFSIZE, OwnedRecord, Parser, RecordRefIter, RecordSet, RefRecord,  # noqa: F401
                     default_engine, each_zipped, parse_path)

"""CPU checks of the drop-in boundary: the built library loads, exports every symbol that
include/fastq_b200.h declares, the ctypes mirrors match the header's struct layouts, and the
product path refuses to run without a CUDA device (no CPU fallback, no oracle import)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fastq_b200.h")


@pytest.fixture(scope="module")
def so_path():
    from fastq_rs_b200 import _lib
    return _lib.build()


def header_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fqb_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from fastq_rs_b200 import _lib
    assert header_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(so_path):
    L = C.CDLL(so_path)
    for name in header_functions():
        assert hasattr(L, name), name
    L.fqb_abi_version.restype = C.c_uint32
    from fastq_rs_b200 import _lib
    assert L.fqb_abi_version() == _lib.ABI_VERSION == 3


def test_no_torch_types_in_signatures():
    txt = open(HEADER).read()
    assert "torch" not in txt and "at::" not in txt and "std::" not in txt


def test_struct_layouts_match_header(so_path, tmp_path):
    """Compile a C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from fastq_rs_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "fastq_b200.h"
int main(void) {
    printf("%zu %zu %zu\n", sizeof(fqb_config), sizeof(fqb_shard), sizeof(fqb_result));
    printf("%zu %zu %zu %zu\n", offsetof(fqb_shard, line_base), offsetof(fqb_shard, flags),
           offsetof(fqb_shard, d_index), offsetof(fqb_shard, index_cap));
    printf("%zu %zu\n", offsetof(fqb_result, n_records), offsetof(fqb_result, tail_offset));
    printf("%zu %zu\n", offsetof(fqb_config, slot_bytes), offsetof(fqb_config, n_slots));
    return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = [int(x) for x in out]
    S, R, Cf = _lib.Shard, _lib.Result, _lib.Config
    exp = [C.sizeof(Cf), C.sizeof(S), C.sizeof(R), S.line_base.offset, S.flags.offset, S.d_index.offset,
           S.index_cap.offset, R.n_records.offset, R.tail_offset.offset, Cf.slot_bytes.offset, Cf.n_slots.offset]
    assert got == exp


def test_stats_layout_and_messages(so_path):
    from fastq_rs_b200 import _lib
    L = _lib.lib()
    for P in (1, 150, 300, 4096):
        assert L.fqb_stats_len_hist_off(P) == 8
        assert L.fqb_stats_base_hist_off(P) == 8 + P + 2
        assert L.fqb_stats_qual_hist_off(P) == 8 + P + 2 + 6 * P
        assert L.fqb_stats_words(P) == 8 + P + 2 + 6 * P + 256 * P
    # the reference's exact texts (src/records.rs:145,159,236; src/lib.rs:281,289)
    assert L.fqb_strerror(1) == b"Fastq headers must start with '@'"
    assert L.fqb_strerror(2) == b"Sequence and quality not separated by +"
    assert L.fqb_strerror(3) == b"Sequence and quality length mismatch"
    assert L.fqb_strerror(4) == b"Fastq record is too long"
    assert L.fqb_strerror(5) == b"Possibly truncated input file"


def test_bad_arguments_are_rejected_without_a_device(so_path):
    from fastq_rs_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.fqb_create(None, C.byref(h)) == _lib.E_ARG
    cfg = _lib.Config(99, 0, 150, 0, 0, 0, 0)            # wrong ABI version
    assert L.fqb_create(C.byref(cfg), C.byref(h)) == _lib.E_ARG
    cfg = _lib.Config(_lib.ABI_VERSION, 0, 0, 0, 0, 0, 0)  # max_len = 0
    assert L.fqb_create(C.byref(cfg), C.byref(h)) == _lib.E_ARG


def test_product_fails_loudly_without_cuda(so_path):
    """No CPU fallback: without a device the engine refuses to come up."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import fastq_rs_b200 as fq
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fq.Engine(max_len=150)
    with pytest.raises(RuntimeError):
        fq.Parser(b"@a\nA\n+\nI\n").count()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fastq_rs_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fastq_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "fastq_oracle", "oracle/", "oracle."):
                    assert needle not in txt, (os.path.join(dirpath, f), needle)
    code = ("import sys; sys.path.insert(0, %r); import fastq_rs_b200; "
            "assert not [m for m in sys.modules if m.startswith('oracle')]" % ROOT)
    subprocess.check_call([sys.executable, "-c", code])

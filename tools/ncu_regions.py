"""Instruction / stall-sample breakdown of the stream kernel by source region (ncu report with
--import-source on):  python tools/ncu_regions.py gpurun_out/scan_full.ncu-rep [n_records] [--lines]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
nrec = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 13379960
show_lines = "--lines" in sys.argv
def regions(path, pats):
    """region boundaries = lines matching the given (regex, name) patterns, in file order"""
    out = []
    for i, l in enumerate(open(path), 1):
        for pat, name in pats:
            if re.search(pat, l):
                out.append((i, name))
    return out
import os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fastq_rs_b200", "csrc")
spec = {
    "fq_stream.cu": regions(os.path.join(root, "fq_stream.cu"), [
        (r"uint32_t nlbits3\(", "nlbits"), (r"Window win_load\(", "win_load"), (r"void scan_rank_store\(", "win_scan"),
        (r"uint32_t infer_start\(", "infer"), (r"^struct WinAcc", "srounds (8-lane scanned rounds)"),
        (r"^struct FRounds", "frounds (predicted rounds)"), (r"bool no_newline32\(", "pred checks"),
        (r"uint32_t stream_pass\(", "stream_pass"), (r"^struct StepK", "line_steps (var)"),
        (r"^struct RecSink", "validate_block (var)"), (r"^struct Shape", "pred_pass / flex_pass"),
        (r"uint32_t win_count_newlines\(", "win_count_newlines"), (r"bool desc_put\(", "desc / drain"),
        (r"^struct RangeState", "var_loop"), (r"void __launch_bounds__\(C::NTHREADS, 1\) fq_stream_kernel", "prologue"),
        (r"// ---- where the first record", "first_window"), (r"// ---- stream through the range", "loop_head"),
        (r"// ---- scanned window ----", "loop_scanned"), (r"// line ends of the consumed records that lie in the owned bytes of the shard\n                n_lines", "loop_tail"),
        (r"// ---- drain", "drain")]),
    "fq_hist.cuh": regions(os.path.join(root, "fq_hist.cuh"), [
        (r"void named_bar\(", "hist:asm_helpers"), (r"void trace_ev\(", "hist:trace"), (r"uint32_t nlmask16s7\(", "hist:nlmask"),
        (r"void flush_hist\(", "hist:flush"), (r"void account_record\(", "hist:account"), (r"record_global\(", "hist:record_global"),
        (r"^struct LaneConst", "hist:rounds"), (r"uint32_t small_prefix\(", "hist:small_prefix")]),
}
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; tot = {}; samp = {}; ti = ts = 0; lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try: inst = float(r[ie]); sm = float(r[ns])
        except ValueError: continue
        ti += inst; ts += sm
        fn = (cur_file or "?").split("/")[-1]
        key = "other:" + fn
        if fn in spec:
            ln = int(r[0]); key = "pre:" + fn
            for lo, name in spec[fn]:
                if ln >= lo: key = name
        tot[key] = tot.get(key, 0) + inst; samp[key] = samp.get(key, 0) + sm
        lines.append((inst, fn, int(r[0]), r[1].strip()[:105]))
print(f"total {ti:.3e} warp-instr = {ti / nrec:.1f}/record, {ts:.0f} samples")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k] / ti * 100:6.2f}%i {tot[k] / nrec:7.1f}/rec {samp[k] / ts * 100:6.2f}%s  {k}")
if show_lines:
    for inst, fn, ln, src in sorted(lines, reverse=True)[:60]:
        print(f"{inst / nrec:6.2f}/rec {fn[:12]:12s} L{ln}: {src}")

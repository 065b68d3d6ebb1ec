"""Per-source-line instruction and stall summary from an ncu report:
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if len(r) > 8 and r[0] == "Line No")
ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():
        try:
            inst = float(r[ie]); samp = float(r[ns])
        except ValueError:
            continue
        st = sorted(((float(r[i]) if r[i] not in ("", "-") else 0.0, h) for i, h in stall_cols), reverse=True)[:2]
        lines.append((inst, samp, r[0], r[1].strip()[:95], st))
ti = sum(l[0] for l in lines); ts = sum(l[1] for l in lines)
print(f"total warp-instructions {ti:.3e}, samples {ts:.0f}")
print("by instructions:")
for inst, samp, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{inst/ti*100:5.1f}%i {samp/ts*100:5.1f}%s L{ln:>4}: {src}  [{', '.join(f'{h[6:]}={v:.0f}' for v,h in st if v)}]")
print("by stall samples:")
for inst, samp, ln, src, st in sorted(lines, key=lambda l: -l[1])[:20]:
    print(f"{inst/ti*100:5.1f}%i {samp/ts*100:5.1f}%s L{ln:>4}: {src}  [{', '.join(f'{h[6:]}={v:.0f}' for v,h in st if v)}]")

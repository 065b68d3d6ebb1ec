"""Aggregate ncu per-line instruction counts over source line ranges: python tools/ncu_ranges.py rep file.cu a-b:name ..."""
import csv, subprocess, sys, io
rep = sys.argv[1]; fname = sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    r, name = a.split(":"); lo, hi = r.split("-"); ranges.append((int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; tot = {}; samp = {}; ti = ts = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try: inst = float(r[ie]); sm = float(r[ns])
        except ValueError: continue
        ti += inst; ts += sm
        key = "other:" + (cur_file or "?").split("/")[-1]
        if cur_file and cur_file.endswith(fname):
            ln = int(r[0])
            for lo, hi, name in ranges:
                if lo <= ln <= hi: key = name; break
            else: key = "unranged:" + fname
        tot[key] = tot.get(key, 0) + inst; samp[key] = samp.get(key, 0) + sm
print(f"total {ti:.3e} warp-instr, {ts:.0f} samples")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k]/ti*100:6.2f}%i {samp[k]/ts*100:6.2f}%s  {k}")

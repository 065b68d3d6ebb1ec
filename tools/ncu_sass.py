"""SASS of a kernel in an ncu report with per-record execution counts:
   python tools/ncu_sass.py rep n_records [min_per_record]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nrec = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if len(r) > 8 and r[0] == "Address")
ie, ns, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot = 0
for r in rows:
    if len(r) == len(hdr) and r[0].startswith("0x"):
        try: c = float(r[ie]); sm = float(r[ns])
        except ValueError: continue
        tot += c
        if c / nrec >= thr:
            print(f"{c/nrec:6.2f} {sm:6.0f} {r[0][-5:]} {r[isrc].strip()}")
print("total per record", tot / nrec)

"""Warp-state samples of an .ncu-rep aggregated by CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py REPORT.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
items, tot, fpath, hdr = {}, 0, "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":      # only the per-line aggregate rows
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    if n == 0:
        continue
    key = (fpath, int(r[0]))
    e = items.setdefault(key, [0, r[1].strip()[:90], {}, 0])
    e[0] += n
    e[3] += int(r[ix["Instructions Executed"]] or 0)
    for h, i in stall_cols:
        if r[i] not in ("", "0"):
            e[2][h[6:]] = e[2].get(h[6:], 0) + int(r[i])
    tot += n
print("total samples", tot)
for (f, ln), (n, src, st, ex) in sorted(items.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100 * n / tot:5.1f}% {f}:{ln:<4d} {src:90s} {dict(sorted(st.items(), key=lambda x: -x[1])[:3])}")

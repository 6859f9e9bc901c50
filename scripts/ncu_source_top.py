"""top source lines by warp-stall samples of an ncu report (--import-source on).  usage: python scripts/ncu_source_top.py rep [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0; lines = []
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        f = r[1].split('/')[-1]; hdr = rows[i + 2]; ci = {}
        for k, n in enumerate(hdr): ci.setdefault(n, k)
        j = i + 3
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            rr = rows[j]
            if len(rr) >= len(hdr) and rr[0].isdigit():
                try: lines.append((int(rr[ci["# Samples"]] or 0), f, int(rr[0]), rr[1][:130]))
                except ValueError: pass
            j += 1
        i = j
    else: i += 1
tot = sum(l[0] for l in lines); print("samples", tot)
for s, f, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]: print(f"{s:6d} {100 * s / tot:5.1f}% {f}:{n}  {src}")

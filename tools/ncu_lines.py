"""Aggregate an ncu report's stall samples / executed instructions per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr and len(r) > iinst and r[0] != "":
        try:
            data.append((int(r[isamp]), int(r[iinst]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print("total samples", ts, "total warp-instructions", ti)
for d in sorted(data, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d | %s" % (100.0 * d[0] / ts, 100.0 * d[1] / ti, d[2], d[3], d[4]))

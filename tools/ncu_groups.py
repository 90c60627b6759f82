"""Group an ncu report's executed warp-instructions / stall samples by source-line ranges."""
import csv
import subprocess
import sys

rep = sys.argv[1]
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
            data.append((fname, int(r[0]), int(r[isamp]), int(r[iinst])))
        except ValueError:
            pass
ts, ti = sum(d[2] for d in data), sum(d[3] for d in data)
print("total samples", ts, "warp-instr", ti)
files = sorted(set(d[0] for d in data))
for f in files:
    lines = sorted([d for d in data if d[0] == f], key=lambda d: d[1])
    print("==", f)
    # print in buckets of contiguous lines (gap > 6 starts a new bucket)
    bucket = []
    def flush():
        if bucket:
            s = sum(d[2] for d in bucket); i = sum(d[3] for d in bucket)
            if 100.0 * i / ti >= 0.4 or 100.0 * s / ts >= 0.4:
                print("  lines %4d-%4d : %5.1f%% samples %5.1f%% instr" % (bucket[0][1], bucket[-1][1], 100.0 * s / ts, 100.0 * i / ti))
    for d in lines:
        if bucket and d[1] - bucket[-1][1] > 6:
            flush(); bucket = []
        bucket.append(d)
    flush()

"""source-level view of one kernel of an `ncu --set full --import-source on` report: warp-sampling totals per window of
SASS instructions (stall reasons + dominant opcodes), to see which loops the samples sit in.
  python tools/ncu_source_windows.py report.ncu-rep [window] > profiles/<name>.txt"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
w = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1] if len(rows[0]) > 1 else rows[0])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, SRC = ix["# Samples"], ix["Source"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S]) for r in data)
print(f"instructions {len(data)}, warp samples {tot}; windows of {w} SASS instructions with >= 0.5 % of the samples")
print("first instr | samples | share | top stall reasons | dominant opcodes")
tc = collections.Counter()
for i in range(0, len(data), w):
    seg = data[i:i + w]
    s = sum(int(r[S]) for r in seg)
    c = collections.Counter()
    for r in seg:
        for st in stalls:
            c[st[6:]] += int(r[ix[st]])
            tc[st[6:]] += int(r[ix[st]])
    if s < 0.005 * tot:
        continue
    ops = collections.Counter((r[SRC].split()[1] if r[SRC].strip().startswith("@") else r[SRC].split()[0]) for r in seg if r[SRC].strip())
    print(f"{i:6d} | {s:6d} | {100 * s / tot:5.1f} % | " + " ".join(f"{k}:{v}" for k, v in c.most_common(4)) + " | " +
          " ".join(f"{k}x{v}" for k, v in ops.most_common(5)))
print("whole kernel: " + " ".join(f"{k}:{100 * v / max(1, sum(tc.values())):.1f}%" for k, v in tc.most_common(10)))

"""Stall samples of one kernel per CUDA source line, from an .ncu-rep captured with --import-source on (read here, no GPU).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"][0]
h = rows[hi]
ws = h.index("Warp Stall Sampling (All Samples)")
agg, src, tot = collections.Counter(), {}, 0
for r in rows[hi + 1:]:
    if len(r) <= ws or not r[0].isdigit() or r[2] != "-":  # per-line rows only (SASS rows carry an address)
        continue
    try:
        s = int(r[ws])
    except ValueError:
        continue
    agg[r[0]] += s
    tot += s
    src.setdefault(r[0], r[1])
print("total samples %d" % tot)
for line, s in agg.most_common(top):
    print("%6d %5.1f%%  L%s: %s" % (s, 100.0 * s / max(tot, 1), line, src[line].strip()[:130]))

"""Per-source-line stall samples from `ncu -i rep --page source --print-source cuda,sass --csv` output (exported on the GPU box).
  python tools/ncu_source_csv.py file.csv [top_n] [--no-barrier] [--sass LINE]"""
import collections
import csv
import sys

path = sys.argv[1]
args = [a for a in sys.argv[2:] if not a.startswith("--")]
top = int(args[0]) if args else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No"][0]
h = rows[hi]
ws = h.index("Warp Stall Sampling (All Samples)")
ie = h.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_")]
want_sass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else None
agg, src, tot, reasons, inst = collections.Counter(), {}, 0, collections.defaultdict(collections.Counter), collections.Counter()
cur_line = None
for r in rows[hi + 1:]:
    if len(r) <= ws:
        continue
    if r[0].isdigit() and r[2] == "-":
        cur_line = int(r[0])
        try:
            s = int(r[ws])
        except ValueError:
            continue
        bar = 0
        for i, n in stall_cols:
            try:
                v = int(r[i])
            except ValueError:
                v = 0
            reasons[cur_line][n[6:]] += v
            if n == "stall_barrier":
                bar = v
        if "--no-barrier" in sys.argv:
            s -= bar
        agg[cur_line] += s
        tot += s
        src.setdefault(cur_line, r[1])
        try:
            inst[cur_line] += int(r[ie])
        except ValueError:
            pass
    elif want_sass is not None and cur_line == want_sass and r[2] != "-":
        print("   ", r[2], r[3][:90], "samples", r[ws], "exec", r[ie])
print("total samples %d" % tot)
for line, s in agg.most_common(top):
    rs = ", ".join("%s %d" % (k, v) for k, v in reasons[line].most_common(3) if v)
    print("%6d %5.1f%%  inst %8d  L%d: %s   [%s]" % (s, 100.0 * s / max(tot, 1), inst[line], line, src[line].strip()[:100], rs))

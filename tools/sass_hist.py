"""Opcode histogram of one kernel of a built library, whole function and innermost loop (no GPU needed).
  python tools/sass_hist.py <lib.so|.o> <substring of the mangled kernel name> [--dump file]"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, "\n  instructions:", len(ins))
    if dump:
        open(dump, "w").write("\n".join("%05x  %s" % i for i in ins))
    # innermost loop = the backward branch with the largest span
    best = None
    for a, t in ins:
        m = re.search(r"BRA\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    body = [t for a, t in ins if best and best[0] <= a <= best[1]] if best else [t for _, t in ins]

    def op(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0].split(".")[0]
    c = Counter(op(t) for t in body)
    print("  loop %s: %d instructions" % (("%x..%x" % best) if best else "(none)", len(body)))
    print("  " + ", ".join("%s %d" % kv for kv in c.most_common(24)))
    dfma = [t for t in body if op(t) == "DFMA"]
    if dfma:
        def nsrc(t):
            regs = re.findall(r"-?\|?(U?R\d+|RZ|c\[|[0-9.e+-]+)", t.split(None, 1)[1])
            srcs = [r for r in regs[1:] if re.match(r"R\d+", r)]
            return len(set(srcs))
        three = [t for t in dfma if nsrc(t) == 3]
        print("  DFMA %d: three distinct register sources %d (of those with .reuse %d)" % (len(dfma), len(three), sum(".reuse" in t for t in three)))

#!/bin/bash
# round 2, last check of the committed state: smoke, the whole GPU suite, a default bench run
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u_bench.json"))
print(d["value"], d["steps"], d["warmup"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print(json.dumps(d["config3_classifier"])); print(json.dumps(d["config5_classifier"]["many_sequences"]))
PY

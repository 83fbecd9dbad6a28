#!/bin/bash
# round 2, GPU call P: full GPU suite + bench on the sixteen-warp linking kernel with the enumerate / link overlap
set -x
cd "$GRAFT_REPO_ROOT"
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench.json"))
print(json.dumps(d["config3_classifier"])); print(json.dumps(d["config5_classifier"]["one_sequence"])); print(json.dumps(d["config5_classifier"]["many_sequences"]))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
PY
tail -3 gpurun_out/r2p_bench.err

#!/bin/bash
# one visit: the whole GPU suite, then the bench line (GPU arm only, every configuration)
TAG=${1:-q}; OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
timeout 240 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("cfg2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
for k,v in d["configs"].items():
    print(k, v.get("chain_steps_per_s", v.get("samples_per_s_total")), v.get("ms", v.get("eval_ms")), (v.get("roofline") or {}).get("frac"), ((v.get("roofline") or {}).get("fp64_issue") or {}).get("frac"), v.get("launch"))
PY

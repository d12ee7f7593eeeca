#!/bin/bash
# What the driver runs at round end, on one GPU: the GPU suite, smoke(), bench.py (both arms) with wall-clock times.
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
( time timeout 300 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1 ) 2> $OUT/time_pytest_$TAG.txt; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu_$TAG.log; grep real $OUT/time_pytest_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err ) 2> $OUT/time_bench_ref_$TAG.txt; echo "ref rc=$?"; grep real $OUT/time_bench_ref_$TAG.txt
( time timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err ) 2> $OUT/time_bench_$TAG.txt; echo "bench rc=$?"; grep real $OUT/time_bench_$TAG.txt
python - <<PY
import json
for f in ("$OUT/bench_ref_$TAG.json", "$OUT/bench_$TAG.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d.get("impl","gpu"), d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline",{}).get("value"), d.get("clocks"), d.get("gpu_launches"))
PY

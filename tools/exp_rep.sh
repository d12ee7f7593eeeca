#!/bin/bash
# A/B of the replicated-table evaluation kernel (cfg 5): parity tests, then bench cfg 5 with plain / replicated tables.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_propagation.py -x -q > $OUT/pytest_rep.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_rep.log
for v in 0 1; do
  PBU_REP_TABLES=$v python bench.py --no-cpu --steps 3 --configs 5 2>$OUT/bench_rep_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg5']; print('rep=$v', c['eval_ms'], c['statistics_ms'], c['samples_per_s_eval'], c['roofline']['fp64_issue'])"
done

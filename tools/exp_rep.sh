#!/bin/bash
# A/B of the replicated-table kernels (cfg 3 step kernel, cfg 5 evaluation): parity tests, then bench with plain / replicated tables.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_propagation.py tests/test_gpu_full_size.py -x -q > $OUT/pytest_rep.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_rep.log
for v in 0 1; do
  PBU_REP_TABLES=$v python bench.py --no-cpu --steps 3 --configs 3 2>$OUT/bench_rep_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg3']; print('rep=$v', c['ms'], c['chain_steps_per_s'], c['launch'], c['roofline']['fp64_issue']['frac'])"
done
python tools/experiments.py rep > $OUT/exp_rep.log 2>&1; cat $OUT/exp_rep.log

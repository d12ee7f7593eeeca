#!/bin/bash
# cfg 2: start offset between the warps of a scheduler (PBU_STAGGER, cycles; empty = the planner's value)
for v in "" 0 1500 2800 4000 5600 8000; do
  if [ -n "$v" ]; then export PBU_STAGGER=$v; else unset PBU_STAGGER; fi
  timeout 120 python bench.py --no-cpu --no-configs --steps 8 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stagger [$v]', d['ms_per_step'], d['roofline']['frac'])"
done

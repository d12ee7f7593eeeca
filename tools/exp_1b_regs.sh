#!/bin/bash
for lib in "" r96 r80; do
  if [ -n "$lib" ]; then export PBU_LIB=$PWD/pybitup_b200/_lib/variants/lib$lib.so; else unset PBU_LIB; fi
  python bench.py --no-cpu --steps 3 --configs 1b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg1b']; print('lib [$lib]', c['chain_steps_per_s'], c['roofline']['frac'], c['launch'])"
done

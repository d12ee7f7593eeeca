#!/bin/bash
# cfg 4 (Ito-SDE d=10 N=1e4, streamed records): the compiled layouts, forced one by one
for v in "" "1,1,0,0,0,0" "1,2,0,0,0,0" "1,2,1,0,0,0" "4,1,0,0,0,0" "8,1,0,0,0,0"; do
  PBU_FORCE_VARIANT=$v python bench.py --no-cpu --steps 3 --configs 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg4']; print('variant [$v]', c['chain_steps_per_s'], c['ms'], c['roofline']['frac'], c['launch'])"
done

for v in "" "1,1,0,0" "1,2,0,0" "2,2,0,0" "2,2,1,0" "4,4,0,1" "4,4,1,1" "4,2,0,0"; do
  PBU_FORCE_VARIANT=$v python bench.py --no-cpu --steps 3 --configs 1b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg1b']; print('variant [$v]', c['chain_steps_per_s'], c['roofline']['frac'], c['launch'])"
done

#!/bin/bash
# N-GPU visit (gpurun --gpus N -- bash tools/gpu_multi.sh N tag): sharded == unsharded over NCCL (tests/test_gpu_multi.py,
# N <= 4), then bench.py under torchrun at N ranks (GPU arm only; every configuration, the moment all-reduce inside e2e).
N=${1:-2}; TAG=${2:-multi}; OUT=gpurun_out; mkdir -p $OUT
if [ "$N" -le 4 ]; then
  python -m pytest tests/test_gpu_multi.py -x -q > $OUT/pytest_multi_${N}gpu_$TAG.log 2>&1; echo "pytest multi rc=$?"; tail -2 $OUT/pytest_multi_${N}gpu_$TAG.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --no-cpu \
  > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
print("cfg2", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["e2e"].get("shard_equals_whole"), d["e2e"].get("pooled_statistics_ms"))
for k,v in d["configs"].items():
    print(k, v.get("chain_steps_per_s", v.get("samples_per_s_total")), v.get("ms", v.get("eval_ms")), v.get("scaling"))
PY

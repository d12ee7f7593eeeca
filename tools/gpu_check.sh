#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), exploratory timings, ncu launch list + full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
python bench.py --steps 3 --warmup 3 --no-cpu --iters 100 > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mh_step|mh_coop|eval_|broadcast|fill|elementwise|Fill' -c 60 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --iters 100 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_one.py cubic 100 3 > $OUT/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'mh_step|mh_coop' -s 1 -c 1 -f -o $OUT/prof_cubic_$TAG \
    python tools/profile_one.py cubic 100 3 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python tools/profile_propagation.py > $OUT/prop_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file $OUT/launches_prop_$TAG.csv \
    python tools/profile_propagation.py > $OUT/prop_ncu_$TAG.log 2>&1; echo "ncu propagation launches rc=$?"

#!/bin/bash
# ncu evidence for the round: launch list of a bench run + ONE full capture per hot kernel (each after its plain run exited 0).
# Usage (under gpurun, from the repo root): bash tools/gpu_profile.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
# the .ncu-rep files (tens of MB each with the source page) stay on the box: what comes back is the raw-page summary
# (tools/ncu_summary.py) and the hottest SASS lines of the source page (tools/ncu_hot.py)
summarise() {
  python tools/ncu_summary.py $OUT/$1.ncu-rep > $OUT/$1.json 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1.src.csv 2>/dev/null && python tools/ncu_hot.py $OUT/$1.src.csv 60 > $OUT/$1.hot.txt 2>/dev/null
  rm -f $OUT/$1.ncu-rep $OUT/$1.src.csv
}
python bench.py --steps 3 --warmup 3 --no-cpu --iters 100 --no-configs > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mh_step|mh_coop|eval_|broadcast|fill|aos_to_soa|chain_sums|pooled|elementwise|Fill|vectorized' -c 120 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --iters 100 --no-configs > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
for cfg in cubic arrhenius linear isde; do
  python tools/profile_one.py $cfg 20 3 > $OUT/plain_${cfg}_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'mh_step|mh_coop|isde_step' -s 1 -c 1 -f -o $OUT/prof_${cfg}_$TAG \
      python tools/profile_one.py $cfg 20 3 > $OUT/ncu_${cfg}_$TAG.log 2>&1; echo "ncu $cfg rc=$?"
  summarise prof_${cfg}_$TAG
done
python tools/profile_propagation.py 3125000 > $OUT/plain_prop_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'eval_model' -c 1 -f -o $OUT/prof_evalmodel_$TAG \
    python tools/profile_propagation.py 3125000 > $OUT/ncu_prop_$TAG.log 2>&1; echo "ncu eval_model rc=$?"
summarise prof_evalmodel_$TAG
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file $OUT/launches_prop_$TAG.csv \
    python tools/profile_propagation.py > $OUT/prop_ncu_$TAG.log 2>&1; echo "ncu propagation launches rc=$?"

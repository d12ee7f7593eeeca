#!/bin/bash
# per-SASS-line executed-instruction counts of one kernel (ncu source page), brought back as CSV (gzip):
#   bash tools/exp_src.sh <cfg of tools/profile_one.py> <steps> <tag>
CFG=${1:-linear}; STEPS=${2:-100}; TAG=${3:-src}
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/profile_one.py $CFG $STEPS 3 > $OUT/plain_${CFG}_$TAG.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mh_step|mh_coop|isde_step' -s 1 -c 1 -f -o $OUT/prof_${CFG}_$TAG \
    python tools/profile_one.py $CFG $STEPS 3 > $OUT/ncu_${CFG}_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/prof_${CFG}_$TAG.ncu-rep > $OUT/prof_${CFG}_$TAG.json 2>/dev/null
ncu -i $OUT/prof_${CFG}_$TAG.ncu-rep --page source --csv > $OUT/prof_${CFG}_$TAG.src.csv 2>/dev/null
gzip -f $OUT/prof_${CFG}_$TAG.src.csv; rm -f $OUT/prof_${CFG}_$TAG.ncu-rep; ls -la $OUT/prof_${CFG}_$TAG*

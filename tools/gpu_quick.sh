#!/bin/bash
# quick GPU visit: parity tests + exploratory timings.  Usage: bash tools/gpu_quick.sh <tag> [probe args]
TAG=${1:-q}; shift
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu_$TAG.log | cut -c1-300
timeout 900 python tools/perf_probe.py "$@" > $OUT/probe_$TAG.log 2>&1; echo "probe rc=$?"; cut -c1-420 $OUT/probe_$TAG.log | tail -40

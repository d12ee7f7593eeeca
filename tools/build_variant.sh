#!/bin/bash
# Experimental build of the library with extra compile flags for some translation units, for A/B runs on the GPU box:
#   tools/build_variant.sh <name> "<extra nvcc flags>" <file.cu> [<file.cu> ...]
# compiles the listed files of pybitup_b200/csrc with the flags into csrc/build_<name>/, links them with the objects of
# the main build (make first) into pybitup_b200/_lib/variants/lib<name>.so; select it with PBU_LIB=<path> (pybitup_b200/_capi.py).
set -e
NAME=$1; FLAGS=$2; shift 2
cd "$(dirname "$0")/../pybitup_b200/csrc"
mkdir -p build_$NAME ../_lib/variants
OBJS=""
for f in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $FLAGS -c $f -o build_$NAME/${f%.cu}.o &
done
wait
for o in build/*.o; do
  b=$(basename $o)
  if [ -f build_$NAME/$b ]; then OBJS="$OBJS build_$NAME/$b"; else OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_lib/variants/lib$NAME.so $OBJS
echo "built pybitup_b200/_lib/variants/lib$NAME.so"

#!/bin/bash
# Build kernel-tuning variants of libug4b200.so under gpurun_variants/<name>/ (travels to the GPU box; git-ignored).
#   scripts/build_variants.sh name1 "-DUG_XS_NST=3 -DUG_XS_MINCTA=1" name2 "..." ...
set -eu
ROOT=$(cd $(dirname $0)/.. && pwd)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=$ROOT/gpurun_variants/$name
  mkdir -p $d
  ( UG4B200_LIBDIR=$d UG4B200_EXTRA_NVCC="$flags" python $ROOT/ugcore_b200/build.py --force > $d/build.log 2>&1 && echo "built $name" || { echo "FAILED $name"; tail -5 $d/build.log; } ) &
done
wait

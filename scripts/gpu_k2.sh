#!/bin/bash
set -u
echo "--- fmad=false VI"; python scripts/kbench.py 7 | cut -c1-420
echo "--- fmad=true VI"; UG4B200_LIBDIR=$PWD/gpurun_variants/fma python scripts/kbench.py 7 | cut -c1-420
echo "--- fmad=true plain"; UG4B200_NO_COMPRESS=1 UG4B200_LIBDIR=$PWD/gpurun_variants/fma python scripts/kbench.py 7 | cut -c1-420

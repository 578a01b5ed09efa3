#!/bin/bash
# usage: bash tools/r02_ncu_lib.sh <lib.so> <tag> [fixture]: full ncu capture of the pixel kernel of one library build
LIB=$1; TAG=$2; FIX=${3:-bunny_4k}
RT_B200_LIB=$PWD/$LIB timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel --launch-skip 2 -c 1 -o gpurun_out/$TAG -f python tools/prof_target.py $FIX > gpurun_out/$TAG.log 2>&1
tail -2 gpurun_out/$TAG.log

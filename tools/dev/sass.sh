#!/bin/bash
# usage: tools/dev/sass.sh [extra nvcc flags]  -> /tmp/dev.sass (instructions only), prints register / spill summary
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo -Xptxas -v "$@" -cubin -o /tmp/dev.cubin tools/dev/dev_kernel.cu 2>&1 | grep -E "error|Used|spill" 
cuobjdump -sass /tmp/dev.cubin | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-100 > /tmp/dev.sass
grep -c "^        /\*" /tmp/dev.sass

#!/bin/bash
# kstat.sh <inst-name> [extra nvcc flags]: compile one instantiation file, print registers/spills of
# the bench variants and dump SASS to /tmp/<inst-name>.sass  (development aid)
set -e
name=$1; shift
cd /root/repo/flint_b200/csrc
nvcc -std=c++17 -O3 --fmad=false -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v -I . "$@" -c inst/inst_$name.cu -o /tmp/$name.o 2>&1 \
  | grep -A2 "Compiling entry function" | grep -E "Compiling|registers|spill" | sed -e 's/ptxas info    ://' -e "s/Compiling entry function '_ZN5flint20erk_integrate_kernelINS_//" -e "s/EEEvNS_5KArgsE' for 'sm_100a'//" | paste - - - | cut -c1-220
cuobjdump -sass /tmp/$name.o > /tmp/$name.sass

#!/bin/bash
# kstat.sh <inst-name> [extra nvcc flags]: compile one instantiation file, print registers/spills of
# its kernels and dump SASS to /tmp/<inst-name>.sass  (development aid)
name=$1; shift
cd /root/repo/flint_b200/csrc
nvcc -std=c++17 -O3 --fmad=false -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v -I . "$@" -c inst/inst_$name.cu -o /tmp/$name.o > /tmp/$name.log 2>&1
python3 - /tmp/$name.log <<'PY'
import re,sys
t=open(sys.argv[1]).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t):
    n=m.group(1)
    n=re.sub(r'_ZN5flint','',n); n=re.sub(r'INS_\d+Method','<',n); n=re.sub(r'ENS_\d+Sys',',',n); n=n.replace('EEEvNS_5KArgsE','>')
    n=re.sub(r'ELb([01])',r',\1',n); n=re.sub(r'ELi(\d)',r',m\1',n)
    print("%-60s stack %5s spillst %5s spillld %5s regs %s"%(n,m.group(2),m.group(3),m.group(4),m.group(5)))
PY
cuobjdump -sass /tmp/$name.o > /tmp/$name.sass

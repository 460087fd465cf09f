#!/bin/bash
# register / spill report of every kernel (ptxas -v), no GPU needed
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xptxas -v -I include "$@" -c -o /tmp/mrt_regs.o mass_raytrace_b200/csrc/mrt_cuda.cu 2>&1 | python3 -c "
import sys,re,subprocess
name=None
for l in sys.stdin:
    m=re.search(r\"Compiling entry function '(\S+)'\",l)
    if m: name=subprocess.run(['c++filt',m.group(1)],capture_output=True,text=True).stdout.strip().split('(')[0]
    m=re.search(r'Used (\d+) registers',l)
    if m and name: print(f'{name:50s} {m.group(1)} regs', end='')
    m=re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads',l)
    if m and name: stack=m.groups()
    if 'Used' in l and name: print(f'   stack {stack[0]} spill st {stack[1]} ld {stack[2]}'); name=None
"

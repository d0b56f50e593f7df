#!/bin/bash
# per-kernel register / stack / spill report (no GPU needed)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --extended-lambda -Xptxas -v -c -o /tmp/vrj_regs.o vanrijn_b200/csrc/vanrijn_cuda.cu 2>&1 \
 | grep -E "Compiling entry|Used|stack frame" | sed 's/ptxas info    : //' | paste - - - \
 | sed -E "s/Compiling entry function '([^']*)' for 'sm_100a'/\1/" | while read -r sym rest; do echo "$(echo $sym | c++filt | sed -E 's/\(.*//; s/void vrj:://') | $rest"; done | sed -E 's/ bytes cumulative stack size//; s/used 0 barriers, //'

#!/bin/bash
# Builds stand-alone variants of the attention kernel: tools/ubench/build_att.sh <name> "<nvcc -D flags>" ...
cd "$(dirname "$0")"; mkdir -p bin
while [ $# -ge 2 ]; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo $2 \
        -I../../synt_isic_b200/csrc -o bin/att_$1 att_knock.cu 2>&1 | grep -iE "error|warning.*spill" &
    shift 2
done
wait

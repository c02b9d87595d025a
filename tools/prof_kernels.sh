#!/bin/bash
# Runs on the GPU box (gpurun): GPU tests, MUFU microbenchmark, then ncu --set full of the first conv_tc2 launches
# of one sampling step and of the first attention launch.  Reports stay in /tmp; CSV pages come back.
TAG=${1:-r1b}
NCONV=${2:-12}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"
tail -3 gpurun_out/pytest_$TAG.log
[ -x tools/ubench/mufu ] && timeout 60 tools/ubench/mufu > gpurun_out/mufu_$TAG.log 2>&1
python tools/ncu_target.py --steps 1 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc2 -c $NCONV -o /tmp/prof_conv_$TAG \
    python tools/ncu_target.py --steps 1 > gpurun_out/ncu_conv_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_tc_kernel -c 1 -o /tmp/prof_attn_$TAG \
    python tools/ncu_target.py --steps 1 > gpurun_out/ncu_attn_$TAG.log 2>&1
for k in conv attn; do
    f=/tmp/prof_${k}_$TAG.ncu-rep
    [ -f $f ] || continue
    ncu -i $f --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
    ncu -i $f --page source --csv > gpurun_out/${TAG}_${k}_source.csv 2>/dev/null
    ls -la $f
done
ls -la gpurun_out | tail -12

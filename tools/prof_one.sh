#!/bin/bash
# ncu --set full of the first COUNT launches matching REGEX in one eager sampling step; CSV pages come back.
TAG=$1; REGEX=$2; COUNT=${3:-1}; SKIP=${4:-0}
mkdir -p gpurun_out
python tools/ncu_target.py --steps 1 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --import-source on --clock-control none -k regex:$REGEX -s $SKIP -c $COUNT -o /tmp/prof_$TAG \
    python tools/ncu_target.py --steps 1 > gpurun_out/ncu_$TAG.log 2>&1
f=/tmp/prof_$TAG.ncu-rep
ncu -i $f --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i $f --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la $f gpurun_out/${TAG}_*

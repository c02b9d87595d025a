#!/bin/bash
# Round-end evidence on the GPU box: (1) bench.py plain, then the same command under
# `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` (launch list);
# (2) `ncu --set full` of representative launches of the top kernels; CSV pages come back, reports stay in /tmp;
# (3) the 512-image ResNet18 forward: launch list + full set of the front-end kernel.
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick"
$CMD > gpurun_out/bench_plain_$TAG.json 2> gpurun_out/bench_plain_$TAG.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
[ "$2" = "launches-only" ] && exit 0
python tools/ncu_target.py --steps 1 > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain target failed"; exit 1; }
# conv_tc2 launches of one eager step: 0-3 = N 64 at 128x128, 4-7 = N 128 at 64x64, 8,9 = N 256 at 32x32, 10 = qkv 1x1, 11 = out-proj
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_tc2 -c 12 -o /tmp/prof_conv_$TAG \
    python tools/ncu_target.py --steps 1 > gpurun_out/ncu_conv_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attention_tc_kernel|conv_out3_mma|conv_in_tc|gn_apply" -c 5 -o /tmp/prof_misc_$TAG \
    python tools/ncu_target.py --steps 1 > gpurun_out/ncu_misc_$TAG.log 2>&1
python tools/ncu_resnet.py > gpurun_out/plain_resnet_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}_resnet.csv python tools/ncu_resnet.py > gpurun_out/ncu_resnet_$TAG.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:stem_tc -c 1 -o /tmp/prof_stem_$TAG \
    python tools/ncu_resnet.py > gpurun_out/ncu_stem_$TAG.log 2>&1
for k in conv misc stem; do
    f=/tmp/prof_${k}_$TAG.ncu-rep
    [ -f $f ] || continue
    ncu -i $f --page raw --csv > gpurun_out/${TAG}_${k}_raw.csv 2>/dev/null
    ls -la $f
done
ls -la gpurun_out | tail -12

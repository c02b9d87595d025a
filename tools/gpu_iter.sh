#!/bin/bash
# One optimisation iteration on the GPU box: GPU tests (fail fast), a short bench, the per-layer profile.
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
rc=$?
tail -15 gpurun_out/pytest_$TAG.log
echo "pytest rc=$rc"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("ms_per_step",d["ms_per_step"],"img/s",d["value"],"e2e",d["e2e"]["value"])
    print(d["step_breakdown_ms"])
    print("roofline",d["roofline"]["achieved"],d["roofline"]["frac"],"clocks",d["clocks"])
    print("time_shap",d.get("time_shap"))
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
timeout 300 python tools/profile_layers.py --out gpurun_out/layers_$TAG.txt > gpurun_out/layers_$TAG.log 2>&1
head -45 gpurun_out/layers_$TAG.txt

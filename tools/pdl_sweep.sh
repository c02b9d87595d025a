#!/bin/bash
# ms/step of the sampling loop for the SYNT_PDL modes (0 off, 2 GEMM kernels only, 1 every kernel)
for m in 0 2 1; do
  SYNT_PDL=$m python bench.py --quick --no-cpu-baseline 2>/dev/null > /tmp/pdl_$m.json
  python - "$m" <<'PY'
import json, sys
m = sys.argv[1]
d = json.loads(open(f"/tmp/pdl_{m}.json").read().strip().splitlines()[-1])
print("SYNT_PDL", m, "ms_per_step", round(d["ms_per_step"], 4), "img/s", round(d["value"], 4))
PY
done

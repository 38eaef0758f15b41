#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_r3b.json 2> gpurun_out/e2e_probe_r3b.err; echo "probe rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/e2e_probe_r3b.json"))
for x in r[3:]: print(x)
PY

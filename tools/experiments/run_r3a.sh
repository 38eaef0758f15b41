#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_r3a.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r3a.log
timeout 200 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_r3a.json 2> gpurun_out/e2e_probe_r3a.err; echo "probe rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/e2e_probe_r3a.json"))
for x in r[2:]: print(x)
PY

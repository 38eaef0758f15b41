#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_probe.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_probe.log
timeout 200 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_probe.json 2> gpurun_out/e2e_probe_probe.err; echo "probe rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/e2e_probe_probe.json"))
for x in r[3:]: print(x)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_probe_probe.csv python tools/experiments/e2e_probe.py 4096 > /dev/null 2>&1; echo "ncu rc=$?"

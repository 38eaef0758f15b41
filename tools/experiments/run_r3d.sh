#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r3d
: > $O.jsonl
timeout 400 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_r3d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r3d.log
run() { FE_KERNELS=phased timeout 200 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $@" >> $O.jsonl; }
run 9 2048 0 10
run 3 3500 0 10
run 4 4096 0 10
python - <<'PY'
import json
for line in open("gpurun_out/r3d.jsonl"):
    try:
        r = json.loads(line)
        print(r["kind"], r["n"], r["perturb"], r.get("phased"))
    except Exception:
        print(line.strip()[:300])
PY

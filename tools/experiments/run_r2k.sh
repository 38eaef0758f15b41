#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2k.jsonl
run() { FE_KERNELS=ws timeout 90 python tools/experiments/exp_ws.py $1 >> gpurun_out/r2k.jsonl 2>> gpurun_out/r2k.log || echo "FAILED $1" >> gpurun_out/r2k.jsonl; }
run "4 600 1 5"; run "4 4096 0 10"; run "4 4096 1 10"; run "4 8192 0 5"
python - <<'PY'
import json
for line in open("gpurun_out/r2k.jsonl"):
    try:
        r = json.loads(line)
        print(r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY
timeout 900 python -m pytest tests -m "gpu and not slow" -q --timeout 600 2>&1 | tail -5
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_f.json 2> gpurun_out/e2e_probe_f.err; tail -42 gpurun_out/e2e_probe_f.json

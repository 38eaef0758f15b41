#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2m.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2m.jsonl 2>> gpurun_out/r2m.log || echo "FAILED $1 $2" >> gpurun_out/r2m.jsonl; }
for v in p8c6d2 p8c7d1 p4c10d2; do run $v "4 4096 0 10"; done
FE_KERNELS=ws timeout 90 python tools/experiments/exp_ws.py 4 4096 1 10 >> gpurun_out/r2m.jsonl 2>> gpurun_out/r2m.log
python - <<'PY'
import json
for line in open("gpurun_out/r2m.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"), r.get("ws_stats", {}).get("mode"), r.get("ws_stats", {}).get("ntiles"))
    except Exception:
        print(line.strip())
PY
timeout 900 python -m pytest tests -m "gpu and not slow" -q --timeout 600 2>&1 | tail -5
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_g.json 2> gpurun_out/e2e_probe_g.err; tail -42 gpurun_out/e2e_probe_g.json

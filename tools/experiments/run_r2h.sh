#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2h.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 90 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2h.jsonl 2>> gpurun_out/r2h.log || echo "FAILED $1 $2" >> gpurun_out/r2h.jsonl; }
run rel "4 600 1 5"
for v in rel ef p5c9 p7c7; do run $v "4 4096 0 10"; done
run rel "4 4096 1 10"; run rel "3 3500 1 10"; run ef "3 3500 1 10"
python - <<'PY'
import json
for line in open("gpurun_out/r2h.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"), r.get("ws_shared"))
    except Exception:
        print(line.strip())
PY
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_c.json 2>&1; tail -48 gpurun_out/e2e_probe_c.json
timeout 900 python -m pytest tests -m "gpu and not slow" -q --timeout 600 2>&1 | tail -5
timeout 600 python bench.py --steps 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -c 3500 gpurun_out/bench_r2h.json; tail -3 gpurun_out/bench_r2h.err

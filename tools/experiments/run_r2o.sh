#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2o.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2o.jsonl 2>> gpurun_out/r2o.log || echo "FAILED $1 $2" >> gpurun_out/r2o.jsonl; }
run rs_p8c8d2 "4 600 1 5"
for v in rs_p8c8d2 rs_p8c8d1 rs_m6 rs_m4; do run $v "4 4096 0 10"; done
run rs_p8c8d2 "4 4096 1 10"; run rs_p8c8d2 "4 8192 0 5"
python - <<'PY'
import json
for line in open("gpurun_out/r2o.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY

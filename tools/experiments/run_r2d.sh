#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2d.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2d.jsonl 2>> gpurun_out/r2d.log || echo "FAILED $1 $2" >> gpurun_out/r2d.jsonl; }
run p7c8d1 "4 600 1 5"
for v in p7c8d1 p6c8d2 p7c7d2 p5c9d2 p6c9d1 p5c10d1; do
  run $v "4 4096 0 10"; run $v "3 3500 1 10"
done
run p7c8d1 "4 4096 1 10"
for v in m6 m12 m10 m14 m10d2 m2 m4; do run $v "4 4096 0 10"; done
python - <<'PY'
import json
for line in open("gpurun_out/r2d.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY

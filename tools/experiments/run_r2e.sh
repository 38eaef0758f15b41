#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2e.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2e.jsonl 2>> gpurun_out/r2e.log || echo "FAILED $1 $2" >> gpurun_out/r2e.jsonl; }
run d1c1k "4 600 1 5"
for v in m10c1k m10c512 m10c256 m10c2k m10d2c512 d1c1k d1c512 d1c2k d1c8k d1c256 d2c1k d2c512; do
  run $v "4 4096 0 10"
done
run d1c1k "3 3500 1 10"; run d2c512 "3 3500 1 10"; run d1c1k "4 4096 1 10"
python - <<'PY'
import json
for line in open("gpurun_out/r2e.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2c.jsonl
for v in pw11cw4 pw9cw6 pw7cw8 pw13cw2 pw10cw5 w2pw11 w3pw11; do
  for args in "4 4096 0 10" "4 4096 1 10" "3 3500 1 10"; do
    FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$v.so timeout 200 python tools/experiments/exp_ws.py $args >> gpurun_out/r2c.jsonl 2>> gpurun_out/r2c.log || echo "FAILED $v $args" >> gpurun_out/r2c.jsonl
  done
done
python - <<'PY'
import json
for line in open("gpurun_out/r2c.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY
FE_KERNELS=ws timeout 200 python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/plain_ws.log 2>&1 &&
FE_KERNELS=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_ws -s 2 -c 1 -o gpurun_out/prof_ws_r2c python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/ncu_ws.log 2>&1
tail -2 gpurun_out/ncu_ws.log
timeout 600 python -m pytest tests -m "gpu and not slow" -q --timeout 600 -x 2>&1 | tail -5

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2n.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2n.jsonl 2>> gpurun_out/r2n.log || echo "FAILED $1 $2" >> gpurun_out/r2n.jsonl; }
for v in f_m4 f_m2 f_m6 f_m12 f_m10 f_m14 f_m1; do run $v "4 4096 0 10"; done
python - <<'PY'
import json
for line in open("gpurun_out/r2n.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r.get("ws", {}).get("ms"))
    except Exception:
        print(line.strip())
PY
FE_KERNELS=ws timeout 200 python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/plain_ws.log 2>&1 &&
FE_KERNELS=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_ws -s 2 -c 1 -o gpurun_out/prof_ws_r2n python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/ncu_ws.log 2>&1
tail -2 gpurun_out/ncu_ws.log

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2l.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2l.jsonl 2>> gpurun_out/r2l.log || echo "FAILED $1 $2" >> gpurun_out/r2l.jsonl; }
run u4d2 "4 40 0 3"
run u4d2 "4 600 1 5"
run u4d2 "3 300 1 5"
for v in u4d2 u4d1 u3d2 u4p5c9 u4p7c7 u4m10 u4m4; do run $v "4 4096 0 10"; done
run u4d2 "4 4096 1 10"; run u4d2 "3 3500 1 10"; run u4d2 "4 8192 0 5"
python - <<'PY'
import json
for line in open("gpurun_out/r2l.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"), {k: r.get(k) for k in ("ws_bad", "ws_nan", "ws_first_bad")} if not r.get("ws", {}).get("equal_rows", True) else "")
    except Exception:
        print(line.strip())
PY

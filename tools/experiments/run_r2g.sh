#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2g.jsonl 2>> gpurun_out/r2g.log || echo "FAILED $1 $2" >> gpurun_out/r2g.jsonl; }
run r10q5 "4 40 0 3"
run r10q5 "4 600 1 5"
run r10q5 "3 300 1 5"
for v in r10q5 r10q7 r10q3 r10q8 p6r10q5 p6c9 m10r m10rq7 m4r; do
  run $v "4 4096 0 10"
done
run r10q5 "4 4096 1 10"; run r10q5 "3 3500 1 10"; run r10q5 "4 8192 0 5"
python - <<'PY'
import json
for line in open("gpurun_out/r2g.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"), {k: r.get(k) for k in ("ws_bad", "ws_nan")} if not r.get("ws", {}).get("equal_rows", True) else "")
    except Exception:
        print(line.strip())
PY

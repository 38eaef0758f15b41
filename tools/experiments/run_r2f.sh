#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2f.jsonl
run() { FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$1.so timeout 60 python tools/experiments/exp_ws.py $2 >> gpurun_out/r2f.jsonl 2>> gpurun_out/r2f.log || echo "FAILED $1 $2" >> gpurun_out/r2f.jsonl; }
run d2c "4 600 1 5"
for v in m10d1 m10d2a m10d2c m10d4c m10d8c d1 d2a d2c d4c d3c; do
  run $v "4 4096 0 10"
done
python - <<'PY'
import json
for line in open("gpurun_out/r2f.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip())
PY
timeout 200 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_a.json 2>&1
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 200 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_b.json 2>&1
tail -45 gpurun_out/e2e_probe_a.json; tail -45 gpurun_out/e2e_probe_b.json

#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2b.jsonl
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe.json 2> gpurun_out/e2e_probe.err
for v in base whatif1 whatif2 whatif3 pw11cw4 pw5cw8 pw7cw4; do
  for args in "4 4096 0 10" "3 3500 1 10"; do
    FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$v.so timeout 200 python tools/experiments/exp_ws.py $args >> gpurun_out/r2b.jsonl 2>> gpurun_out/r2b.log || echo "FAILED $v $args" >> gpurun_out/r2b.jsonl
  done
done
python - <<'PY'
import json
for line in open("gpurun_out/r2b.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r.get("ws"))
    except Exception:
        print(line.strip())
PY
cat gpurun_out/e2e_probe.json
# ncu: full capture of the ws kernel on cfg2 (release lib)
FE_KERNELS=ws timeout 200 python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/plain_ws.log 2>&1 &&
FE_KERNELS=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_ws -s 2 -c 2 -o gpurun_out/prof_ws_r2b python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/ncu_ws.log 2>&1
tail -3 gpurun_out/ncu_ws.log

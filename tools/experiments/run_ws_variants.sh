#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/ws_variants
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=ws timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
run release 4 40 0 3
run release 4 600 1 5
run release 3 300 1 5
run release 4 4096 0 10
run release 4 4096 1 10
run release 3 3500 0 10
run release 4 8192 0 5
python - <<'PY'
import json
for line in open("gpurun_out/ws_variants.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip()[:300])
PY
tail -5 $O.log
FE_LIB=tools/experiments/variants/lib_trace.so timeout 150 python tools/experiments/trace_ws.py 4 4096 0 2>&1 | tail -1

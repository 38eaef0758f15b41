#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2y
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=phased timeout 200 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=phased FE_LIB=tools/experiments/variants/lib_$lib.so timeout 200 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
run release 9 37 1 3
run release 9 300 1 5
run q9old 9 2048 0 10
run release 9 2048 0 10
run q9u9 9 2048 0 10
run q9u1 9 2048 0 10
run release 9 2048 1 10
python - <<'PY'
import json
for line in open("gpurun_out/r2y.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("phased"))
    except Exception:
        print(line.strip()[:300])
PY
tail -5 $O.log

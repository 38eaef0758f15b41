#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2p.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=ws timeout 150 python tools/experiments/exp_ws.py $@ >> gpurun_out/r2p.jsonl 2>> gpurun_out/r2p.log || echo "FAILED $lib $@" >> gpurun_out/r2p.jsonl; else FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> gpurun_out/r2p.jsonl 2>> gpurun_out/r2p.log || echo "FAILED $lib $@" >> gpurun_out/r2p.jsonl; fi; }
run release 4 40 0 3
run release 4 600 1 5
run release 3 300 1 5
run release 4 4096 0 10
run rs 4 600 1 5
for v in rs b5 r2 p7c7 m2 m4 m6 rsm2; do run $v 4 4096 0 10; done
run release 4 4096 1 10
run rs 4 4096 1 10
run release 3 3500 0 10
run release 4 8192 0 5
python - <<'PY'
import json
for line in open("gpurun_out/r2p.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip()[:300])
PY
tail -5 gpurun_out/r2p.log

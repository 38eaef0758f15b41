#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r3f
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=phased timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=phased FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
for v in release q9s3k q9s6k q9s12k; do
run $v 9 2048 0 10
done
python - <<'PY'
import json
for line in open("gpurun_out/r3f.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("phased"))
    except Exception:
        print(line.strip()[:300])
PY
tail -3 $O.log

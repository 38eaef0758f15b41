#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r3e
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=ws timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
for v in release noswap deal20 deal20ns; do
run $v 4 600 1 5
run $v 4 4096 0 10
done
run release 4 4096 1 10
run deal20 4 4096 1 10
run release 3 3500 0 10
run deal20 3 3500 0 10
python - <<'PY'
import json
for line in open("gpurun_out/r3e.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip()[:300])
PY
tail -3 $O.log

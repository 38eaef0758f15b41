#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2w
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=ws timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
run release 4 600 1 5
run t64a 4 600 1 5
run t64c 3 300 1 5
for v in release t64a t64b t64c t64d; do run $v 4 4096 0 10; done
run t64a 4 4096 1 10
python - <<'PY'
import json
for line in open("gpurun_out/r2w.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"), r.get("ws_stats",{}).get("ntiles"))
    except Exception:
        print(line.strip()[:300])
PY
tail -5 $O.log

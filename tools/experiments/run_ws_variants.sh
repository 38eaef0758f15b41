#!/bin/bash
# Pattern of the variant runs behind the tables in DESIGN.md: (variant build, mesh) cases through exp_ws.py.
#   build_variants.sh l2h1 "-DFE_WS_L2_HINTS=1" l2h2 "-DFE_WS_L2_HINTS=2";  gpurun -- bash tools/experiments/run_ws_variants.sh
mkdir -p gpurun_out
O=gpurun_out/ws_variants
: > $O.jsonl
run() { lib=$1; shift; if [ "$lib" = release ]; then FE_KERNELS=ws timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; else FE_KERNELS=ws FE_LIB=tools/experiments/variants/lib_$lib.so timeout 150 python tools/experiments/exp_ws.py $@ >> $O.jsonl 2>> $O.log || echo "FAILED $lib $@" >> $O.jsonl; fi; }
for v in ${VARIANTS:-release l2h1 l2h2}; do
  run $v 4 4096 0 20
done
python - <<'PY'
import json
for line in open("gpurun_out/ws_variants.jsonl"):
    try:
        r = json.loads(line)
        print(r["lib"].split("/")[-1], r["kind"], r["n"], r["perturb"], r.get("ws"))
    except Exception:
        print(line.strip()[:300])
PY
tail -3 $O.log

#!/bin/bash
# Round-2 (session 3) measurement batch on one GPU: tests, smoke, bench lines, launch list, ncu captures.
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; echo "ref rc=$?"
for c in cfg3 cfg5 poisson unstructured; do
  timeout 600 python bench.py --config $c > gpurun_out/bench_${c}.json 2> gpurun_out/bench_${c}.err; echo "$c rc=$?"
done
timeout 300 python bench.py --steps 2 --warmup 3 --cg-iters 20 --no-cpu --no-strong > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --cg-iters 20 --no-cpu --no-strong > gpurun_out/ncu_final.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'k_adjacency_fill|k_adjacency_count|k_tile_layout|k_tile_owner|k_tile_ws_records|k_block_sort_small' \
  -c 6 -o gpurun_out/prof_plan_final -f python tools/experiments/e2e_probe.py 4096 > gpurun_out/ncu_plan_final.log 2>&1; echo "ncu plan rc=$?"
python - <<'PY'
import json
for f in ["bench_final","bench_cfg3","bench_cfg5","bench_poisson","bench_unstructured"]:
    try:
        r=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, r["value"], r["ms_per_step"], r["roofline"]["frac"], (r.get("e2e") or {}).get("ms_per_step"), r.get("parity_checked"))
    except Exception as e:
        print(f, "ERR", e)
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not slow" -q --timeout 600 2>&1 | tail -8
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_e.json 2> gpurun_out/e2e_probe_e.err; tail -42 gpurun_out/e2e_probe_e.json
timeout 600 python bench.py --steps 5 > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; python -c "
import json; r=json.load(open('gpurun_out/bench_r2j.json')); print({k: r[k] for k in ('value','ms_per_step')}, r['roofline']['frac'], r['roofline']['ms_per_launch'], r['e2e'], r['config']['symbolic_phase_s'], r['cg']['pcg_iters_per_s'])"; tail -3 gpurun_out/bench_r2j.err
FE_KERNELS=ws timeout 200 python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/plain_ws.log 2>&1 &&
FE_KERNELS=ws timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_ws -s 2 -c 1 -o gpurun_out/prof_ws_r2j python tools/experiments/exp_ws.py 4 4096 0 3 > gpurun_out/ncu_ws.log 2>&1
tail -2 gpurun_out/ncu_ws.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/experiments/e2e_probe.py 4096 > gpurun_out/e2e_probe_d.json 2> gpurun_out/e2e_probe_d.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 500 --csv --log-file gpurun_out/launches_e2e.csv python tools/experiments/e2e_probe.py 4096 > gpurun_out/ncu_e2e.log 2>&1
tail -40 gpurun_out/e2e_probe_d.json
python tools/launch_summary.py gpurun_out/launches_e2e.csv 2>/dev/null | head -50

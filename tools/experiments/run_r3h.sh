#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/pytest_r3h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r3h.log
timeout 600 python bench.py --config cfg3 > gpurun_out/bench_cfg3_r3h.json 2> gpurun_out/bench_cfg3_r3h.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg3_r3h.err
cat gpurun_out/bench_cfg3_r3h.json | head -c 1500

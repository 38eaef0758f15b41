#!/bin/bash
# GPU-box driver for exp_ws.py: small meshes first (a hang costs seconds), each under its own timeout.
mkdir -p gpurun_out
: > gpurun_out/exp_ws.jsonl
for args in "4 40 0 3" "4 600 1 5" "3 300 1 5" "4 4096 0 10" "4 4096 1 10" "3 3500 1 10" "4 8192 0 5"; do
  echo "== $args" >> gpurun_out/exp_ws.log
  timeout 300 python tools/experiments/exp_ws.py $args >> gpurun_out/exp_ws.jsonl 2>> gpurun_out/exp_ws.log || { echo "FAILED $args rc=$?" >> gpurun_out/exp_ws.jsonl; break; }
done
cat gpurun_out/exp_ws.jsonl

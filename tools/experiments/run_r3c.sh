#!/bin/bash
mkdir -p gpurun_out
FE_KERNELS=phased timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assemble_tiled -s 2 -c 1 \
  -o gpurun_out/prof_q9_r3c -f python tools/experiments/exp_ws.py 9 2048 0 3 > gpurun_out/ncu_q9_r3c.log 2>&1
tail -2 gpurun_out/ncu_q9_r3c.log

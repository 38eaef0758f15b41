#!/bin/bash
# ncu --set full of the symbolic / plan kernels on the e2e path (cfg2), one launch each
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'k_adjacency_fill|k_adjacency_count|k_tile_layout|k_tile_ws_records|k_tile_bucket|k_block_sort|k_tile_owner|k_fill_incidence|k_sort_incidence|k_tile_ws_runs|k_tile_conn|k_csr_expand' \
  -c 12 -o gpurun_out/prof_plan_r2z -f python tools/experiments/e2e_probe.py 4096 > gpurun_out/ncu_plan_r2z.log 2>&1
tail -3 gpurun_out/ncu_plan_r2z.log
ls -la gpurun_out/prof_plan_r2z.ncu-rep

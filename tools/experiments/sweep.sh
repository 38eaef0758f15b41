#!/bin/bash
# builds the library with several tile configurations on the GPU box and times the tiled kernel
for cfg in "128,32,3" "64,32,5" "96,32,4" "64,64,4" "128,32,2"; do
  FE_TILE_CONFIG=$cfg python -m feprogram_b200.build --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  grep -A2 "k_assemble_tiledILi4" feprogram_b200/csrc/build/fe_tiles.ptxas.log | grep -E "registers" | head -1
  echo "== config $cfg"; python tools/experiments/dbg.py 2>&1 | grep -E "tiled dbg|max_nodes" | cut -c1-150
done

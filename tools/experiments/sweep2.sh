#!/bin/bash
for defs in "-DFE_SLAB_PAD=0" "-DFE_SLAB_PAD=2" "-DFE_SLAB_PAD=4" "-DFE_SLAB_PAD=6" "-DFE_SLAB_PAD=8" "-DFE_SLAB_PAD=0 -DFE_NO_L2_PREFETCH"; do
  FE_EXTRA_DEFS="$defs" python -m feprogram_b200.build --force > /dev/null 2>&1 || { echo "build failed $defs"; continue; }
  echo "== $defs"; python tools/experiments/dbg.py 2>&1 | grep -E "tiled dbg=0" | cut -c1-150
done

#!/bin/bash
# Developer tool: builds libfeprogram_b200.so variants with extra -D flags into tools/experiments/variants/ (git-ignored
# by *.so, shipped to the GPU box by gpurun).  Usage: build_variants.sh name1 "-DA=1 -DB=2" name2 "-DC" ...
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/experiments/variants
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  FE_EXTRA_DEFS="$defs" python -m feprogram_b200.build --force > /dev/null
  cp feprogram_b200/libfeprogram_b200.so tools/experiments/variants/lib_$name.so
  grep -h -A2 "k_assemble_wsILi4ELb0" feprogram_b200/csrc/build/fe_tiles_ws.ptxas.log | grep "Used" | sed "s/^/$name: /"
done
python -m feprogram_b200.build --force > /dev/null     # restore the release build

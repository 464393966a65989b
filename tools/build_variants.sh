#!/bin/bash
# Compile-time experiment variants of libparrot_b200.so (same sources, different -D knobs) into build_variants/.
# Select one at run time with PARROT_B200_LIB=build_variants/<name>.so ; A/B them back to back in ONE gpurun call.
#   tools/build_variants.sh name1 "-DPB_X=1 -DPB_Y=2" name2 "..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  echo "== $name: $flags"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags -Xptxas -v \
    -o build_variants/$name.so parrot_b200/csrc/api.cu -lcuda 2>&1 | grep -A2 -E "scan_fwd_grouped|scan_bwd_grouped" | grep -E "spill|error" || true
done

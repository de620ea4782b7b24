#!/bin/bash
# tools/_bin/sweep2d_* : one binary per "threads stages rps5 rps3 minblocks [warpspec]".
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
build() {
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I warmup-fir-filter_b200/csrc \
    -DFIRGPU_T2_THREADS=$1 -DFIRGPU_T2_STAGES=$2 -DFIRGPU_T2_RPS5=$3 -DFIRGPU_T2_RPS3=$4 -DFIRGPU_T2_MINBLOCKS=$5 -DFIRGPU_T2_WARPSPEC=${6:-1} \
    -o tools/_bin/sweep2d_T$1_G$2_R$3_$4_M$5_W${6:-1} tools/sweep_tiled.cu
}
for cfg in "$@"; do
  build $cfg &
  while [ "$(jobs -r | wc -l)" -ge 8 ]; do sleep 0.5; done
done
wait
ls tools/_bin | grep sweep2d

#!/bin/bash
# Builds tools/_bin/sweep_* : one binary per tunable configuration "threads sub stages minblocks store_cs".
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
build() {
  local t=$1 s=$2 g=$3 m=$4 c=$5
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I warmup-fir-filter_b200/csrc \
    -DFIRGPU_K_THREADS=$t -DFIRGPU_K_SUB=$s -DFIRGPU_K_STAGES=$g -DFIRGPU_K_MINBLOCKS=$m -DFIRGPU_STORE_CS=$c \
    -o tools/_bin/sweep_T${t}_S${s}_G${g}_M${m}_C${c} tools/sweep_staged.cu
}
for cfg in "$@"; do
  build $cfg &
  while [ "$(jobs -r | wc -l)" -ge 8 ]; do sleep 0.5; done
done
wait
ls tools/_bin

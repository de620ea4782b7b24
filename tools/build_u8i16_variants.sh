#!/bin/bash
# Library variants that differ only in the tile configuration of the uint8-in / int16-out staged kernel
# ("sub,stages,minblocks,ctas_per_sm", 256 threads): tools/_bin/libfirgpu_u8i16_<tag>.so, selected with FIRGPU_LIB.
#   tools/build_u8i16_variants.sh "2,2,2,2" "8,2,1,1" ...
set -e
cd "$(dirname "$0")/../warmup-fir-filter_b200/csrc"
mkdir -p ../../tools/_bin _obj_variants
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ',' '_')
  IFS=, read sub stages minb ctas <<< "$cfg"
  ( nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xfatbin=-compress-all -Xcompiler -fPIC \
      -DFIRGPU_U8I16_SUB=$sub -DFIRGPU_U8I16_STAGES=$stages -DFIRGPU_U8I16_MINBLOCKS=$minb -DFIRGPU_U8I16_CTAS=$ctas -DFIRGPU_STAGED_IN=uint8_t -DFIRGPU_STAGED_OUT=int16_t -DFIRGPU_STAGED_DIL=1 \
      -DFIRGPU_STAGED_NAME=launch_staged_u8_i16_d1 -c fir1d_staged_inst.cu -o _obj_variants/u8_i16_1_$tag.o
    objs=$(ls _obj/*.o | grep -v fir1d_staged_u8_i16_1.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_bin/libfirgpu_u8i16_$tag.so $objs _obj_variants/u8_i16_1_$tag.o \
      -lcudart_static -lpthread -ldl -lrt ) &
done
wait
ls -la ../../tools/_bin/libfirgpu_u8i16_*.so

#!/bin/bash
# Build an experimental variant of the library for A/B runs: scripts/build_variant.sh <suffix> <extra nvcc flags...>
# -> vggt_qwen3_b200/libvq_b200_<suffix>.so (select it with VQ_B200_LIB=...)
set -e
sfx=$1; shift
cd "$(dirname "$0")/.."
objs=""
mkdir -p vggt_qwen3_b200/build_$sfx
for f in common gemm attention rowops aggregator perceiver_ops perceiver_attn perceiver trainops preprocess geom; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC "$@" \
       -c vggt_qwen3_b200/csrc/$f.cu -o vggt_qwen3_b200/build_$sfx/$f.o &
  objs="$objs vggt_qwen3_b200/build_$sfx/$f.o"
done
wait
nvcc -shared -o vggt_qwen3_b200/libvq_b200_$sfx.so $objs -gencode arch=compute_100a,code=sm_100a
echo vggt_qwen3_b200/libvq_b200_$sfx.so

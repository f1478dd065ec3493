#!/bin/bash
# Builds tools/bin/attn_bench_<tag> for a list of "-D..." variants of attention_tcgen05.cu (experiment harness).
# usage: tools/build_attn_variants.sh tag1:"-DX=1 -DY" tag2:"..."
set -e
cd "$(dirname "$0")/.."
C=candle_video_b200/csrc
mkdir -p tools/bin
pids=()
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -I $C $flags \
      -o tools/bin/attn_bench_$tag tools/attn_bench.cu $C/attention_tcgen05.cu $C/tensormap.cc $C/profile.cc $C/options.cc -cudart static \
    && echo "built attn_bench_$tag" ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done

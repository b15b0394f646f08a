#!/bin/bash
# Rebuilds encoder.o on the GPU box with different text_mega.cuh settings and times the batch-1 text tower.
# usage: tools/tmega_sweep.sh "<bn_qkv> <bn_proj> <bn_fc1> <bn_fc2> <stages> <threads> [rot] [bk]" ...   (then restores the default build)
cd meme-search-engine_b200/csrc
for v in "$@"; do
  set -- $v
  rm -f build/encoder.o
  make EXTRA="-DMSE_TMEGA_BN_QKV=$1 -DMSE_TMEGA_BN_PROJ=$2 -DMSE_TMEGA_BN_FC1=$3 -DMSE_TMEGA_BN_FC2=$4 -DMSE_TMEGA_STAGES=$5 -DMSE_TMEGA_THREADS=$6 -DMSE_TMEGA_ROT=${7:-1} -DMSE_TMEGA_BK=${8:-64} -DMSE_TMEGA_BARRIER=${9:-0}" > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  for dbg in ${DEBUGS:-0}; do
    echo -n "bn=$1/$2/$3/$4 stages=$5 threads=$6 rot=${7:-1} bk=${8:-64} barrier=${9:-0} debug=$dbg: "
    (cd ../.. && MSE_TMEGA_DEBUG=$dbg timeout -k 10 120 python tools/text_latency.py 1 2>/dev/null | cut -c50-200)
  done
done
rm -f build/encoder.o; make > /dev/null 2>&1

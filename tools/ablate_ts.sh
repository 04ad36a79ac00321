#!/bin/bash
# k_conv_ts ablations on the cfg-2 tile: per-debug-mask conv time by C_out (TL_TS_DEBUG bits: 1 no MMA, 2 no row loads,
# 4 no epilogue memory ops, 8 no weight copies, 16 no tcgen05.st).  usage: tools/ablate_ts.sh <outdir> [mode] [masks...]
out=$1; mode=${2:-f16}; shift; shift; masks=${@:-0 1 2 4 16 18 19 23}; mkdir -p $out
for dbg in $masks; do
  TL_TS_DEBUG=$dbg timeout 300 python tools/profile_layers.py cfg2_2M $mode > $out/layers_${mode}_dbg$dbg.txt 2>&1
  echo "== TL_TS_DEBUG=$dbg"; sed -n 6,9p $out/layers_${mode}_dbg$dbg.txt
done

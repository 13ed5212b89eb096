#!/bin/bash
# Decomposition of the 16/32-channel full-resolution decoder layers (halo kernel): which side bounds them?
# MSP_CONV_DEBUG: 1 no global stores, 2 no statistics, 128 no A loads, 256 no MMAs
out=gpurun_out/r02d/narrow.txt; mkdir -p gpurun_out/r02d; : > $out
for cfg in "MSP_CONV_DEBUG=0" "MSP_CONV_DEBUG=1" "MSP_CONV_DEBUG=2" "MSP_CONV_DEBUG=3" "MSP_CONV_DEBUG=128" "MSP_CONV_DEBUG=256" "MSP_CONV_DEBUG=384" "MSP_CONV_DEBUG=387" "MSP_CONV_MULTI=0" "MSP_CONV_NOHALO=1"; do
  echo "== $cfg" >> $out
  env $cfg python tools/bench_conv.py --model unet50 --batch 24 --only d4_c --kinds fprop,dgrad >> $out 2>&1
  env $cfg python tools/bench_conv.py --model unet50 --batch 24 --only d3_c1 --kinds fprop >> $out 2>&1
done

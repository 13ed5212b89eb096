#!/bin/bash
# Decomposition of the 16/32-channel full-resolution decoder layers (halo kernel): which side bounds them?
# MSP_CONV_DEBUG: 1 no global stores, 2 no statistics, 4 no accumulator read (multi epilogue), 128 no A loads, 256 no MMAs
out=gpurun_out/r02e/narrow.txt; mkdir -p gpurun_out/r02e; : > $out
for cfg in "MSP_CONV_DEBUG=0" "MSP_CONV_DEBUG=1" "MSP_CONV_DEBUG=3" "MSP_CONV_DEBUG=4" "MSP_CONV_DEBUG=260" "MSP_CONV_DEBUG=388" "MSP_CONV_DEBUG=387" \
           "MSP_CONV_MULTI=0 MSP_CONV_DEBUG=1" "MSP_CONV_MULTI=0 MSP_CONV_DEBUG=3" "MSP_CONV_MULTI=0 MSP_CONV_DEBUG=259" "MSP_CONV_MULTI=0 MSP_CONV_DEBUG=387"; do
  echo "== $cfg" >> $out
  env $cfg python tools/bench_conv.py --model unet50 --batch 24 --only d4_c --kinds fprop >> $out 2>&1
done

#!/bin/bash
out=gpurun_out/r02g/narrow.txt; mkdir -p gpurun_out/r02g; : > $out
for cfg in "MSP_CONV_DEBUG=0" "MSP_CONV_DEBUG=4" "MSP_CONV_DEBUG=128" "MSP_CONV_DEBUG=132" "MSP_CONV_MULTI=0"; do
  echo "== $cfg" >> $out
  env $cfg python tools/bench_conv.py --model unet50 --batch 24 --only d4_ --kinds fprop,dgrad >> $out 2>&1
  env $cfg python tools/bench_conv.py --model unet50 --batch 24 --only d3_ --kinds fprop,dgrad >> $out 2>&1
done

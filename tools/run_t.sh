mkdir -p gpurun_out/r02z
for v in 0 2; do echo "== MSP_CONV_2CTA=$v"; MSP_CONV_2CTA=$v timeout 200 python tools/bench_conv.py --model r50 --batch 256 --only stem --kinds fprop 2>&1 | tail -3; MSP_CONV_2CTA=$v timeout 200 python tools/bench_conv.py --model r50 --batch 256 --only L0b --kinds fprop,dgrad 2>&1 | tail -8; done > gpurun_out/r02z/pair.txt 2>&1
true

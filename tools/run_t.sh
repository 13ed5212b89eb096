mkdir -p gpurun_out/r02t
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02t/tests.log 2>&1
timeout 600 python tools/bench_conv.py --model unet50 --batch 24 --only d4_ > gpurun_out/r02t/conv.txt 2>&1
timeout 600 python tools/bench_conv.py --model unet50 --batch 24 --only d3_ >> gpurun_out/r02t/conv.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02t/bench.log 2> gpurun_out/r02t/bench.err
true

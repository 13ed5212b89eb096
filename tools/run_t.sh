mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02x/tests.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02x/bench.log 2> gpurun_out/r02x/bench.err
MSP_WGRAD_CHUNK=100000 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02x/bench_nochunk.log 2> gpurun_out/r02x/bench_nochunk.err
true

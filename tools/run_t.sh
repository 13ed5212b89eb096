mkdir -p gpurun_out/r02r
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02r/tests.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --graph-timeline gpurun_out/r02r/tl > gpurun_out/r02r/bench.log 2> gpurun_out/r02r/bench.err
true

mkdir -p gpurun_out/r02y
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_hotpath_gpu.py -x -q -k "maxpool or resnet50 or resnet18 or eval_mode" > gpurun_out/r02y/tests.log 2>&1
python tools/bench_pool.py > gpurun_out/r02y/pool_new.txt 2>&1
true

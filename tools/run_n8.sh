mkdir -p gpurun_out/r02n8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02n8/bench_n8.log 2> gpurun_out/r02n8/bench_n8.err
true

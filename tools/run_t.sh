mkdir -p gpurun_out/r02f8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02f8/bench_n8.log 2> gpurun_out/r02f8/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 3 --no-parity-check > gpurun_out/r02f8/bench_n4.log 2> gpurun_out/r02f8/bench_n4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity-check > gpurun_out/r02f8/bench_n2.log 2> gpurun_out/r02f8/bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f8/bench_n1.log 2> gpurun_out/r02f8/bench_n1.err
true

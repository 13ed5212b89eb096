mkdir -p gpurun_out/r02o
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --workload cfg3 --secondary none --no-parity-check --graph-timeline gpurun_out/r02o/tl_n2 > gpurun_out/r02o/bench_n2.log 2> gpurun_out/r02o/bench_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload cfg3 --secondary none --no-cpu-baseline --graph-timeline gpurun_out/r02o/tl_n1 > gpurun_out/r02o/bench_n1.log 2> gpurun_out/r02o/bench_n1.err
true

mkdir -p gpurun_out/r02u
python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --secondary none > gpurun_out/r02u/plain_cfg2.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02u/ncu_traffic_cfg2.csv python bench.py --workload cfg2 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --secondary none > gpurun_out/r02u/ncu_cfg2.log 2>&1
python bench.py --workload cfg3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --secondary none > gpurun_out/r02u/plain_cfg3.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 4400 --csv --log-file gpurun_out/r02u/ncu_traffic_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --secondary none > gpurun_out/r02u/ncu_cfg3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02u/bench.log 2> gpurun_out/r02u/bench.err
true

python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/per_op_now.json > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err
python bench.py --steps 10 --warmup 3 --global-batch 32 --no-latency --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err

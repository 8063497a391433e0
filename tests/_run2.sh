for ch in 16 32 64 256; do
SDN_PRE_CHUNK=$ch timeout 300 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --profile-out gpurun_out/per_op_ch$ch.json > gpurun_out/bench_ch$ch.json 2> gpurun_out/bench_ch$ch.err
done

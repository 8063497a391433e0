timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.txt
for pdl in 1 0; do
SDN_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --global-batch 32 --no-latency --no-cpu-baseline > gpurun_out/bench_b32_pdl$pdl.json 2> gpurun_out/bench_b32_pdl$pdl.err
SDN_PDL=$pdl timeout 300 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/bench_b256_pdl$pdl.json 2> gpurun_out/bench_b256_pdl$pdl.err
done

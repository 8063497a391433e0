timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --profile-out gpurun_out/per_op_dg.json > gpurun_out/bench_dg.json 2> gpurun_out/bench_dg.err

python -m pytest tests -x -q -m gpu > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.txt
python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/bench_e2e2.json 2> gpurun_out/bench_e2e2.err

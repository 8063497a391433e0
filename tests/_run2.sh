timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.txt
for i in 1 2; do
python tests/gpu_ablate.py
SDN_LIB_NAME=libsdn_b200_alt.so python tests/gpu_ablate.py
done > gpurun_out/ab_mma2.txt 2>&1
for lib in libsdn_b200.so libsdn_b200_alt.so libsdn_b200.so libsdn_b200_alt.so; do
SDN_LIB_NAME=$lib python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['value'], d['ms_per_step'])"
done >> gpurun_out/ab_mma2.txt 2>&1

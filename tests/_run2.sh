timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.txt
for ov in 1 0 1 0; do
SDN_WGRAD_OVERLAP=$ov timeout 300 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap $ov', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done > gpurun_out/ab_overlap.txt 2>&1
for ov in 1 0; do
SDN_WGRAD_OVERLAP=$ov timeout 300 python bench.py --steps 10 --warmup 3 --global-batch 32 --no-latency --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b32 overlap $ov', d['value'], d['ms_per_step'])"
done >> gpurun_out/ab_overlap.txt 2>&1

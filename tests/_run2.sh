for w in 1 2 1 2; do
SDN_CONVT_WGRAD_WAVES=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-latency --no-cpu-baseline --profile-out gpurun_out/per_op_w$w.json > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_w$w.json').read().strip().splitlines()[-1]); p=json.load(open('gpurun_out/per_op_w$w.json'))
print($w, round(d['value'],1), round(d['ms_per_step'],3), ' '.join(f"L{r['layer']}:{r['ms']/r['calls']:.3f}" for r in p['rows'] if r['name']=='convT_wgrad'))
PY
done > gpurun_out/convt_waves.txt 2>&1

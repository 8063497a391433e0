set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/per_op_final.json > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/bench_pre_ncu.json 2> gpurun_out/bench_pre_ncu.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 740 -c 490 --csv --log-file gpurun_out/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1

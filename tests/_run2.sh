set -x
python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/bench_pre_ncu.json 2> gpurun_out/bench_pre_ncu.err && \
timeout 1200 ncu --set full --clock-control none -k regex:"bn_bwd_reduce|bn_bwd_apply|wgrad_tr_kernel|conv_gemm_kernel" -s 258 -c 86 -o /tmp/ncu_full python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la /tmp/ncu_full.ncu-rep
ncu -i /tmp/ncu_full.ncu-rep --page raw --csv > gpurun_out/ncu_full_raw.csv 2> gpurun_out/ncu_full_raw.err
SZ=$(stat -c %s /tmp/ncu_full.ncu-rep); if [ "$SZ" -lt 45000000 ]; then cp /tmp/ncu_full.ncu-rep gpurun_out/r1_ncu_full.ncu-rep; fi

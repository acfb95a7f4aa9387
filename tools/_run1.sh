set -x
timeout -k 5 900 python bench.py > gpurun_out/bench_r1_3dB_fused.json 2> gpurun_out/bench_r1_3dB_fused.err
tail -c 600 gpurun_out/bench_r1_3dB_fused.json
timeout -k 5 300 python bench.py --no-cpu --snr 5.0 > gpurun_out/bench_r1_5dB_fused.json 2>/dev/null
timeout -k 5 300 python bench.py --no-cpu --snr 4.0 > gpurun_out/bench_r1_4dB_fused.json 2>/dev/null
timeout -k 5 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_bench.log 2>&1 && \
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_fused.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 1 -f -o gpurun_out/prof_r1_bench_fused python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench2.log 2>&1
tail -3 gpurun_out/ncu_bench2.log

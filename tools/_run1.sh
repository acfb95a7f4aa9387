set -x
timeout -k 5 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 900 python bench.py > gpurun_out/bench_r1_3dB_fused.json 2> gpurun_out/bench_r1_3dB_fused.err
for snr in 4.0 5.0; do timeout -k 5 300 python bench.py --no-cpu --snr $snr > gpurun_out/bench_r1_${snr%.0}dB_fused.json 2>/dev/null; done
timeout -k 5 400 python bench.py --no-cpu --precision fp64 --demap exact --frames 512 --steps 2 > gpurun_out/bench_r1_fp64.json 2>/dev/null
for f in gpurun_out/bench_r1_3dB_fused.json gpurun_out/bench_r1_4dB_fused.json gpurun_out/bench_r1_5dB_fused.json gpurun_out/bench_r1_fp64.json; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('REC', '$f'.split('/')[-1], 'value',round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'dec ms', round(d['roofline']['launch_ms'],1), 'frac', round(d['roofline']['frac'],3), 'it', round(d['avg_iterations'],2), d['config']['schedule'], d['cpu_baseline'] and round(d['cpu_baseline']['value'],2), d['clocks'])"; done
timeout -k 5 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_bench.log 2>&1 && \
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_fused.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo LAUNCHLIST $?

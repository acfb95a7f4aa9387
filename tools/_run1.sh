set -x
timeout -k 5 900 python bench.py > gpurun_out/bench_r1_3dB_fused.json 2> gpurun_out/bench_r1_3dB_fused.err
for snr in 4.0 5.0; do timeout -k 5 300 python bench.py --no-cpu --snr $snr > gpurun_out/bench_r1_${snr%.0}dB_fused.json 2>/dev/null; done
timeout -k 5 400 python bench.py --no-cpu --precision fp64 --frames 512 --steps 2 > gpurun_out/bench_r1_fp64.json 2>/dev/null
timeout -k 5 400 python bench.py --no-cpu --precision fp64 --frames 512 --steps 2 --schedule 2 --lanes 512 > gpurun_out/bench_r1_fp64_fused.json 2>/dev/null
for f in gpurun_out/bench_r1_3dB_fused.json gpurun_out/bench_r1_4dB_fused.json gpurun_out/bench_r1_5dB_fused.json gpurun_out/bench_r1_fp64.json gpurun_out/bench_r1_fp64_fused.json; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$f'.split('/')[-1], 'value',round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'dec ms', round(d['roofline']['launch_ms'],1), 'frac', round(d['roofline']['frac'],3), 'it', round(d['avg_iterations'],2), d['config']['schedule'])"; done

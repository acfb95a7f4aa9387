set -x
timeout -k 5 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -k 5 300 python tools/sweep_decode.py --frames 2048 --lanes 512,1024 --schedules 2 --fused 32:1:0:0:4 2>&1 | tail -2
QAMRECON_FUSED_STORE_POST=0 timeout -k 5 300 python tools/sweep_decode.py --frames 2048 --lanes 1024 --schedules 2 --fused 32:1:0:0:4 2>&1 | tail -1
timeout -k 5 200 python tools/sweep_decode.py --frames 1024 --snr 5.0 --lanes 512,1024 --schedules 0,2 --fused 32:1:0:0:4 2>&1 | tail -4
QAMRECON_FUSED_STORE_POST=0 timeout -k 5 200 python tools/sweep_decode.py --frames 1024 --snr 5.0 --lanes 1024 --schedules 2 --fused 32:1:0:0:4 2>&1 | tail -1
timeout -k 5 300 python bench.py --no-cpu --steps 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bench value',round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'dec ms', round(d['roofline']['launch_ms'],1), 'frac', round(d['roofline']['frac'],3))"

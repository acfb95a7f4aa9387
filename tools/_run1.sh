timeout -k 5 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for snr in 3.0 5.0; do timeout -k 5 300 python bench.py --no-cpu --snr $snr --steps 4 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('REC $snr value',round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1))"; done

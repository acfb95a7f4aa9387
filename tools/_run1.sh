for snr in 3.0; do
for cfg in "0 512" "2 1024" "2 512"; do set -- $cfg
timeout -k 5 300 python bench.py --steps 3 --warmup 3 --no-cpu --snr $snr --schedule $1 --lanes $2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sched $1 lanes $2', d['config']['workload'][61:74], d['value'], d['e2e']['value'], d['roofline']['launch_ms'], d['roofline']['frac'])"
done; done

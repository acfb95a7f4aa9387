timeout -k 5 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout -k 5 300 python tools/sweep_decode.py --frames 2048 --lanes 1024 --schedules 2 --fused 32:1:0:0:4 --stages 2>&1 | tail -4

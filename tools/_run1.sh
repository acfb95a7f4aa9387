timeout -k 5 300 python tools/sweep_decode.py --frames 2048 --lanes 512,1024 --schedules 2 --fused 32:1:0:0:4 2>&1 | tail -2

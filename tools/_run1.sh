set -x
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain_fixture or refill or fp32_batch" 2>&1 | tail -3
timeout -k 5 400 python tools/sweep_decode.py --frames 2048 --lanes 512,1024 --schedules 2 --fused 32:1:0:0:1:1000,32:1:0:0:2:1000,32:1:0:0:4:1000,32:1:0:0:1:900,32:1:0:0:2:900,32:1:0:0:4:900,32:1:0:0:2:750 2>&1 | tail -15

for pf in 0 4 6; do
QAMRECON_FUSED_PREFETCH=$pf timeout -k 5 300 python tools/sweep_decode.py --frames 2048 --lanes 1024 --schedules 2 --fused 32:1:0:$pf:4 2>&1 | tail -1
done

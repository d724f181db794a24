run() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 6 --warmup 3 --gather-only --gather-micro-batches $MB 2>/dev/null | tail -1 >> gpurun_out/r2_g8.jsonl; }
PORT=29521 MB=4 run RUA_MULTI_CTAS_PER_SM=0
PORT=29522 MB=4 run RUA_MULTI_CTAS_PER_SM=2
PORT=29523 MB=4 run RUA_MULTI_CTAS_PER_SM=8
PORT=29524 MB=8 run RUA_MULTI_CTAS_PER_SM=4
PORT=29525 MB=1 run RUA_MULTI_CTAS_PER_SM=0
PORT=29526 MB=2 run RUA_MULTI_CTAS_PER_SM=16

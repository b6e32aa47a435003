#!/bin/bash
run() { echo -n "ctas=$1 reserve=$2 rows=$3 every=$4: "; YC_NCCL_CTAS=$1 YC_RESERVE_SMS=$2 YC_GATHER_ROWS=$3 YC_GATHER_EVERY=$4 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus $N --steps 200 --warmup 5 --profile 2>&1 | tail -1; }
N=${N:-2}
run 0 0 4096 8
run 0 0 4096 1

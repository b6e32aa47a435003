#!/bin/bash
N=${N:-2}
python bench.py --steps 200 --warmup 5 --profile 2>&1 | tail -1
for every in 8 1; do
echo -n "N=$N every=$every: "; YC_GATHER_EVERY=$every python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 100)) bench.py --gpus $N --steps 200 --warmup 5 --profile 2>&1 | grep profile_run | tail -1
done

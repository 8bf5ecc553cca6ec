#!/bin/bash
# 2-GPU weak-scaling run of bench.py (launched the way the driver does)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3

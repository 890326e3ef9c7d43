#!/bin/bash
# where do the ~60 us per evaluation of the row-sharded path go? (2 GPUs)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]/20*1000,1), "us/eval", d.get("comm"), round(d["value"],1))'
echo "single GPU, n=5e7 (half shard), PDL on:";  LRB_PDL=1 python bench.py --n 50000000 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
echo "single GPU, n=5e7 (half shard), PDL off:"; LRB_PDL=0 python bench.py --n 50000000 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
echo "N=2 p2p:";  $TR --nproc-per-node 2 bench.py --gpus 2 --comm p2p 2>&1 | tail -1 | python -c "$P"
echo "N=2 nccl:"; $TR --nproc-per-node 2 bench.py --gpus 2 --comm nccl 2>&1 | tail -1 | python -c "$P"
echo "two independent single-GPU runs at n=5e7 side by side (no exchange, both GPUs busy):"
(CUDA_VISIBLE_DEVICES=0 python bench.py --n 50000000 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P") &
(CUDA_VISIBLE_DEVICES=1 python bench.py --n 50000000 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P") &
wait

#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]/20*1000,1), "us/eval", d.get("comm"), round(d["value"],1))'
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k p2p 2>&1 | tail -2
for rep in 1 2; do
  echo "p2p, release-only:";  $TR --nproc-per-node 2 bench.py --gpus 2 --comm p2p --steps 40 2>&1 | tail -1 | python -c "$P"
  echo "p2p, fence per thread:"; LRB_P2P_FENCE=1 $TR --nproc-per-node 2 bench.py --gpus 2 --comm p2p --steps 40 2>&1 | tail -1 | python -c "$P"
done
echo "single GPU half shard:"; python bench.py --n 50000000 --no-cpu-baseline --steps 40 2>&1 | tail -1 | python -c "$P"

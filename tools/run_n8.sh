#!/bin/bash
# 8-GPU validation: dist parity tests, c3 strong scaling (p2p / nccl), c4 chain-sharded, c5 fp64
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_dist_n8.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
for n in 8 4; do
  timeout 200 $TR --nproc-per-node $n bench.py --gpus $n --comm p2p 2>&1 | tail -1 | tee gpurun_out/bench_r1_n${n}_p2p.json | cut -c1-400
done
timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --comm nccl 2>&1 | tail -1 | tee gpurun_out/bench_r1_n8_nccl.json | cut -c1-300
timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --workload c4 2>&1 | tail -1 | tee gpurun_out/bench_r1_c4_n8.json | cut -c1-400
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --workload c5 --steps 40 2>&1 | tail -1 | tee gpurun_out/bench_r1_c5_n8.json | cut -c1-700

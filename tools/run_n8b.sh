#!/bin/bash
# final-build 8-GPU validation: dist parity (incl. many-chain), c3 strong scaling N=8,4,2 with comm=auto
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_dist_n8b.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
for n in 8 4 2; do
  timeout 200 $TR --nproc-per-node $n bench.py --gpus $n 2>&1 | tail -1 | tee gpurun_out/bench_r1c_n${n}_auto.json | cut -c1-330
done
timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --workload c4 2>&1 | tail -1 | tee gpurun_out/bench_r1c_c4_n8.json | cut -c1-200
timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200

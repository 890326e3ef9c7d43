#!/bin/bash
# 2-GPU validation of drive mode: parity test + bench N=2 (drive vs static kernel) + N=1 at the half shard
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["mode"], d["n_gpus"], round(d["ms_per_step"]/20*1000,1), "us/eval", d.get("comm"), round(d["value"],1), "evals/s e2e", round(d["e2e"]["value"],1), d["digest"]["final_lpost"], d["digest"]["accepted"], d["digest"]["ranks_agree"])'
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -4
for m in "" "--deterministic"; do
  timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 --no-secondary $m 2>gpurun_out/n2.err | tail -1 | python -c "$P" || tail -5 gpurun_out/n2.err
done
for m in "" "--deterministic"; do
  timeout 300 python bench.py --n 50000000 --no-cpu-baseline --no-secondary $m 2>/dev/null | tail -1 | python -c "$P"
done
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2>gpurun_out/n2.err; tail -c 1500 gpurun_out/r2_bench_n2.json

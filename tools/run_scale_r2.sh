#!/bin/bash
# same-box strong-scaling sweep N=1,2,4,8 (what the driver's SCALE run does), drive mode; A/B with the static kernel at N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
timeout 300 python bench.py --gpus 1 --no-cpu-baseline 2>gpurun_out/scale.err | tail -1 > gpurun_out/scale_r2_n1.json
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n bench.py --gpus $n 2>>gpurun_out/scale.err | tail -1 > gpurun_out/scale_r2_n$n.json
done
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --no-secondary --deterministic 2>>gpurun_out/scale.err | tail -1 > gpurun_out/scale_r2_n8_static.json
timeout 300 $TR --nproc-per-node 8 bench.py --gpus 8 --no-secondary --comm nccl 2>>gpurun_out/scale.err | tail -1 > gpurun_out/scale_r2_n8_nccl.json
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
python - <<'PY'
import json
base=None
for tag in ("n1","n2","n4","n8","n8_static","n8_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/scale_r2_{tag}.json").read())
    except Exception as e:
        print(tag,"failed",e); continue
    base=base or d["value"]
    n=d["n_gpus"]
    c4=(d.get("secondary") or {}).get("c4",{})
    print(tag, d["mode"], d.get("comm"), round(d["value"],1), "evals/s", round(d["ms_per_step"]/20*1000,1), "us/eval", "eff", round(d["value"]/(n*base),4),
          "e2e", round(d["e2e"]["value"],1), "lpost", d["digest"]["final_lpost"], "x_l2", d["digest"]["final_x_l2"], "acc", d["digest"]["accepted"],
          "| c4", round(c4.get("chain_iters_per_s",0)), c4.get("frac"))
PY

#!/bin/bash
# same-box strong-scaling sweep N=1,2,4,8 (what the driver's SCALE run does)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
python bench.py --gpus 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/scale_r1_n1.json
for n in 2 4 8; do
  timeout 200 $TR --nproc-per-node $n bench.py --gpus $n 2>&1 | tail -1 > gpurun_out/scale_r1_n$n.json
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    d=json.loads(open(f"gpurun_out/scale_r1_n{n}.json").read())
    base=base or d["value"]
    print(n, round(d["value"],1), "evals/s", round(d["ms_per_step"],3), "ms/step", "eff", round(d["value"]/(n*base),4), d.get("comm"))
PY

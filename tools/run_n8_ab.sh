#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29655"
for rep in 1 2; do for pdl in 1 0; do
  echo "PDL=$pdl"; LRB_PDL=$pdl timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 40 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['comm'], d['accept_rate'], d['roofline']['achieved'])"
done; done

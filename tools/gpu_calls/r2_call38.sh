#!/bin/bash
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > $O/r2c38_bench2.json 2> $O/r2c38_bench2.err; tail -2 $O/r2c38_bench2.err | cut -c1-250
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c38_bench2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'weak',d['weak'],'e2e',round(d['e2e']['value']))
P

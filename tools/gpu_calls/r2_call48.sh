#!/bin/bash
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 20 --warmup 3 --no-extra > $O/r2c48_bench4.json 2> $O/r2c48_bench4.err; tail -1 $O/r2c48_bench4.err | cut -c1-200
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c48_bench4.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'weak',d['weak'],'e2e',round(d['e2e']['value']), d['roofline']['step']['kernels_ms'])
P

#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialised or contingency or partial_what or scenario_resident" > $O/r2c34_tests.log 2>&1; tail -4 $O/r2c34_tests.log
python bench.py --no-e2e --no-cpu-baseline --no-callbacks --steps 5 > $O/r2c34_bench.json 2>/dev/null
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c34_bench.json').read().strip().splitlines()[-1])
for k,v in d['extra_configs'].items():
    print(k,{a:(b if not isinstance(b,dict) else {x:round(y,4) for x,y in b.items()}) for a,b in v.items()})
P

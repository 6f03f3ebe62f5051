#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident" > $O/r2c32_tests.log 2>&1; tail -12 $O/r2c32_tests.log
timeout 300 python - 4096 <<'P' 2>&1 | grep -v "^$"
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
S = int(sys.argv[1])
for env in ({}, {"PS_NO_PBG": "1"}):
    os.environ.update(env)
    print("--", env or "default (gather-from-shared-memory residual kernel)", flush=True)
    for what in (7, 1):
        sweep.run("case13659pegase", "PBRARVmodel", S, [16], what=what, steps=8)
    sweep.run("case13659pegase", "PBRARVmodel", 512, [16], what=7, steps=8)
    for k in env: del os.environ[k]
P

#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident" > $O/r2c26_tests.log 2>&1; tail -5 $O/r2c26_tests.log
for n in 3 9; do echo "-- stages $n"; PS_PBQ_STAGES=$n python tools/sweep.py ncu_g 2>&1 | grep -v "^$"; done
timeout 300 python - 4096 <<'P' 2>&1 | grep -v "^$"
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
S = int(sys.argv[1])
for env in ({}, {"PS_NO_PBQ": "1"}):
    os.environ.update(env)
    print("--", env or "default (copy-engine-fed residual kernel)", flush=True)
    for what in (7, 1):
        sweep.run("case13659pegase", "PBRARVmodel", S, [16], what=what, steps=8)
    for k in env: del os.environ[k]
P
ncu --set full --clock-control none --import-source on -k regex:pb_bus_g_tma -c 1 -o $O/r02_pbq2 python tools/sweep.py ncu_g > $O/r2c26_ncu.log 2>&1; tail -2 $O/r2c26_ncu.log

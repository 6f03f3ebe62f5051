#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "specialised or partial_what or code_paths" 2>&1 | tail -4
python - <<'PY' > $O/r2c10_gonly.log 2>&1
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
for case, F, S in (("case2869pegase", "PNPAPVmodel", 1024), ("case9241pegase", "PNPAPVmodel", 512), ("case13659pegase", "PNRARVmodel", 512), ("case13659pegase", "CBRARVmodel", 512)):
    for what in (1, 3, 7):
        sweep.run(case, F, S, [16], what=what)
PY
cat $O/r2c10_gonly.log
for L in libpowersense_b200 libps_bg6 libps_bg8; do echo $L; PS_LIB=$PWD/powersense.jl_b200/$L.so python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import sweep
sweep.run('case13659pegase','PBRARVmodel',2048,[8],what=1)
sweep.run('case13659pegase','PBRAPVmodel',1024,[8],what=1)
"; done 2>&1 | grep -v "^$" > $O/r2c10_bg.log; cat $O/r2c10_bg.log

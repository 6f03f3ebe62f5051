#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu -k "specialised or partial_what or code_paths or contingency or random_networks or batch_vs_oracle" > $O/r2c42_tests.log 2>&1; tail -3 $O/r2c42_tests.log
for env in "" "PS_NO_GLIST=1"; do for c in 4 8 16; do
echo "== $env PS_CHUNK_GL=$c"
env $env PS_CHUNK_GL=$c python - <<'P' 2>&1 | grep -v "^$\|kernels"
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case2869pegase", "PNPAPVmodel", 1024, [16], what=1, steps=8)
sweep.run("case9241pegase", "PNPAPVmodel", 512, [16], what=1, steps=8)
sweep.run("case13659pegase", "PNRARVmodel", 512, [16], what=1, steps=8)
P
done; done

#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "PBRA" > $O/r2c29_tests.log 2>&1; tail -3 $O/r2c29_tests.log
for lib in libps_jhrows0.so libpowersense_b200.so; do for c in 4 8 16; do
echo "== $lib PS_CHUNK_PB=$c"
PS_LIB=$PWD/powersense.jl_b200/$lib PS_CHUNK_PB=$c python - <<'P' 2>&1 | grep -v "^$"
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case13659pegase", "PBRARVmodel", 4096, [16], what=7, steps=8)
sweep.run("case13659pegase", "PBRARVmodel", 4096, [16], what=3, steps=8)
P
done; done

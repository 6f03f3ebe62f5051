#!/bin/bash
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r2c43_tests.log 2>&1; tail -3 $O/r2c43_tests.log
timeout 600 python tests/stress_paths.py 60 > $O/r2c43_stress.log 2>&1; tail -1 $O/r2c43_stress.log
for c in 1 2 4; do echo "== PS_CHUNK_GL=$c"; PS_CHUNK_GL=$c python - <<'P' 2>&1 | grep -v "^$\|kernels"
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case2869pegase", "PNPAPVmodel", 1024, [16], what=1, steps=8)
sweep.run("case9241pegase", "PNPAPVmodel", 512, [16], what=1, steps=8)
P
done
python bench.py > $O/r2c43_bench.json 2> $O/r2c43_bench.err; tail -1 $O/r2c43_bench.err | cut -c1-250

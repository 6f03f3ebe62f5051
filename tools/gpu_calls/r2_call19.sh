#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slp.py -x -q -m gpu -k "scenario_resident or code_paths or batch_vs or quadratic or case300_gpu or full_size" 2>&1 | tail -4
python - <<'PY' > $O/r2c19_s512.log 2>&1
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep, os
for S in (512, 2048):
    sweep.run("case13659pegase", "PBRARVmodel", S, [16], what=7)
os.environ["PS_NO_PBSCN"] = "1"
sweep.run("case13659pegase", "PBRARVmodel", 512, [16], what=7)
PY
grep -E "what=|kernels" $O/r2c19_s512.log
python bench.py --no-extra > $O/r2c19_bench.json 2> $O/r2c19_bench.err; tail -1 $O/r2c19_bench.err | cut -c1-400

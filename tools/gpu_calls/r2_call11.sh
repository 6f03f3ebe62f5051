#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident or specialised or partial_what or code_paths or full_size" 2>&1 | tail -4
python - <<'PY' > $O/r2c11_gonly.log 2>&1
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=1)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=1, scal=False)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [4, 16], what=1)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=3)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=7)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=7, scal=False)
os.environ["PS_SYNTH_ORDER"] = "file"
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=7)
os.environ["PS_PBSCN"] = "1"
sweep.run("case13659pegase", "PBRARVmodel", 2048, [8], what=7)
PY
cat $O/r2c11_gonly.log

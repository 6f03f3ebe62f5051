#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu -k "specialised or partial_what or code_paths or contingency or random_networks or batch_vs_oracle or layouts" > $O/r2c46_tests.log 2>&1; tail -3 $O/r2c46_tests.log
cat > /tmp/ab.py <<'P'
import os, subprocess, sys
ROOT = os.getcwd()
code = r'''
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case2869pegase", "PNPAPVmodel", 1024, [16], what=7, steps=8)
sweep.run("case2869pegase", "PNPAPVmodel", 1024, [16], what=3, steps=8)
sweep.run("case9241pegase", "PNPAPVmodel", 512, [16], what=7, steps=8)
sweep.run("case13659pegase", "PNRARVmodel", 512, [16], what=7, steps=8)
'''
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, PS_LIB=os.path.join(ROOT, "powersense.jl_b200", lib))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
        print(f"==== {lib} (rep {rep})"); print("\n".join(l for l in out.stdout.splitlines() if l.strip()), flush=True)
        if out.returncode: print(out.stderr[-800:])
P
python /tmp/ab.py libps_prev.so libpowersense_b200.so > $O/r2c46_ab.log 2>&1; grep -v "^   kern" $O/r2c46_ab.log; grep "kernels" $O/r2c46_ab.log | cut -c1-120

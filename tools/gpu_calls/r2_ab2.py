import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case13659pegase", "PBRARVmodel", 2048, [16], what=7)
sweep.run("case13659pegase", "PBRARVmodel", 2048, [16], what=1)
'''
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, PS_LIB=os.path.join(ROOT, "powersense.jl_b200", lib))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
        print(f"==== {lib} (rep {rep})")
        print("\n".join(l for l in out.stdout.splitlines() if l.strip()))
        if out.returncode:
            print(out.stderr[-1500:])

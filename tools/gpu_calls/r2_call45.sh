#!/bin/bash
for env in "" "PS_NO_BULK=1"; do echo "== $env"; env $env python - <<'P' 2>&1 | grep -v "^$"
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import sweep
sweep.run("case13659pegase", "PBRARVmodel", 4096, [16], what=3, steps=8)
sweep.run("case13659pegase", "PBRARVmodel", 4096, [16], what=4, steps=8)
P
done

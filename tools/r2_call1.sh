#!/bin/bash
# round 2, GPU call 1: baseline bench with callbacks + ncu sections for the gather-bound kernels
set -x
O=gpurun_out
python bench.py --callbacks --steps 10 > $O/r2c1_bench.json 2> $O/r2c1_bench.err
python tools/sweep.py ncu > $O/r2c1_plain_ncu.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pb_bus -s 4 -c 2 -o $O/r02_pb_bus python tools/sweep.py ncu > $O/r2c1_ncu1.log 2>&1
python tools/sweep.py ncu_g > $O/r2c1_plain_ncu_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'branch_g|pb_bus_g' -s 4 -c 2 -o $O/r02_gonly python tools/sweep.py ncu_g > $O/r2c1_ncu2.log 2>&1
python tools/sweep.py ncu_c3 > $O/r2c1_plain_ncu_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bus_warp|branch_direct' -s 4 -c 2 -o $O/r02_c3 python tools/sweep.py ncu_c3 > $O/r2c1_ncu3.log 2>&1
ls -la $O

#!/bin/bash
O=gpurun_out
for n in 3 6 9; do echo "-- stages $n"; PS_PBQ_STAGES=$n python tools/sweep.py ncu_g 2>&1 | grep -v "^$"; done
ncu --set full --clock-control none --import-source on -k regex:pb_bus_g_tma -c 1 -o $O/r02_pbq python tools/sweep.py ncu_g > $O/r2c25_ncu.log 2>&1; tail -3 $O/r2c25_ncu.log

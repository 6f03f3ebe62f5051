#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"bus_warp" -c 2 -o $O/r02_final_c3b python tools/sweep.py ncu_c3 > $O/r2c47_ncu.log 2>&1; tail -1 $O/r2c47_ncu.log

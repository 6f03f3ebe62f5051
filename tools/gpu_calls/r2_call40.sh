#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"branch_bulk|pb_bus" -c 6 -o $O/r02_final_pb python tools/sweep.py ncu > $O/r2c40_ncu1.log 2>&1; tail -1 $O/r2c40_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:"bus_warp|branch_direct" -c 4 -o $O/r02_final_c3 python tools/sweep.py ncu_c3 > $O/r2c40_ncu2.log 2>&1; tail -1 $O/r2c40_ncu2.log

#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pb_bus_g_gather -c 1 -o $O/r02_pbg python tools/sweep.py ncu_g > $O/r2c33_ncu.log 2>&1; tail -2 $O/r2c33_ncu.log

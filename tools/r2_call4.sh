#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident or code_paths" > $O/r2c5_newtests.log 2>&1
tail -3 $O/r2c5_newtests.log
python tools/sweep.py pbs > $O/r2c5_pbs_r4.log 2>&1
PS_LIB=$PWD/powersense.jl_b200/libps_r6.so python tools/sweep.py pbs > $O/r2c5_pbs_r6.log 2>&1
cat $O/r2c5_pbs_r4.log $O/r2c5_pbs_r6.log
python tools/sweep.py ncu > $O/r2c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pb_bus_g -s 2 -c 1 -o $O/r02_pbs3 python tools/sweep.py ncu > $O/r2c5_ncu.log 2>&1

#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_slp.py tests/test_gpu_opf.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/n1_sweep.py PNPAPVmodel case300 411 24 > $O/r2c9_n1_case300.log 2>&1; tail -6 $O/r2c9_n1_case300.log
PS_MAX_ITER=60 timeout 600 python tools/n1_sweep.py PNPAPVmodel case2869pegase:feasible 64 28 > $O/r2c9_n1_case2869.log 2>&1; tail -6 $O/r2c9_n1_case2869.log

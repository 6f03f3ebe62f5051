#!/bin/bash
O=gpurun_out
PS_MAX_ITER=12 timeout 1500 python tools/n1_sweep.py PNPAPVmodel case9241pegase:feasible 64 16 > $O/r2c31_n1_case9241.log 2>&1; echo rc=$?; grep -v "^  branch" $O/r2c31_n1_case9241.log | tail -12

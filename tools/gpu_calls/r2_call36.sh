#!/bin/bash
O=gpurun_out
timeout 1000 python tests/stress_paths.py 120 > $O/r2c36_stress.log 2>&1; echo rc=$?; tail -4 $O/r2c36_stress.log

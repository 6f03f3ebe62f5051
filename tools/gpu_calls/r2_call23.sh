#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident" > $O/r2c23_tests.log 2>&1; tail -3 $O/r2c23_tests.log
python tools/sweep.py side > $O/r2c23_side.log 2>&1; cat $O/r2c23_side.log
python tools/sweep.py side 2048 > $O/r2c23_side2048.log 2>&1; grep -v "^   kern" $O/r2c23_side2048.log
python tools/sweep.py latency > $O/r2c23_latency.log 2>&1; grep 13659 $O/r2c23_latency.log

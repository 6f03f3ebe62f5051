#!/bin/bash
O=gpurun_out
python tools/sweep.py side > $O/r2c22_side.log 2>&1; cat $O/r2c22_side.log | grep -v "^   kernels" 
python tools/latency_probe.py > $O/r2c22_latency.log 2>&1; cat $O/r2c22_latency.log

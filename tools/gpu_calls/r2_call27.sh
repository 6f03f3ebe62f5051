#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident" > $O/r2c27_tests.log 2>&1; tail -3 $O/r2c27_tests.log
python tools/r2_ab3.py libps_ahead0.so libps_ahead4.so libpowersense_b200.so libps_ahead16.so > $O/r2c27_ab.log 2>&1; cat $O/r2c27_ab.log

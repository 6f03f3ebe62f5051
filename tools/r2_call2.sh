#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scenario_resident or hessian_order or cache_keyed or rounded_powers or code_paths" > $O/r2c2_newtests.log 2>&1
tail -5 $O/r2c2_newtests.log
python bench.py --callbacks --steps 10 > $O/r2c2_bench.json 2> $O/r2c2_bench.err
tail -3 $O/r2c2_bench.err
python -m pytest tests -x -q -m gpu > $O/r2c2_tests.log 2>&1
tail -5 $O/r2c2_tests.log

#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r2c17_tests.log 2>&1; tail -4 $O/r2c17_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/r2c17_bench.json 2> $O/r2c17_bench.err; tail -2 $O/r2c17_bench.err

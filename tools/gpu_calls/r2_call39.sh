#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu > $O/r2c39_tests.log 2>&1; tail -3 $O/r2c39_tests.log
python bench.py > $O/r2c39_bench.json 2> $O/r2c39_bench.err; tail -1 $O/r2c39_bench.err | cut -c1-300

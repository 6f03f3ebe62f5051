#!/bin/bash
set -x
O=gpurun_out
python bench.py > $O/r2c7_bench.json 2> $O/r2c7_bench.err; echo rc=$?
tail -3 $O/r2c7_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c7_ref.json 2>> $O/r2c7_bench.err
python -m pytest tests/test_gpu_smoke_bench.py -x -q -m gpu 2>&1 | tail -5
python tools/sweep.py ncu > $O/r2c7_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:branch_bulk -s 2 -c 2 -o $O/r02_branch_bulk python tools/sweep.py ncu > $O/r2c7_ncu.log 2>&1

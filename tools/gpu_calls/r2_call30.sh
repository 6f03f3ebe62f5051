#!/bin/bash
O=gpurun_out
python bench.py > $O/r2c30_bench.json 2> $O/r2c30_bench.err; tail -1 $O/r2c30_bench.err | cut -c1-400
python bench.py --impl reference --steps 5 --warmup 1 > $O/r2c30_ref.json 2>/dev/null; cut -c1-300 $O/r2c30_ref.json
python tools/sweep.py all9 > $O/r2c30_all9.log 2>&1; grep -v "^$" $O/r2c30_all9.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench.csv python bench.py --no-e2e --no-cpu-baseline > $O/r2c30_ncu_bench.log 2>&1; tail -1 $O/r2c30_ncu_bench.log | cut -c1-200

#!/bin/bash
set -x
O=gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch_vs_oracle or specialised or code_paths or scenario_resident" 2>&1 | tail -4
python bench.py > $O/r2c12_bench.json 2> $O/r2c12_bench.err; tail -2 $O/r2c12_bench.err

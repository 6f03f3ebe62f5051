#!/bin/bash
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r2c28_tests.log 2>&1; tail -3 $O/r2c28_tests.log
python tools/latency_probe.py > $O/r2c28_latency.log 2>&1; cat $O/r2c28_latency.log
python tools/latency_probe.py case300 PNPAPVmodel > $O/r2c28_latency300.log 2>&1; cat $O/r2c28_latency300.log
python tools/latency_probe.py case14 PNPAPVmodel > $O/r2c28_latency14.log 2>&1; cat $O/r2c28_latency14.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c28_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/r2c28_memcheck.log

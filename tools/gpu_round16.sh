#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary16.txt; tail -n 5 gpurun_out/$name.log | cut -c1-1500; }
rm -f gpurun_out/summary16.txt
run tests_gpu_all python -m pytest tests/ -x -q -m gpu --durations=12
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_default python bench.py
run bench_reference python bench.py --impl reference

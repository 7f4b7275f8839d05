#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary_final.txt; tail -n 4 gpurun_out/$name.log | cut -c1-1200; }
rm -f gpurun_out/summary_final.txt
run tests_gpu_all python -m pytest tests/ -x -q -m gpu --durations=8
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench_default python bench.py
run bench_reference python bench.py --impl reference --steps 20 --warmup 3

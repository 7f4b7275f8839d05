#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary4.txt; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary4.txt
run tests_all python -m pytest tests -q -m gpu -x
run extract python tools/time_extract.py
run bench python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run stages python tools/tune_stages.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_retrieval3.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_launch3.log 2>&1
echo "launch list exit $?"

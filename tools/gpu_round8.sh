#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary8.txt; tail -n 5 gpurun_out/$name.log | cut -c1-600; }
rm -f gpurun_out/summary8.txt
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
  python bench.py --only train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_launch_train.log 2>&1
echo "train launch list exit $?"

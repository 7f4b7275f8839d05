#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary2.txt; tail -n 4 gpurun_out/$name.log | cut -c1-300; }
rm -f gpurun_out/summary2.txt
run retr_tests python -m pytest tests/test_retrieval_gpu.py -q -m gpu
run extract python tools/time_extract.py
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_retrieval2.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_launch2.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract.csv \
  python tools/time_extract.py > gpurun_out/prof_launch_extract.log 2>&1
echo "extract launch list exit $?"

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary7.txt; tail -n 8 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary7.txt
run tests_all python -m pytest tests -q -m gpu
run extract python tools/time_extract.py 64,256
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract5.csv \
  python tools/time_extract.py 256 1 > gpurun_out/prof_launch_extract5.log 2>&1
echo "extract launch list exit $?"

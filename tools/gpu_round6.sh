#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary6.txt; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary6.txt
run tests_all python -m pytest tests -q -m gpu
run extract python tools/time_extract.py 64,256
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract4.csv \
  python tools/time_extract.py 256 1 > gpurun_out/prof_launch_extract4.log 2>&1
echo "extract launch list exit $?"
# stage-3 dwconv + fc1 + fc2 with source counters (3rd forward, block ~15 of stage 3)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv7_ln_kernel|gemm_tn_kernel" -s 296 -c 3 \
  -o gpurun_out/prof_extract_s3b -f python tools/time_extract.py 256 1 > gpurun_out/prof_extract_s3b.log 2>&1
echo "s3 capture exit $?"

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary26.txt; tail -n 8 gpurun_out/$name.log | cut -c1-300; }
rm -f gpurun_out/summary26.txt
run tests_vit_ret python -m pytest tests/test_vit_gpu.py tests/test_retrieval_gpu.py -q -k "training or l2_variant"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1100 --launch-count 520 --csv --log-file gpurun_out/launches_vit_train.csv \
  python tools/time_vit_train.py 128 1 > gpurun_out/prof_launch_vit_train.log 2>&1
echo "launch list exit $?"

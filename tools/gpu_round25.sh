#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary25.txt; tail -n 14 gpurun_out/$name.log | cut -c1-300; }
rm -f gpurun_out/summary25.txt
run tests_vit python -m pytest tests/test_vit_gpu.py -q -s -k "training"
run time_vit_train python tools/time_vit_train.py 128 5

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary20.txt; tail -n 25 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary20.txt
run tests_vit python -m pytest tests/test_vit_gpu.py -x -q
run time_vit python tools/time_vit.py vit_base_patch16_224 256 10

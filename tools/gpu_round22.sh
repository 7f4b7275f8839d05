#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary22.txt; tail -n 12 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary22.txt
run tests_vit_heads python -m pytest tests/test_vit_gpu.py tests/test_heads_gpu.py -x -q
run time_vit python tools/time_vit.py vit_base_patch16_224 256 10
run time_vit_l python tools/time_vit.py vit_large_patch14_clip_336 64 5

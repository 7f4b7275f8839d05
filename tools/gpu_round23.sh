#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary23.txt; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary23.txt
run tests_vit python -m pytest tests/test_vit_gpu.py -x -q
for w in 16 8 4; do
  VDK_ATT_WARPS=$w timeout 120 python tools/time_vit.py vit_base_patch16_224 256 10 2>&1 | tail -1 | sed "s/^/warps=$w /"
done
VDK_ATT_WARPS=8 timeout 120 python tools/time_vit.py vit_large_patch14_clip_336 64 5 2>&1 | tail -1 | sed "s/^/warps=8 /"

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary3.txt; tail -n 12 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary3.txt
run debug_overflow python tools/debug_overflow.py
run heads_optim python -m pytest tests/test_heads_gpu.py tests/test_optim_gpu.py -q -m gpu
# source-level profile of the early (slow-path heavy) ranges: launches 2 and 3 of score_filter<false> in a step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_kernel -s 13 -c 2 \
  -o gpurun_out/prof_score_early -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_early.log 2>&1
echo "early profile exit $?"

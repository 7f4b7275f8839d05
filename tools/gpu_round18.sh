#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary18.txt; tail -n 12 gpurun_out/$name.log | cut -c1-600; }
rm -f gpurun_out/summary18.txt
run tests_train python -m pytest tests/test_convnext_train_gpu.py tests/test_gemm_gpu.py -x -q -s -k "not subprocess" --durations=3
timeout 300 python tools/prof_train_kernels.py 2 128 10 fc > gpurun_out/prof_fc4_st2.log 2>&1; cat gpurun_out/prof_fc4_st2.log
run bench_train python bench.py --only train --steps 10 --warmup 3 --no-cpu-baseline

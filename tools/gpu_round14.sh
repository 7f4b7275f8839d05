#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary14.txt; tail -n 4 gpurun_out/$name.log | cut -c1-600; }
rm -f gpurun_out/summary14.txt
run tests_gemm python -m pytest tests/test_gemm_gpu.py tests/test_convnext_train_gpu.py -x -q
VDK_GEMM_PAIR=2 run tests_gemm_pair python -m pytest tests/test_gemm_gpu.py -x -q
for st in 2 0; do
  timeout 300 python tools/prof_train_kernels.py $st 128 10 fc > gpurun_out/prof_fc2_st${st}.log 2>&1
done
cat gpurun_out/prof_fc2_st*.log
run tests_all python -m pytest tests/test_convnext_gpu.py tests/test_heads_gpu.py tests/test_cbir_gpu.py -x -q
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary13.txt; tail -n 6 gpurun_out/$name.log | cut -c1-700; }
rm -f gpurun_out/summary13.txt
run tests_dw python -m pytest tests/test_convnext_train_gpu.py -x -q -k "dwconv or layernorm or batchnorm"
run tests_dwf python -m pytest tests/test_convnext_gpu.py -x -q -k "dwconv"
for st in 0 2; do
  timeout 300 python tools/prof_train_kernels.py $st 128 10 dwconv > gpurun_out/prof_dw3_st${st}.log 2>&1
done
cat gpurun_out/prof_dw3_st*.log
VDK_GEMM_PAIR=1 run tests_gemm_pair python -m pytest tests/test_gemm_gpu.py -x -q
for st in 2 0 3; do
  VDK_GEMM_PAIR=0 timeout 300 python tools/prof_train_kernels.py $st 128 10 fc > gpurun_out/prof_fc_st${st}_pair0.log 2>&1
  VDK_GEMM_PAIR=1 timeout 300 python tools/prof_train_kernels.py $st 128 10 fc > gpurun_out/prof_fc_st${st}_pair1.log 2>&1
done
tail -n 20 gpurun_out/prof_fc_st*.log
run tests_all python -m pytest tests/test_convnext_gpu.py tests/test_convnext_train_gpu.py tests/test_heads_gpu.py -x -q
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline

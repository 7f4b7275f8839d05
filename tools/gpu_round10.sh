#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary10.txt; tail -n 6 gpurun_out/$name.log | cut -c1-700; }
rm -f gpurun_out/summary10.txt
run tests_gemm python -m pytest tests/test_gemm_gpu.py tests/test_convnext_train_gpu.py tests/test_convnext_gpu.py -x -q --durations=5
for st in 2 0; do
  VDK_GEMM_AUXPIPE=0 timeout 300 python tools/prof_train_kernels.py $st 128 10 fc > gpurun_out/prof_k_st${st}_pipe0.log 2>&1
  timeout 300 python tools/prof_train_kernels.py $st 128 10 > gpurun_out/prof_k_st${st}_pipe7.log 2>&1
done
cat gpurun_out/prof_k_st*.log
timeout 600 ncu --set full --import-source on --clock-control none -o gpurun_out/train_kernels_st2 -f \
  python tools/prof_train_kernels.py 2 128 1 > gpurun_out/ncu_train_kernels.log 2>&1
echo "ncu exit $?"
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline

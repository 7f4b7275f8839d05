#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary11.txt; tail -n 6 gpurun_out/$name.log | cut -c1-700; }
rm -f gpurun_out/summary11.txt
run tests_dw python -m pytest tests/test_convnext_gpu.py tests/test_convnext_train_gpu.py tests/test_cbir_gpu.py -x -q --durations=5
for st in 0 1 2 3; do
  VDK_DWCONV_CHUNKED=0 timeout 300 python tools/prof_train_kernels.py $st 128 10 dwconv7_ln,bwd-data > gpurun_out/prof_dw_st${st}_old.log 2>&1
  timeout 300 python tools/prof_train_kernels.py $st 128 10 dwconv > gpurun_out/prof_dw_st${st}_new.log 2>&1
done
cat gpurun_out/prof_dw_st*.log
timeout 600 ncu --set full --import-source on --clock-control none -o gpurun_out/train_kernels_st2 -f \
  python tools/prof_train_kernels.py 2 128 1 "GELU only,gelu',dwconv7_ln,wgrad" > gpurun_out/ncu_train_kernels.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary12.txt; tail -n 6 gpurun_out/$name.log | cut -c1-700; }
rm -f gpurun_out/summary12.txt
run tests_dw python -m pytest tests/test_convnext_gpu.py tests/test_convnext_train_gpu.py -x -q --durations=3
for st in 0 2; do
  timeout 300 python tools/prof_train_kernels.py $st 128 10 dwconv > gpurun_out/prof_dw2_st${st}.log 2>&1
done
cat gpurun_out/prof_dw2_st*.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_tn|dwconv7" -o /tmp/train_kernels_st2 -f \
  python tools/prof_train_kernels.py 2 128 1 "GELU only,gelu',dwconv7_ln,wgrad" > gpurun_out/ncu_train_kernels.log 2>&1
echo "ncu exit $?"; ls -la /tmp/*.ncu-rep
ncu -i /tmp/train_kernels_st2.ncu-rep --page raw --csv > gpurun_out/ncu_train_raw.csv 2>/dev/null
ncu -i /tmp/train_kernels_st2.ncu-rep --page details --csv > gpurun_out/ncu_train_details.csv 2>/dev/null
ncu -i /tmp/train_kernels_st2.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_train_source_sass.csv 2>/dev/null
sz=$(stat -c %s /tmp/train_kernels_st2.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/train_kernels_st2.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -8
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline

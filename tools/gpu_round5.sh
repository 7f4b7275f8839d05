#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 500 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary5.txt; tail -n 6 gpurun_out/$name.log | cut -c1-400; }
rm -f gpurun_out/summary5.txt
run tests_cg python -m pytest tests/test_convnext_gpu.py tests/test_gemm_gpu.py tests/test_retrieval_gpu.py tests/test_heads_gpu.py -q -m gpu
run extract python tools/time_extract.py 64,256
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract3.csv \
  python tools/time_extract.py 256 1 > gpurun_out/prof_launch_extract3.log 2>&1
echo "extract launch list exit $?"
run bench python bench.py --steps 10 --warmup 3

#!/bin/bash
# First GPU bring-up: each group in its own process so a trapped kernel cannot poison later groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/$name.log; }
run gemm_tiny python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "plain_fp32_out and 128-128-64"
run gemm_all python -m pytest tests/test_gemm_gpu.py -q -m gpu
run retr_prepare python -m pytest tests/test_retrieval_gpu.py -q -m gpu -k "rows_prepare or error_within"
run retr_single python -m pytest tests/test_retrieval_gpu.py -q -m gpu -k "single_range"
run retr_rest python -m pytest tests/test_retrieval_gpu.py -q -m gpu -k "not single_range and not rows_prepare and not error_within and not full_size"
run retr_full python -m pytest tests/test_retrieval_gpu.py -q -m gpu -k "full_size"
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 5 --warmup 3
cat gpurun_out/summary.txt

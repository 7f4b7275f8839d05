#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary17.txt; tail -n 4 gpurun_out/$name.log | cut -c1-1200; }
rm -f gpurun_out/summary17.txt
run tests_train python -m pytest tests/test_convnext_train_gpu.py tests/test_engine_gpu.py tests/test_optim_gpu.py -x -q
run bench_train python bench.py --only train --steps 10 --warmup 3 --no-cpu-baseline
run multi_gpu_check python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
run bench_2gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline

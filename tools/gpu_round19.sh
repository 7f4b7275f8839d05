#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary19.txt; tail -n 6 gpurun_out/$name.log | cut -c1-900; }
rm -f gpurun_out/summary19.txt
run tests_train python -m pytest tests/test_convnext_train_gpu.py -x -q -k "ranges or gradients or fused"
run multi_gpu_check python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
run bench_2gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --only train --steps 10 --warmup 3 --no-cpu-baseline

#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary_final3.txt; tail -n 3 gpurun_out/$name.log | cut -c1-500; }
rm -f gpurun_out/summary_final3.txt
run tests_vit python -m pytest tests/test_vit_gpu.py -x -q -k "ranges or toy_training or circleloss"
run bench_2gpu_train python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --only train --steps 10 --warmup 3

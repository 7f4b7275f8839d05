#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary_final2.txt; tail -n 3 gpurun_out/$name.log | cut -c1-600; }
rm -f gpurun_out/summary_final2.txt
run multi_gpu_check python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/check_multi_gpu.py
run bench_2gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3
run bench_ref_2gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 3 --warmup 1

#!/bin/bash
mkdir -p gpurun_out
for st in 2 0; do
  timeout 300 python tools/prof_train_kernels.py $st 128 10 "fc1 fwd" > gpurun_out/prof_fc3_st${st}.log 2>&1
done
cat gpurun_out/prof_fc3_st*.log

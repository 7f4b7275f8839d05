#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary9.txt; tail -n 6 gpurun_out/$name.log | cut -c1-700; }
rm -f gpurun_out/summary9.txt
run tests_train python -m pytest tests/test_convnext_train_gpu.py tests/test_heads_gpu.py tests/test_engine_gpu.py -x -q --durations=8
run bench_train python bench.py --only train --steps 5 --warmup 3 --no-cpu-baseline
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2330 --launch-count 800 --csv --log-file gpurun_out/launches_train2.csv \
  python bench.py --only train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_launch_train2.log 2>&1
echo "train launch list exit $?"

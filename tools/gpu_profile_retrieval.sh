#!/bin/bash
# Launch list (per-kernel device time) + one full ncu capture of the dominant retrieval kernel.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_retrieval.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_launch.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_filter_kernel -s 6 -c 2 \
  -o gpurun_out/prof_score -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out

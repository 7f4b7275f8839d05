#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract2.csv \
  python tools/time_extract.py 256 1 > gpurun_out/prof_launch_extract2.log 2>&1
echo "extract launch list exit $?"
# full capture: one stage-1 dwconv, stage-1 fc1 (GELU), stage-3 dwconv, stage-3 fc1/fc2 of the 3rd (timed) forward
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv7_ln_kernel|gemm_tn_kernel" -s 240 -c 4 \
  -o gpurun_out/prof_extract_s1 -f python tools/time_extract.py 256 1 > gpurun_out/prof_extract_s1.log 2>&1
echo "s1 capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv7_ln_kernel|gemm_tn_kernel" -s 296 -c 3 \
  -o gpurun_out/prof_extract_s3 -f python tools/time_extract.py 256 1 > gpurun_out/prof_extract_s3.log 2>&1
echo "s3 capture exit $?"

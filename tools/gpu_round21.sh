#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary21.txt; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
rm -f gpurun_out/summary21.txt
# (a) DRAM traffic of the dominant GEMM launches (stage-3 MLP of the batch-256 forward)
timeout 300 ncu --set full --clock-control none -k regex:gemm_tn --launch-skip 17 --launch-count 2 -o /tmp/traffic_gemm -f \
  python tools/time_extract.py 256 1 > gpurun_out/ncu_traffic_gemm.log 2>&1
ncu -i /tmp/traffic_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_traffic_gemm_raw.csv 2>/dev/null
# (b) DRAM traffic of the retrieval kernel on the last gallery range
timeout 300 ncu --set full --clock-control none -k regex:score_filter --launch-skip 3 --launch-count 1 -o /tmp/traffic_ret -f \
  python bench.py --only retrieval --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_traffic_ret.log 2>&1
ncu -i /tmp/traffic_ret.ncu-rep --page raw --csv > gpurun_out/ncu_traffic_ret_raw.csv 2>/dev/null
# (c) launch list of one ViT-B/16 forward
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"vdk|gemm_tn|attention" --csv --log-file gpurun_out/launches_vit.csv \
  python tools/time_vit.py vit_base_patch16_224 256 1 > gpurun_out/prof_launch_vit.log 2>&1
echo "ncu done"; ls -la gpurun_out/ncu_traffic_* gpurun_out/launches_vit.csv
run bench_extract python bench.py --only extract --steps 10 --warmup 3 --no-cpu-baseline

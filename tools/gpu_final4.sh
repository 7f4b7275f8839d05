#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/ -x -q -m gpu > gpurun_out/tests_head.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/tests_head.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_head.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/smoke_head.log | cut -c1-300

#!/bin/bash
# Round 2, seventh GPU call: any-hit shadow rays — parity, then A/B on the plastic configurations.
set -u
out=gpurun_out/r2c7; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_parity.log
timeout 1500 python -m pytest tests/test_fullsize_gpu.py tests/test_shipped_scenes_gpu.py -m gpu -q > $out/pytest_rest.log 2>&1; echo "fullsize+shipped rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_rest.log
FJ_SWEEP_WORKLOAD=config4 bash tools/sweep.sh "FJGPU_ANYHIT=0" "FJGPU_ANYHIT=1" > $out/sweep_anyhit.log 2>&1
FJ_SWEEP_WORKLOAD=config2 bash tools/sweep.sh "FJGPU_ANYHIT=0" "FJGPU_ANYHIT=1" >> $out/sweep_anyhit.log 2>&1
cat $out/sweep_anyhit.log

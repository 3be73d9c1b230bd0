#!/bin/bash
# Round 2, twenty-second GPU call: FP32 triangle packets read as 3 x 16 B (no parity selects) against the 32 + 16 split.
set -u
out=gpurun_out/r2c22; mkdir -p $out
FJGPU_TRI3=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trace_closest or variants_bit_exact or frame_matches" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log
tail -n 2 $out/pytest.log
bash tools/sweep.sh "FJGPU_TRI3=0" "FJGPU_TRI3=1" "FJGPU_TRI3=0" "FJGPU_TRI3=1" > $out/sweep.log 2>&1
FJ_SWEEP_WORKLOAD=config4 bash tools/sweep.sh "FJGPU_TRI3=0" "FJGPU_TRI3=1" >> $out/sweep.log 2>&1
cat $out/sweep.log

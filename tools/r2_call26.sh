#!/bin/bash
# Round 2, twenty-sixth GPU call: leaf phases of whole rounds only (FJGPU_B1_WHOLE=1) — parity, then A/B.
set -u
out=gpurun_out/r2c26; mkdir -p $out
FJGPU_B1_WHOLE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trace_closest or variants_bit_exact or frame_matches or variants_same_frame" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log
tail -n 2 $out/pytest.log
bash tools/sweep.sh "FJGPU_B1_WHOLE=0" "FJGPU_B1_WHOLE=1" "FJGPU_B1_WHOLE=1 FJGPU_B1_MIN=24" "FJGPU_B1_WHOLE=1 FJGPU_B1_MIN=28" "FJGPU_B1_WHOLE=1 FJGPU_B1_MIN=16" > $out/sweep.log 2>&1
cat $out/sweep.log

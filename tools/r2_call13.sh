#!/bin/bash
# Round 2, thirteenth GPU call: k_extend_ring with the cheap transitions inline and the heavy phases (leaf tests, instance
# entries) gated by the number of waiting lanes — parity of every variant, then the thresholds.
set -u
out=gpurun_out/r2c13; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 5 $out/pytest_parity.log
E="FJGPU_EXTEND=3 FJGPU_RING=0"
bash tools/sweep.sh "FJGPU_EXTEND=2" "$E" "$E FJGPU_B2_MIN=8" "$E FJGPU_B2_MIN=12" "$E FJGPU_B1_MIN=24" "$E FJGPU_B1_MIN=32" "$E FJGPU_B1_MIN=24 FJGPU_B2_MIN=8" \
  "$E FJGPU_B1_MIN=28 FJGPU_B2_MIN=12" "$E FJGPU_B1_MIN=48 FJGPU_B2_MIN=16" "$E FJGPU_B1_MIN=24 FJGPU_B2_MIN=8 FJGPU_PHASE_A_MIN=12" "$E FJGPU_B1_MIN=24 FJGPU_B2_MIN=8 FJGPU_PHASE_A_MIN=20" \
  "$E FJGPU_B1_MIN=24 FJGPU_B2_MIN=8 FJGPU_REFILL=8" "FJGPU_EXTEND=3" "FJGPU_EXTEND=3 FJGPU_B1_MIN=24 FJGPU_B2_MIN=8" "FJGPU_EXTEND=3 FJGPU_B1_MIN=24 FJGPU_B2_MIN=8 FJGPU_REFILL=16" > $out/sweep.log 2>&1
cat $out/sweep.log

#!/bin/bash
# Round 2, tenth GPU call: k_extend_ring after the L1-wavefront diet (16-B record loads, target in the state word, one 32-B hit
# store) and with packed FMAs — parity, then A/B.
set -u
out=gpurun_out/r2c10; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 5 $out/pytest_parity.log
bash tools/sweep.sh "FJGPU_EXTEND=2" "FJGPU_EXTEND=3" "FJGPU_EXTEND=3 FJGPU_FMA2=0" "FJGPU_EXTEND=3 FJGPU_REFILL=24" "FJGPU_EXTEND=3 FJGPU_REFILL=32" \
  "FJGPU_EXTEND=3 FJGPU_REFILL=32 FJGPU_PHASE_A_MIN=14" "FJGPU_EXTEND=3 FJGPU_REFILL=32 FJGPU_PHASE_A_MIN=18" "FJGPU_EXTEND=3 FJGPU_REFILL=32 FJGPU_STACK_SMEM=12" \
  "FJGPU_EXTEND=3 FJGPU_REFILL=32 FJGPU_EXTEND_MINBLOCKS=8" "FJGPU_EXTEND=3 FJGPU_REFILL=32 FJGPU_EXTEND_MINBLOCKS=6" > $out/sweep.log 2>&1
cat $out/sweep.log

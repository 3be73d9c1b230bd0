#!/bin/bash
# Round 2, ninth GPU call: k_extend_ring (per-warp ring of prepared rays) — parity of every variant, then A/B and thresholds.
set -u
out=gpurun_out/r2c9; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 5 $out/pytest_parity.log
bash tools/sweep.sh "FJGPU_EXTEND=2" "FJGPU_EXTEND=3" "FJGPU_EXTEND=3 FJGPU_REFILL=8" "FJGPU_EXTEND=3 FJGPU_REFILL=24" "FJGPU_EXTEND=3 FJGPU_REFILL=32" \
  "FJGPU_EXTEND=3 FJGPU_STACK_SMEM=12" "FJGPU_EXTEND=3 FJGPU_PHASE_A_MIN=12" "FJGPU_EXTEND=3 FJGPU_PHASE_A_MIN=20" "FJGPU_EXTEND=3 FJGPU_PHASE_A_MIN=24" \
  "FJGPU_EXTEND=3 FJGPU_EXTEND_MINBLOCKS=6" "FJGPU_EXTEND=3 FJGPU_EXTEND_MINBLOCKS=8 FJGPU_REFILL=8" "FJGPU_EXTEND=3 FJGPU_CARVEOUT_PCT=60" "FJGPU_EXTEND=3 FJGPU_CARVEOUT_PCT=80" > $out/sweep.log 2>&1
cat $out/sweep.log

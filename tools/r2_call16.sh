#!/bin/bash
# Round 2, sixteenth GPU call: occupancy / threshold re-sweep of the default kernel after gating.
set -u
out=gpurun_out/r2c16; mkdir -p $out
bash tools/sweep.sh "FJGPU_EXTEND=3" "FJGPU_EXTEND_MINBLOCKS=8" "FJGPU_EXTEND_MINBLOCKS=8 FJGPU_REFILL=4" "FJGPU_EXTEND_MINBLOCKS=6" "FJGPU_STACK_SMEM=8" "FJGPU_REFILL=4" "FJGPU_REFILL=6" "FJGPU_REFILL=10" \
  "FJGPU_B1_MIN=20" "FJGPU_B1_MIN=28" "FJGPU_B2_MIN=6" "FJGPU_B2_MIN=10" "FJGPU_PHASE_A_MIN=14" "FJGPU_PHASE_A_MIN=18" "FJGPU_FMA2=0 FJGPU_STACK_SMEM=8 FJGPU_RING=1" > $out/sweep.log 2>&1
cat $out/sweep.log

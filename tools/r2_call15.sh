#!/bin/bash
# Round 2, fifteenth GPU call: k_shade with the next record group prefetched into L2 — parity of the frames, then A/B.
set -u
out=gpurun_out/r2c15; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "frame or chunk or batch" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest.log
bash tools/sweep.sh "FJGPU_SHADE_PREFETCH=0" "FJGPU_SHADE_PREFETCH=1" "FJGPU_SHADE_PREFETCH=0" "FJGPU_SHADE_PREFETCH=1" "FJGPU_SHADE_PREFETCH=1 FJGPU_SHADE_MINBLOCKS=6" "FJGPU_SHADE_PREFETCH=1 FJGPU_SHADE_CTAS=4" > $out/sweep.log 2>&1
FJ_SWEEP_WORKLOAD=config4 bash tools/sweep.sh "FJGPU_SHADE_PREFETCH=0" "FJGPU_SHADE_PREFETCH=1" >> $out/sweep.log 2>&1
cat $out/sweep.log

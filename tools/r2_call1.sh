#!/bin/bash
# Round 2, first GPU call: parity of everything that changed (shared-memory traversal stack, staged tree top, moving
# triangles, queue growth in later batches, full-size frames), then the A/B sweeps of the new k_extend2 knobs and the
# ncu --set full captures of the kernels that ship.  Everything lands in gpurun_out/r2c1/.
set -u
out=gpurun_out/r2c1; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
nproc > $out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
bash tools/sweep.sh "FJGPU_STACK_SMEM=12" "FJGPU_STACK_SMEM=8" "FJGPU_STACK_SMEM=16" \
  "FJGPU_EXTEND_MINBLOCKS=6" "FJGPU_EXTEND_MINBLOCKS=8" \
  "FJGPU_TOP_NODES=21" "FJGPU_TOP_NODES=85" "FJGPU_TOP_NODES=341 FJGPU_EXTEND_MINBLOCKS=6" \
  "FJGPU_STACK_SMEM=8 FJGPU_CARVEOUT_PCT=50" "FJGPU_STACK_SMEM=12 FJGPU_CARVEOUT_PCT=55" > $out/sweep.log 2>&1
cat $out/sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_extend2 -s 2 -c 2 -o $out/k_extend2_r2 \
  python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_extend.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 2 -o $out/k_shade_r2 \
  python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_shade.log 2>&1
python profiles/ncu_summary.py $out/k_extend2_r2.ncu-rep > $out/k_extend2_r2_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_shade_r2.ncu-rep > $out/k_shade_r2_ncu_full.txt 2>&1
ls -la $out

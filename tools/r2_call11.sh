#!/bin/bash
# Round 2, eleventh GPU call: why did neither fewer instructions nor fewer L1 wavefronts move k_extend?  ncu --set full of the ring
# kernel (ring and direct refill) on the full-size north star, plus the A/B lines.
set -u
out=gpurun_out/r2c11; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants" > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_parity.log
bash tools/sweep.sh "FJGPU_EXTEND=2" "FJGPU_EXTEND=3" "FJGPU_EXTEND=3 FJGPU_RING=0" "FJGPU_EXTEND=3 FJGPU_RING=0 FJGPU_STACK_SMEM=8" "FJGPU_EXTEND=3 FJGPU_RING=0 FJGPU_REFILL=8" > $out/sweep.log 2>&1
cat $out/sweep.log
FJGPU_EXTEND=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 4 -c 1 -o $out/k_ring \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_ring.log 2>&1
FJGPU_EXTEND=3 FJGPU_RING=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 4 -c 1 -o $out/k_direct \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_direct.log 2>&1
python profiles/ncu_summary.py $out/k_ring.ncu-rep > $out/k_ring_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_direct.ncu-rep > $out/k_direct_ncu_full.txt 2>&1
ls -la $out

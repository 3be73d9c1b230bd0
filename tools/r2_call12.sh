#!/bin/bash
# Round 2, twelfth GPU call: ncu --set full of a BOUNCE launch (incoherent rays: two thirds of k_extend's time) for the ring
# kernel, its direct-refill variant and k_extend2.
set -u
out=gpurun_out/r2c12; mkdir -p $out
FJGPU_EXTEND=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 5 -c 1 -o $out/k_ring_b \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_ring.log 2>&1
FJGPU_EXTEND=3 FJGPU_RING=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 5 -c 1 -o $out/k_direct_b \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_direct.log 2>&1
python profiles/ncu_summary.py $out/k_ring_b.ncu-rep > $out/k_ring_b_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_direct_b.ncu-rep > $out/k_direct_b_ncu_full.txt 2>&1
ls -la $out

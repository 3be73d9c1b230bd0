#!/bin/bash
# Round 2, nineteenth GPU call: cheap transitions once per outer iteration against inline, next-node prefetch; the bench's
# clock sampler started before the warm-up with a 1-s period (the driver's command twice: is the gap to the kernels gone?).
set -u
out=gpurun_out/r2c19; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "parity rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_parity.log
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('bench 20 steps: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  gap %.1f ms  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['ms_per_step']-sum(k.values()), d['e2e']['value'], d['clocks']))" | tee -a $out/bench20.log
done
bash tools/sweep.sh "FJGPU_TRANSIT_INLINE=1" "FJGPU_TRANSIT_INLINE=0" "FJGPU_TRANSIT_INLINE=0 FJGPU_B1_MIN=20 FJGPU_B2_MIN=6" "FJGPU_TRANSIT_INLINE=0 FJGPU_B1_MIN=28 FJGPU_B2_MIN=10" "FJGPU_TRANSIT_INLINE=0 FJGPU_PHASE_A_MIN=14" "FJGPU_TRANSIT_INLINE=0 FJGPU_PHASE_A_MIN=18" \
  "FJGPU_NODE_PREFETCH=1" "FJGPU_NODE_PREFETCH=1 FJGPU_TRANSIT_INLINE=0" "FJGPU_TRANSIT_INLINE=0 FJGPU_REFILL=12" > $out/sweep.log 2>&1
cat $out/sweep.log

#!/bin/bash
# Round 2, thirty-second GPU call: the same with the batch budget cached (no cudaMemGetInfo per frame); batching tests first.
set -u
out=gpurun_out/r2c32; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch or overflow or budget or regrow" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log; tail -n 2 $out/pytest.log
for lms in default default default; do
  if [ $lms = default ]; then unset FJ_CLOCK_LMS; else export FJ_CLOCK_LMS=$lms; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('lms $lms: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  gap %.1f ms  per-step wall %s  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['ms_per_step']-sum(k.values()), {a: round(b,1) for a,b in d['step_wall_ms'].items()}, d['clocks']))" | tee -a $out/gap.log
done

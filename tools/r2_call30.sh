#!/bin/bash
# Round 2, thirtieth GPU call: where do 3-14 ms per step go between the bench's step time and its kernels?  Per-step wall times,
# with the clock sampler and without it (FJ_CLOCK_LMS=60000: no poll inside the region).
set -u
out=gpurun_out/r2c30; mkdir -p $out
for lms in default default 60000 60000; do
  if [ $lms = default ]; then unset FJ_CLOCK_LMS; else export FJ_CLOCK_LMS=$lms; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('lms $lms: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  gap %.1f ms  per-step wall %s  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['ms_per_step']-sum(k.values()), {a: round(b,1) for a,b in d['step_wall_ms'].items()}, d['clocks']))" | tee -a $out/gap.log
done

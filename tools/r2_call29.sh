#!/bin/bash
# Round 2, twenty-ninth GPU call: the bench's clock sampler at about three polls per timed region — the driver's command three
# times (gap between the step time and the three kernels), one short run.
set -u
out=gpurun_out/r2c29; mkdir -p $out
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('bench 20 steps: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  gap %.1f ms  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['ms_per_step']-sum(k.values()), d['e2e']['value'], d['clocks']))" | tee -a $out/bench20.log
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('bench 3 steps: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['clocks']))" | tee -a $out/bench20.log

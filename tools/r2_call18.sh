#!/bin/bash
# Round 2, eighteenth GPU call: what are the 15 ms per step between the bench's step time and its three kernels on 20-step runs?
# (nvidia-smi polling period), then config 5 with the kernel that ships.
set -u
out=gpurun_out/r2c18; mkdir -p $out
for lms in 200 1000 5000 200; do
  FJ_CLOCK_LMS=$lms timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('lms $lms: %.1f Mrays/s  step %.1f ms  kernels %.1f ms  gap %.1f ms  e2e %.1f  clocks %s' % (d['value'], d['ms_per_step'], sum(k.values()), d['ms_per_step']-sum(k.values()), d['e2e']['value'], d['clocks']))" | tee -a $out/clock_period.log
done
FJGPU_BUILD=device FJ_PARITY_TILES=2 timeout 1500 python bench.py --workload config5 --steps 2 --warmup 3 > $out/bench_config5.json 2> $out/bench_config5.err
tail -c 500 $out/bench_config5.json

#!/bin/bash
# usage: tools/sweep.sh "VAR=val VAR2=val" "VAR=val" ...   — one north-star bench line per configuration
# (FJ_SWEEP_WORKLOAD selects another workload)
for cfg in "$@"; do
  echo -n "$cfg => "
  env $cfg timeout 900 python bench.py --workload ${FJ_SWEEP_WORKLOAD:-north_star} --steps 2 --warmup 2 --no-cpu-baseline --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
f=d['frame']
print('%.1f Mrays/s  %.1f ms/frame  kernels %s  e2e %.1f' % (d['value'], d['ms_per_step'], {k: round(v,1) for k,v in d['kernel_ms_per_step'].items()}, d['e2e']['value']), ' steps/ray %.1f tris/ray %.1f pairs/leaf-phase %.1f rounds %.2f' % (f['node_steps_per_ray'], f['tri_tests_per_ray'], f['tri_pairs_per_leaf_phase'], f['rounds_per_leaf_phase']))"
done

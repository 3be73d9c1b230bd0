#!/bin/bash
# Round 2, thirty-fourth GPU call (8 GPUs): the driver's command at N = 8 on the final repository state.
set -u
out=gpurun_out/r2c34; mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_8gpu.json 2> $out/bench_8gpu.err
tail -1 $out/bench_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d: value %.1f Mrays/s  %.2f ms/step  e2e %.1f (%.2f ms)  kernels %s  step wall %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0), {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items()}, d.get('step_wall_ms')))"

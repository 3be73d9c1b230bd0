#!/bin/bash
# Round 2, twenty-first GPU call (8 GPUs): the multi-GPU paths with the kernels that ship — the driver's command at N = 8, 4, 2
# under torchrun (value, e2e), fjgpu_render_frame_multi / FJ_GPU_COUNT in one process.
set -u
out=gpurun_out/r2c21; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --steps 20 --warmup 5 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
  tail -1 $out/bench_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d: value %.1f Mrays/s  %.2f ms/step  e2e %.1f (%.2f ms)  kernels %s  clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('ms_per_step', 0), {k: round(v,1) for k,v in d.get('kernel_ms_per_step',{}).items()}, d.get('clocks')))" | tee -a $out/summary.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or assemble" > $out/pytest_parity_multi.log 2>&1; echo "parity-multi rc=$?" | tee -a $out/summary.log
timeout 600 python -m pytest tests/test_host_mirror.py -m gpu -q -k "several_gpus" > $out/pytest_fjscene_multi.log 2>&1; echo "fjscene-multi rc=$?" | tee -a $out/summary.log
tail -2 $out/pytest_parity_multi.log $out/pytest_fjscene_multi.log

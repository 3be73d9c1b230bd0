#!/bin/bash
# Round 2, sixth GPU call (2 GPUs): the multi-GPU paths — bench at N = 2 under torchrun (value, e2e with the device un-permute +
# one pinned copy), fjgpu_render_frame_multi / FJ_GPU_COUNT in one process — plus the files that failed before and two A/Bs.
set -u
out=gpurun_out/r2c6; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi or assemble or variants_same_frame or overflow" > $out/pytest_parity_multi.log 2>&1; echo "parity-multi rc=$?" | tee -a $out/summary.log
timeout 900 python -m pytest tests/test_host_mirror.py -m gpu -q -k "several_gpus" > $out/pytest_fjscene_multi.log 2>&1; echo "fjscene-multi rc=$?" | tee -a $out/summary.log
timeout 1500 python -m pytest tests/test_shipped_scenes_gpu.py -m gpu -q > $out/pytest_shipped.log 2>&1; echo "shipped rc=$?" | tee -a $out/summary.log
tail -3 $out/pytest_parity_multi.log $out/pytest_fjscene_multi.log $out/pytest_shipped.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_2gpu.json 2> $out/bench_2gpu.err
tail -c 900 $out/bench_2gpu.json
bash tools/sweep.sh "FJGPU_FARKEY=0" "FJGPU_FARKEY=1" > $out/sweep_farkey.log 2>&1
cat $out/sweep_farkey.log
FJ_SWEEP_WORKLOAD=config3 bash tools/sweep.sh "FJGPU_FARKEY=0" "FJGPU_FARKEY=1" > $out/sweep_farkey_c3.log 2>&1
cat $out/sweep_farkey_c3.log
ls -la $out

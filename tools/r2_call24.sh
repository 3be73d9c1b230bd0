#!/bin/bash
# Round 2, twenty-fourth GPU call: tile batches per frame (FJGPU_SAMPLE_MB: 16 GiB = 3 batches of the north-star frame, 64 GiB = 1),
# then config 5 with the final kernels.
set -u
out=gpurun_out/r2c24; mkdir -p $out
bash tools/sweep.sh "FJGPU_SAMPLE_MB=16384" "FJGPU_SAMPLE_MB=32768" "FJGPU_SAMPLE_MB=65536" "FJGPU_SAMPLE_MB=8192" "FJGPU_SAMPLE_MB=65536" > $out/sweep.log 2>&1
cat $out/sweep.log
FJGPU_BUILD=device FJ_PARITY_TILES=2 timeout 1500 python bench.py --workload config5 --steps 2 --warmup 3 > $out/bench_config5.json 2> $out/bench_config5.err
tail -c 300 $out/bench_config5.json

#!/bin/bash
# The measurements the round-1 notes leave open, in ONE single-GPU gpurun call (about 4 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/round2_first_call.sh'
# Everything lands in gpurun_out/r2_first/.  Nothing printed under ncu is a bench value.
set -u
out=gpurun_out/r2_first; mkdir -p $out
# 1. does the k_shade slow state (74-78 ms instead of 54-60) survive the chunked queue reservation?  The three settings that
#    used to trigger it on one GPU: a device-built tree, another SAH leaf cost, the counter away from the start of its block.
bash tools/sweep.sh "FJGPU_QUEUE_CHUNK=1" "FJGPU_QUEUE_CHUNK=0" \
  "FJGPU_QUEUE_CHUNK=1 FJGPU_LEAF_COST_X10=10" "FJGPU_QUEUE_CHUNK=0 FJGPU_LEAF_COST_X10=10" \
  "FJGPU_QUEUE_CHUNK=1 FJGPU_BUILD=device" "FJGPU_QUEUE_CHUNK=0 FJGPU_BUILD=device" \
  "FJGPU_QUEUE_CHUNK=1 FJGPU_CTL_OFFSET=8192" "FJGPU_QUEUE_CHUNK=0 FJGPU_CTL_OFFSET=8192" > $out/sweep_k_shade_state.log 2>&1
# 2. ncu --set full of the two kernels that changed at the end of round 1 (north-star scene at 480x270: one launch = 8.8 M rays)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_extend2 -s 2 -c 2 -o $out/k_extend2_coop \
  python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_extend.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 2 -o $out/k_shade_chunked \
  python bench.py --workload profile --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_shade.log 2>&1
python profiles/ncu_summary.py $out/k_extend2_coop.ncu-rep > $out/k_extend2_coop_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_shade_chunked.ncu-rep > $out/k_shade_chunked_ncu_full.txt 2>&1
# 3. the motion-blur workload with every kernel of the end of round 1
timeout 120 python bench.py --workload north_star_motion --steps 5 --warmup 3 > $out/bench_north_star_motion.json 2> $out/bench_motion.err
ls -la $out

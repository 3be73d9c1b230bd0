#!/bin/bash
# Round 2, second GPU call: the whole GPU suite (bridge, shipped scenes, full-size frames, new kernel variants), the A/B of the
# branch-free stack ops / PRMT byte conversion, a default bench line with parity + cpu_baseline, and the DRAM-resident configs.
set -u
out=gpurun_out/r2c2; mkdir -p $out
nproc > $out/nproc.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=20 > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
bash tools/sweep.sh "FJGPU_MAGIC=0" "FJGPU_MAGIC=1" "FJGPU_MAGIC=1 FJGPU_REFILL=8" "FJGPU_MAGIC=0 FJGPU_STACK_SMEM=16" > $out/sweep.log 2>&1
cat $out/sweep.log
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 1500 $out/bench_north_star.json
for cfg in "FJGPU_TOP_NODES=0" "FJGPU_TOP_NODES=85" "FJGPU_MAGIC=1"; do
  echo "config3 $cfg" >> $out/config3.log
  env $cfg timeout 600 python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu-baseline --no-parity 2>>$out/config3.err | tail -1 >> $out/config3.log
done
cat $out/config3.log | cut -c1-600
ls -la $out

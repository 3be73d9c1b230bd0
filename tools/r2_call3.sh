#!/bin/bash
# Round 2, third GPU call: the whole GPU suite, tree-shape and scheduling sweeps of the new k_extend2, the k_shade state sweep
# round 1 left open, ncu --set full of the kernels that ship (a full-size north-star launch), configs 4 and 5.
set -u
out=gpurun_out/r2c3; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
tail -4 $out/pytest.log
bash tools/sweep.sh "FJGPU_LEAF_COST_X10=15" "FJGPU_LEAF_COST_X10=10" "FJGPU_LEAF_COST_X10=20" "FJGPU_LEAF_COST_X10=30" \
  "FJGPU_MAX_LEAF=6 FJGPU_LEAF_COST_X10=10" "FJGPU_MAX_LEAF=8 FJGPU_LEAF_COST_X10=8" "FJGPU_MAX_LEAF=2" \
  "FJGPU_REFILL=8" "FJGPU_REFILL=16" "FJGPU_PHASE_A_MIN=12" "FJGPU_PHASE_A_MIN=20" "FJGPU_PHASE_A_MIN=24 FJGPU_REFILL=8" > $out/sweep_tree.log 2>&1
cat $out/sweep_tree.log
bash tools/sweep.sh "FJGPU_QUEUE_CHUNK=1" "FJGPU_QUEUE_CHUNK=0" "FJGPU_QUEUE_CHUNK=1 FJGPU_CTL_OFFSET=8192" "FJGPU_QUEUE_CHUNK=0 FJGPU_CTL_OFFSET=8192" \
  "FJGPU_QUEUE_CHUNK=1 FJGPU_BUILD=device" "FJGPU_QUEUE_CHUNK=0 FJGPU_BUILD=device" "FJGPU_SORT_BITS=4" > $out/sweep_k_shade_state.log 2>&1
cat $out/sweep_k_shade_state.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_extend2 -s 5 -c 1 -o $out/k_extend2_north_star \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_extend.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 5 -c 1 -o $out/k_shade_north_star \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_shade.log 2>&1
python profiles/ncu_summary.py $out/k_extend2_north_star.ncu-rep > $out/k_extend2_north_star_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_shade_north_star.ncu-rep > $out/k_shade_north_star_ncu_full.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_north_star.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/launches.log 2>&1
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > $out/bench_config4.json 2> $out/bench_config4.err
tail -c 400 $out/bench_config4.json
FJGPU_BUILD=device FJ_PARITY_TILES=2 timeout 1500 python bench.py --workload config5 --steps 2 --warmup 3 > $out/bench_config5.json 2> $out/bench_config5.err
tail -c 400 $out/bench_config5.json
ls -la $out

#!/bin/bash
# Round 2, seventeenth GPU call: evidence for the kernels that ship — ncu --set full of three consecutive k_extend_ring launches
# (camera round + two bounce rounds) and one k_shade launch of the full-size north star, the launch list of the bench command,
# and the bench lines (driver commands first: reference arm, own arm; then the other workloads).
set -u
out=gpurun_out/r2c17; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 4 -c 3 -o $out/k_extend_ring \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_extend.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 5 -c 1 -o $out/k_shade \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_shade.log 2>&1
python profiles/ncu_summary.py $out/k_extend_ring.ncu-rep > $out/k_extend_ring_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_shade.ncu-rep > $out/k_shade_ncu_full.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/launches.log 2>&1
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 600 $out/bench_north_star.json; tail -n 4 $out/bench_north_star.err
for w in north_star_motion config2 config3 config4; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > $out/bench_$w.json 2> $out/bench_$w.err
  tail -c 300 $out/bench_$w.json
done
ls -la $out

#!/bin/bash
# Round 2, eighth GPU call: the evidence of the final kernels — ncu --set full of three k_extend2 launches and one k_shade launch
# of the full-size north star, the launch list of the bench command, the driver's own two commands (reference arm first), config 4.
set -u
out=gpurun_out/r2c8; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend2 -s 4 -c 3 -o $out/k_extend2_final \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_extend.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 5 -c 1 -o $out/k_shade_final \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_shade.log 2>&1
python profiles/ncu_summary.py $out/k_extend2_final.ncu-rep > $out/k_extend2_final_ncu_full.txt 2>&1
python profiles/ncu_summary.py $out/k_shade_final.ncu-rep > $out/k_shade_final_ncu_full.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/launches.log 2>&1
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
tail -c 700 $out/bench_reference_arm.json
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 500 $out/bench_north_star.json; tail -n 4 $out/bench_north_star.err
timeout 900 python bench.py --workload config4 --steps 3 --warmup 3 > $out/bench_config4.json 2> $out/bench_config4.err
tail -c 300 $out/bench_config4.json
ls -la $out

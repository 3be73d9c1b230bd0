#!/bin/bash
# Round 2, twenty-third GPU call: evidence for the FINAL kernels — whole GPU suite, smoke, ncu --set full of three consecutive
# k_extend_ring launches and one k_shade launch, the launch list, the driver's two commands, the other workloads.
set -u
out=gpurun_out/r2c23; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/pytest_all.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 4 -c 3 -o $out/k_extend_ring \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_extend.log 2>&1
python profiles/ncu_summary.py $out/k_extend_ring.ncu-rep > $out/k_extend_ring_ncu_full.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/launches.log 2>&1
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 500 $out/bench_north_star.json
for w in north_star_motion config2 config3 config4; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > $out/bench_$w.json 2> $out/bench_$w.err
  tail -c 200 $out/bench_$w.json
done

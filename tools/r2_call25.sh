#!/bin/bash
# Round 2, twenty-fifth GPU call: default batch budget 64 GiB (the north-star frame is ONE batch: 4 k_extend launches of 141 M
# camera samples' rays instead of 12) — parity of the batching tests, ncu --set full of the three launch kinds, launch list, the
# driver's command, config 5.
set -u
out=gpurun_out/r2c25; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fullsize_gpu.py -m gpu -q -x -k "batch or fullsize or region or config" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log
tail -n 2 $out/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_ring -s 4 -c 3 -o $out/k_extend_ring \
  python bench.py --workload north_star --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_extend.log 2>&1
python profiles/ncu_summary.py $out/k_extend_ring.ncu-rep > $out/k_extend_ring_ncu_full.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $out/launches.log 2>&1
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 400 $out/bench_north_star.json
FJGPU_BUILD=device FJ_PARITY_TILES=2 timeout 1500 python bench.py --workload config5 --steps 2 --warmup 3 > $out/bench_config5.json 2> $out/bench_config5.err
tail -c 300 $out/bench_config5.json
grep -E "== kernel|time_duration|dram__bytes" $out/k_extend_ring_ncu_full.txt

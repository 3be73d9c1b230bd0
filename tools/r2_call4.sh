#!/bin/bash
# Round 2, fourth GPU call: the rest of the GPU suite (parity file onwards, host mirror, shipped scenes), the slot-order
# tie-break A/B, config 3 with parity + cpu_baseline, the driver's own bench command.
set -u
out=gpurun_out/r2c4; mkdir -p $out
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py tests/test_shipped_scenes_gpu.py -m gpu -q --durations=12 > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
tail -6 $out/pytest.log
bash tools/sweep.sh "FJGPU_REFILL=12" "FJGPU_REFILL=8" > $out/sweep.log 2>&1
cat $out/sweep.log
timeout 1200 python bench.py --workload config3 --steps 3 --warmup 3 > $out/bench_config3.json 2> $out/bench_config3.err
tail -c 300 $out/bench_config3.json
timeout 1200 python bench.py --workload config2 --steps 5 --warmup 3 > $out/bench_config2.json 2> $out/bench_config2.err
tail -c 300 $out/bench_config2.json
timeout 600 python bench.py --workload north_star_motion --steps 3 --warmup 3 > $out/bench_north_star_motion.json 2> $out/bench_motion.err
tail -c 300 $out/bench_north_star_motion.json
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 300 $out/bench_north_star.json
ls -la $out

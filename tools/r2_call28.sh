#!/bin/bash
# Round 2, twenty-eighth GPU call: final state of the repository — whole GPU suite, smoke, the driver's two commands.
set -u
out=gpurun_out/r2c28; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/pytest_all.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/summary.log
tail -n 3 $out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 600 $out/bench_north_star.json

#!/bin/bash
# Round 2, thirty-third GPU call: the driver's two commands on the final repository state (batch budget cached).
set -u
out=gpurun_out/r2c33; mkdir -p $out
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 900 $out/bench_north_star.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"

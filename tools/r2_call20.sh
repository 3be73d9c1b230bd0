#!/bin/bash
# Round 2, twentieth GPU call: the final structure (cheap transitions once per outer iteration, compile-time) — whole GPU suite,
# A/B lines, the driver's two commands.
set -u
out=gpurun_out/r2c20; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/pytest_all.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/summary.log
tail -n 4 $out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.log
bash tools/sweep.sh "FJGPU_EXTEND=2" "FJGPU_EXTEND=3" "FJGPU_B1_MIN=20 FJGPU_B2_MIN=6" "FJGPU_REFILL=10" "FJGPU_RING=1" > $out/sweep.log 2>&1
for w in north_star_motion config3 config4; do echo "== $w" >> $out/sweep.log; FJ_SWEEP_WORKLOAD=$w bash tools/sweep.sh "FJGPU_EXTEND=3" >> $out/sweep.log 2>&1; done
cat $out/sweep.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench_north_star.json 2> $out/bench_north_star.err
tail -c 700 $out/bench_north_star.json

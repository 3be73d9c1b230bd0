#!/bin/bash
# Round 2, fourteenth GPU call: k_extend_ring (direct refill, gated heavy phases) as the default — the whole GPU suite, then old
# against new on every bench workload.
set -u
out=gpurun_out/r2c14; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q -x > $out/pytest_all.log 2>&1; echo "pytest -m gpu rc=$?" | tee -a $out/summary.log
tail -n 5 $out/pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/summary.log
for w in north_star north_star_motion config2 config4 config3; do
  echo "== $w" >> $out/sweep.log
  FJ_SWEEP_WORKLOAD=$w bash tools/sweep.sh "FJGPU_EXTEND=2" "FJGPU_EXTEND=3" >> $out/sweep.log 2>&1
done
cat $out/sweep.log

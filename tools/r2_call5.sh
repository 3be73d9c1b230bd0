#!/bin/bash
# Round 2, fifth GPU call: the GPU suite file by file (one process each, so a failure cannot poison the next file's context).
set -u
out=gpurun_out/r2c5; mkdir -p $out
for f in test_gpu_parity test_host_mirror test_shipped_scenes_gpu test_bridge_gpu test_fullsize_gpu; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q --durations=5 > $out/pytest_$f.log 2>&1
  echo "$f rc=$?" | tee -a $out/pytest_summary.log
  tail -3 $out/pytest_$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/pytest_summary.log
FJ_SWEEP_WORKLOAD=config4 bash tools/sweep.sh "FJGPU_REFILL=12" > $out/sweep_config4.log 2>&1
cat $out/sweep_config4.log
ls -la $out

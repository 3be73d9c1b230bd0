#!/bin/bash
# Round 2, twenty-seventh GPU call: ray records and triangle packets loaded without allocating in L1 (LDG.NA) — parity, then the
# standard lines (compare with 276.2-277.1 ms of the build before).
set -u
out=gpurun_out/r2c27; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trace_closest or variants_bit_exact or frame_matches" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/summary.log
tail -n 2 $out/pytest.log
bash tools/sweep.sh "FJGPU_EXTEND=3" "FJGPU_EXTEND=3" "FJGPU_EXTEND=3" > $out/sweep.log 2>&1
for w in config3 config4; do echo "== $w" >> $out/sweep.log; FJ_SWEEP_WORKLOAD=$w bash tools/sweep.sh "FJGPU_EXTEND=3" >> $out/sweep.log 2>&1; done
cat $out/sweep.log

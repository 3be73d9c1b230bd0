#!/bin/bash
# Round 2, thirty-sixth GPU call: the parity file once more on the final library (the batch-budget cache touches every frame).
set -u
out=gpurun_out/r2c36; mkdir -p $out
timeout 140 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/pytest_parity.log 2>&1; echo "parity rc=$?"
tail -n 3 $out/pytest_parity.log

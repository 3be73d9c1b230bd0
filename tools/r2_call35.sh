#!/bin/bash
# Round 2, thirty-fifth GPU call: smoke() and one short bench line on the final library (comment-only rebuild since call 33).
set -u
out=gpurun_out/r2c35; mkdir -p $out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 $out/smoke.log
bash tools/sweep.sh "FJGPU_EXTEND=3" 2>&1 | tee $out/sweep.log

#!/bin/bash
# round 2, GPU call 14 (1 GPU): what the driver runs at round end - GPU suite (incl. the slow N=16384 elementwise test),
# smoke(), the default bench and the reference arm
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call14_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== GPU suite =="
timeout 1800 python -m pytest tests -m gpu -q -x --durations=5 > $OUT/r02c14_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -14 $OUT/r02c14_gpu_tests.log | tee -a $S
say "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee -a $S
say "== bench (defaults) =="
timeout 900 python bench.py > $OUT/r02c14_bench.json 2> $OUT/r02c14_bench.err; say "bench rc=$?"
grep "^{" $OUT/r02c14_bench.json | tail -c 3800 | tee -a $S; tail -3 $OUT/r02c14_bench.err | tee -a $S
say "== bench --impl reference (steps 3, warmup 1) =="
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -c 1500 | tee -a $S
say "== done =="

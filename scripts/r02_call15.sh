#!/bin/bash
# round 2, GPU call 15 (1 GPU): GPU suite after the test fix (whole suite, no -x), e2e with 16 upload groups
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call15_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== GPU suite (without the slow N=16384 test, which passed in call 14) =="
timeout 1800 python -m pytest tests -m "gpu and not slow" -q > $OUT/r02c15_gpu_tests.log 2>&1; say "gpu tests rc=$?"; tail -6 $OUT/r02c15_gpu_tests.log | tee -a $S
say "== bench (defaults) =="
timeout 900 python bench.py --skip-parity > $OUT/r02c15_bench.json 2> $OUT/r02c15_bench.err; say "bench rc=$?"
grep "^{" $OUT/r02c15_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f  e2e %.0f  other %s  frac %.3f' % (d['value'], d['e2e']['value'], d['config']['other_api'], d['roofline']['frac']))" | tee -a $S
say "== done =="

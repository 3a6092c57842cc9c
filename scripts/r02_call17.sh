#!/bin/bash
# round 2, GPU call 17 (1 GPU): ncu launch list of bench.py itself (one timed step of the headline workload)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call17_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== bench.py N=65536, 3 warm-up + 1 timed step, without ncu =="
timeout 600 python bench.py --steps 1 --warmup 3 --skip-parity --skip-cpu-baseline --e2e-steps 0 > $OUT/r02c17_bench.json 2> $OUT/r02c17_bench.err; say "rc=$?"
grep "^{" $OUT/r02c17_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f  ms/step %.1f  update share %.4f  frac %.3f  launches %d' % (d['value'], d['ms_per_step'], d['roofline']['share_of_step'], d['roofline']['frac'], d['gpu_launches']))" | tee -a $S
say "== the same command under ncu (launch list; the first 7700 launches = set-up and warm-up are skipped) =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 7700 -c 3200 --csv --log-file $OUT/r02c17_launches_bench_n65536.csv \
  python bench.py --steps 1 --warmup 3 --skip-parity --skip-cpu-baseline --e2e-steps 0 > $OUT/r02c17_ncu.log 2>&1; say "ncu rc=$?"
wc -l $OUT/r02c17_launches_bench_n65536.csv | tee -a $S
say "== done =="

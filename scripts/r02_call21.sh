#!/bin/bash
# round 2, GPU call 21 (1 GPU, what is left of the round's GPU time): generation kernel with read-back stores (lung in
# registers) and the jump kernel with 32 / 48 / 64 segments per CTA on hardware - parity tests for the two candidate
# defaults, then timing of every variant (one python process).
# (ran scripts/rng_final.py as of commit 6f93a39, which still drove the kernel-variant test hook; see the note in that script)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
RNG_FINAL_BUDGET_S=36 timeout 46 python scripts/rng_final.py > $OUT/r02c21_rng_final.jsonl 2> $OUT/r02c21_rng_final.err
echo "rc=$?" | tee $OUT/r02_call21_summary.txt
grep "^{" $OUT/r02c21_rng_final.jsonl | tee -a $OUT/r02_call21_summary.txt
tail -5 $OUT/r02c21_rng_final.err | tee -a $OUT/r02_call21_summary.txt

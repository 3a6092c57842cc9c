#!/bin/bash
# round 2, GPU call 22 (1 GPU, the last ~35 GPU-seconds of the round): the generation kernel with four segments per warp
# (groups of 8 lanes) on hardware - its parity tests, then timing against the one-segment-per-warp kernel.
# (ran scripts/rng_final.py as of commit the commit that added this file, which still drove the kernel-variant test hook; see the note in that script)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
RNG_FINAL_BUDGET_S=16 timeout 20 python scripts/rng_final.py > $OUT/r02c22_rng_final.jsonl 2> $OUT/r02c22_rng_final.err
echo "rc=$?" | tee $OUT/r02_call22_summary.txt
grep "^{" $OUT/r02c22_rng_final.jsonl | tee -a $OUT/r02_call22_summary.txt
tail -5 $OUT/r02c22_rng_final.err | tee -a $OUT/r02_call22_summary.txt

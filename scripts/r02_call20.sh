#!/bin/bash
# round 2, GPU call 20 (1 GPU, the round's last GPU-minute): the rewritten jump kernel (16 segments per CTA, K rows
# streamed through registers), the generation kernel storing from registers, and the pipelined-pieces option on
# hardware: parity tests, then timing over segment bias x pieces x size (one python process).
# (ran scripts/rng_final.py as of commit 612ab67, which still drove the kernel-variant test hook; see the note in that script)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
RNG_FINAL_BUDGET_S=55 timeout 75 python scripts/rng_final.py > $OUT/r02c20_rng_final.jsonl 2> $OUT/r02c20_rng_final.err
echo "rc=$?" | tee $OUT/r02_call20_summary.txt
grep "^{" $OUT/r02c20_rng_final.jsonl | tee -a $OUT/r02_call20_summary.txt
tail -5 $OUT/r02c20_rng_final.err | tee -a $OUT/r02_call20_summary.txt

#!/bin/bash
# round 2, GPU call 19 (1 GPU, the round's last ~2 GPU-minutes): the device dSFMT-19937 generator on hardware -
# its GPU parity tests (bit-exact against the oracle's sequential generator), then the achieved write bandwidth.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call19_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
say "== device dSFMT: GPU parity tests =="
timeout 60 python -m pytest tests/test_zz_gpu_rng.py -q -s > $OUT/r02c19_rng_tests.log 2>&1; say "rc=$?"
tail -4 $OUT/r02c19_rng_tests.log | tee -a $S
grep "device dSFMT" $OUT/r02c19_rng_tests.log | tee -a $S
say "== device dSFMT: achieved bandwidth =="
RNG_SIZES=${RNG_SIZES:-67108864,268435456,1073741824} RNG_REPS=3 timeout 35 python scripts/rng_bench.py > $OUT/r02c19_rng_bench.jsonl 2> $OUT/r02c19_rng_bench.err; say "rc=$?"
cat $OUT/r02c19_rng_bench.jsonl | tee -a $S
tail -3 $OUT/r02c19_rng_bench.err | tee -a $S
say "== smoke() with the rebuilt library =="
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02c19_smoke.log 2>&1; say "rc=$?"
tail -2 $OUT/r02c19_smoke.log | tee -a $S
say "== done =="

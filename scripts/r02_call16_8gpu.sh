#!/bin/bash
# round 2, GPU call 16 (8 GPUs, short): block size of the distributed layout at P = 8 with the faster panel chain
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call16_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
for NB in 512 256; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$((NB/256)) bench.py --gpus 8 --steps 3 --warmup 2 \
     --skip-parity --e2e-steps 0 --block $NB > $OUT/r02c16_bench_nb$NB.json 2> $OUT/r02c16_bench_nb$NB.err
  grep "^{" $OUT/r02c16_bench_nb$NB.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('P=8 N=65536 nb=%d: value %.0f GFLOP/s  ms/step %.1f  update-kernel frac %.3f share %.3f  resid %.2e info %d' % (d['config']['nb'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d['config']['probe_resid'], d['config']['info']))" | tee -a $S
done
say "== done =="

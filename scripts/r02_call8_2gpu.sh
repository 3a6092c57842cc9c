#!/bin/bash
# round 2, GPU call 8 (2 GPUs): in-process multi-GPU (worker threads + ncclCommInitAll), slab-distributed inverse, C unit
# program, torchrun cholDist parity, bench at P = 2 with the leading-block parity and the in-process e2e
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call8_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
nvidia-smi -L | tee -a $S
say "== multi-GPU tests (all visible devices in ONE process) =="
timeout 1200 python -m pytest tests/test_gpu_multigpu.py tests/test_dropin_c.py -m gpu -q -rs > $OUT/r02c8_mgpu_tests.log 2>&1; say "rc=$?"; tail -8 $OUT/r02c8_mgpu_tests.log | tee -a $S
say "== host-pointer POTRI: staged on one GPU vs slab-distributed over all =="
for N in 16384 32768; do AB_N=$N timeout 600 python scripts/mgpu_inverse_ab.py 2>&1 | grep potri | tee -a $S; done
say "== bench --gpus 2, N=65536 =="
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 \
   > $OUT/r02c8_bench_n65536_2gpu.json 2> $OUT/r02c8_bench_2gpu.err; say "bench rc=$?"
tail -c 3500 $OUT/r02c8_bench_n65536_2gpu.json | tee -a $S; tail -5 $OUT/r02c8_bench_2gpu.err | tee -a $S
say "== done =="

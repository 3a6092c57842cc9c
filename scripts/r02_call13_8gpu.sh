#!/bin/bash
# round 2, GPU call 13 (8 GPUs): headline bench at P = 8, in-process multi-GPU tests on 8 devices, distributed POTRI, N = 131072
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
S=$OUT/r02_call13_summary.txt; : > $S
say() { echo "$@" | tee -a $S; }
nvidia-smi -L | wc -l | tee -a $S
say "== bench --gpus 8, N=65536 =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 2 --warmup 3 \
   > $OUT/r02c13_bench_n65536_8gpu.json 2> $OUT/r02c13_bench_8gpu.err; say "bench rc=$?"
grep "^{" $OUT/r02c13_bench_n65536_8gpu.json | tail -c 3500 | tee -a $S; tail -3 $OUT/r02c13_bench_8gpu.err | tee -a $S
say "== host-pointer POTRI, distributed over 8 GPUs =="
for N in 32768 65536; do AB_MODES=distributed AB_N=$N timeout 600 python scripts/mgpu_inverse_ab.py 2>&1 | grep potri | tee -a $S; done
say "== multi-GPU tests (8 devices in ONE process) =="
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -rs > $OUT/r02c13_mgpu_tests.log 2>&1; say "rc=$?"; tail -6 $OUT/r02c13_mgpu_tests.log | tee -a $S
say "== bench --gpus 8, N=131072 (no e2e: the host matrix would be 137 GB) =="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 2 --warmup 3 --order 131072 --e2e-steps 0 \
   > $OUT/r02c13_bench_n131072_8gpu.json 2> $OUT/r02c13_bench_131072.err; say "bench rc=$?"
grep "^{" $OUT/r02c13_bench_n131072_8gpu.json | tail -c 2500 | tee -a $S; tail -3 $OUT/r02c13_bench_131072.err | tee -a $S
say "== done =="

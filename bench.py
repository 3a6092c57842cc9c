#!/usr/bin/env python
"""bench.py — DPOTRF GFLOP/s (N^3/3) at N = 65536 on 1/2/4/8 B200 (BASELINE.json `metric`).

  python bench.py --gpus 1 --steps K --warmup W            # this library, one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W      # one rank per GPU, NCCL
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU dpotrf on host cores

One "step" = one full Cholesky factorisation (lower, FP64) of the synthetic SPD matrix
A = B*B^T + N*I (B: N x 1024, uniform(-0.5, 0.5), fixed seed), restored before each step inside the
timed region.  `value` is whole-job GFLOP/s with the matrix already resident in HBM (for N > 1: in
the 1-D block-cyclic layout across the ranks); `e2e` is the same metric through the host-pointer
C-ABI call (cuMultiGPUDpotrf / host shard -> cholDistDpotrf -> host shard) with the H2D and D2H
copies inside the timed region.  The JSON line also carries the roofline of the dominant kernel
(the DMMA trailing update, timed per launch with CUDA events on its own stream, against the FP64
tensor-core peak measured in the same process) and the reference's CPU dpotrf timed on this box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DPOTRF GFLOP/s (N^3/3) at N=65536, 1/2/4/8 B200, % of FP64 TC peak"
SEED = 20121018
KGEN = 1024  # inner dimension of the generator B (N x 1024 for large N)


# ---------------------------------------------------------------------------------------------
# host-side helpers of the multi-process path (covered by tests/test_host_logic.py with gloo)
# ---------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def exchange_unique_id(make_id, rank):
    """rank 0 creates the 128-byte NCCL id, everyone gets it."""
    dist = _dist()
    box = [make_id() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def owned_columns(n, nb, rank, nranks):
    """global column indices of the rank's packed local columns (block c = lc*nranks + rank)."""
    import numpy as np
    nblk = (n + nb - 1) // nb
    parts = [np.arange(c * nb, min(n, c * nb + nb)) for c in range(rank, nblk, nranks)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)


def combine_info(local_info):
    """first failing column over all ranks (0 = success)."""
    dist = _dist()
    if dist is None:
        return int(local_info)
    allv = [None] * dist.get_world_size()
    dist.all_gather_object(allv, int(local_info))
    pos = [v for v in allv if v > 0]
    return min(pos) if pos else 0


def max_over_ranks(x):
    dist = _dist()
    if dist is None:
        return float(x)
    allv = [None] * dist.get_world_size()
    dist.all_gather_object(allv, float(x))
    return max(allv)


def barrier():
    dist = _dist()
    if dist is not None:
        dist.barrier()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


# ---------------------------------------------------------------------------------------------
def cpu_reference_lib():
    """(ctypes lib, kind): the reference's own CPU code if it was compiled (oracle/_ref), else the port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    if H.ref() is not None:
        return H.CpuLib("ref"), "reference", H
    return H.CpuLib("oracle"), "port", H


def use_all_host_cores():
    """The CPU reference is OpenMP code; torchrun exports OMP_NUM_THREADS=1, so set the thread count
    explicitly to every core this process may run on.  Returns the count used."""
    import ctypes
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(ctypes.c_int(cores))
    except OSError:
        cores = int(os.environ.get("OMP_NUM_THREADS", "1"))
    return cores


def cpu_spd(n, k=KGEN):
    """A = B*B^T + n*I with B (n x min(k, n), column-major) = u - 0.5 from the configs' generator, dSFMT-19937 seed SEED
    (the oracle's restatement of the reference's rng/dsfmt_19937.c; this CPU leg is where bench.py may run oracle/ code)"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    k = min(k, n)
    B = (H.dsfmt_uniform(SEED, n * k) - 0.5).reshape(k, n).T
    A = B @ B.T + n * np.eye(n)
    return np.asfortranarray(A)


def time_cpu_dpotrf(n, steps, warmup):
    import numpy as np
    lib, kind, _ = cpu_reference_lib()
    A0 = cpu_spd(n)
    times = []
    for it in range(warmup + steps):
        A = A0.copy(order="F")
        t0 = time.perf_counter()
        info = lib.potrf("d", "L", A)
        dt = time.perf_counter() - t0
        assert info == 0
        if it >= warmup:
            times.append(dt)
    L = np.tril(A)
    resid = float(np.linalg.norm(L @ L.T - A0) / np.linalg.norm(A0))
    return sum(times) / len(times), kind, resid


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = use_all_host_cores()
    # sample size: the CPU rate FALLS with N (left-looking nb=32 loop), so the larger sample is
    # the more faithful one; take N=8192 when the whole --steps/--warmup run still ends within a few minutes
    # (estimated from one N=4096 factorisation), else N=4096.  Both rates go on record.
    n_small = min(4096, args.n)
    sec_small, kind, _ = time_cpu_dpotrf(n_small, 1, 0)
    rates = {str(n_small): n_small ** 3 / 3.0 / sec_small / 1e9}
    n_s = args.ref_n
    if n_s == 0:
        est = sec_small * 8.0 * 1.6 * (args.steps + args.warmup)  # (8192/4096)^3 and a falling rate
        n_s = 8192 if (est <= 240.0 and args.n >= 8192) else n_small
    sec, kind, resid = time_cpu_dpotrf(n_s, args.steps, args.warmup)
    gf = n_s ** 3 / 3.0 / sec / 1e9
    rates[str(n_s)] = gf
    sample = "reference CPU dpotrf (lower) of a %dx%d SPD matrix of the N=%d workload's family (A = B*B^T + n*I), per step" % (n_s, n_s, args.n)
    line = {"impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "dpotrf lower N=%d, FP64 DMMA trailing update (BASELINE configs[1]/[4] family at the metric's N); "
                                   "this arm: CPU sample of order %d per step (N=%d would take hours)" % (args.n, n_s, args.n),
                       "N": args.n, "uplo": "L", "sample_N": n_s, "sample_resid": resid, "cpu_gflops_by_sample_N": rates,
                       "generator": "dSFMT-19937 seed %d (CPU, oracle/dsfmt_oracle.c = the reference's rng/dsfmt_19937.c) - 0.5, "
                                    "A = B*B^T + n*I with B n x %d column-major" % (SEED, KGEN),
                       "note": "CPU reference (lapack/dpotrf.c + blas/*.c, compiled unmodified) timed on a bounded sample per step; "
                               "its rate falls with N, so the GPU/CPU ratio is understated"},
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
LEAD = 8192  # order of the leading principal block compared elementwise with the reference's CPU dpotrf (SURVEY.md section 8d cfg 5)


def lead_block_parity(A, A0, cols, N, P, rank):
    """Gather the leading LEAD x LEAD block of the factor (and of the input) from the block-cyclic layout to rank 0
    and compare it elementwise with the reference's own CPU dpotrf of that block (the Cholesky factor of a leading
    principal block is the leading block of the factor).  Returns (maxrel, bound, cpu_seconds, kind, cores) on
    rank 0, None elsewhere.  Not timed; the only place besides cpu_baseline where bench.py runs oracle/ code."""
    import numpy as np
    m = min(LEAD, N)
    sel = np.nonzero(cols < m)[0]
    mine = (cols[sel], A[sel][:, :m].cpu().numpy(), A0[sel][:, :m].cpu().numpy())  # rows of the tensors = matrix columns
    dist = _dist()
    parts = [mine]
    if dist is not None:
        parts = [None] * P
        dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    L = np.zeros((m, m), order="F")
    Ain = np.zeros((m, m), order="F")
    for gc, lpart, apart in parts:
        L[:, gc] = lpart.T
        Ain[:, gc] = apart.T
    cores = use_all_host_cores()
    lib, kind, _ = cpu_reference_lib()
    want = Ain.copy(order="F")
    t0 = time.perf_counter()
    info = lib.potrf("d", "L", want)
    sec = time.perf_counter() - t0
    assert info == 0, info
    ref = np.tril(want)
    maxrel = float(np.abs(np.tril(L) - ref).max() / np.abs(ref).max())
    bound = 64.0 * m * 2.220446049250313e-16
    return maxrel, bound, sec, kind, cores


def dram_profile():
    """DRAM bytes of ONE launch of the dominant kernel from the committed `ncu --set full` capture of this round
    (profiles/r02_ncu_dmma_update.json, written by scripts/ncu_summary.py from the .ncu-rep); None if absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_dmma_update.json")))
    except (OSError, ValueError):
        return None


def run_ours(args):
    import numpy as np
    import torch
    from cholesky_b200_loader import load
    chol = load()
    chol.lib()  # fails loudly if the CUDA library is missing: there is no fallback

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    P = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != P:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, P))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if P > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    N, K, Wm = args.n, args.steps, args.warmup

    # FP64 tensor-core peak of THIS GPU, measured now (MEASURED_PEAKS.json has no FP64 entry)
    peak_tf = max(chol.probe_fp64(0, 20000)[0] for _ in range(3)) if rank == 0 else 0.0

    uid = exchange_unique_id(chol.dist_unique_id, rank) if P > 1 else None
    d = chol.Dist(rank, P, uid, nb=args.nb)
    nb = d.block_size("d", N)
    cols = owned_columns(N, nb, rank, P)
    nloc = len(cols)
    cols_t = torch.from_numpy(cols).to(dev)

    # synthetic SPD input, generated in place in the distributed layout (row-major (nloc, N) tensor ==
    # column-major N x nloc with ld = N): A[:, cols] = B * B[cols, :]^T + N * I
    # B (column-major N x KGEN) = u - 0.5 with u from the generator the configs name, dSFMT-19937, produced on the device by the
    # library's cuRng64GetCloseOpen (bit-identical to the reference's CPU stream; every rank generates the whole of B)
    Bt = torch.empty(KGEN, N, dtype=torch.float64, device=dev)
    try:
        rng = chol.Rng64(SEED)
        rng.get(Bt.data_ptr(), KGEN * N)
        torch.cuda.synchronize()
        rng.close()
        generator = ("dSFMT-19937 seed %d on the device (cuRng64GetCloseOpen: the reference's rng/dsfmt_19937.c stream) - 0.5, "
                     "A = B*B^T + N*I with B N x %d column-major" % (SEED, KGEN))
    except Exception as e:  # noqa: BLE001 - input generation only; the measured path has no fallback
        g = torch.Generator(device=dev).manual_seed(SEED)
        Bt = torch.rand(KGEN, N, dtype=torch.float64, device=dev, generator=g)
        generator = "torch.rand (Philox, seed %d) - 0.5 (device dSFMT failed: %r), A = B*B^T + N*I with B N x %d" % (SEED, e, KGEN)
    Bt -= 0.5
    A0 = torch.empty(nloc, N, dtype=torch.float64, device=dev)
    for s in range(0, nloc, 4096):
        e = min(nloc, s + 4096)
        A0[s:e] = Bt[:, cols_t[s:e]].T @ Bt
    A0[torch.arange(nloc, device=dev), cols_t] += float(N)
    A = torch.empty_like(A0)
    torch.cuda.synchronize()

    # P == 1: the timed call is the drop-in entry point cuDpotrf itself (the packed layout of one rank IS the plain
    # column-major matrix); P > 1: cholDistDpotrf, one rank per GPU
    lap = d.lapack

    def step_dist():
        A.copy_(A0)
        return d.potrf("d", N, A.data_ptr(), N)

    def step_cu():
        A.copy_(A0)
        return lap.potrf("d", "L", N, A.data_ptr(), N)

    step = step_cu if P == 1 else step_dist
    timed_api = "cuDpotrf" if P == 1 else "cholDistDpotrf"

    def timed(fn, k):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            inf = fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), inf

    for _ in range(Wm):
        info = step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    chol.launch_count(reset=True)
    lap.profile(True)
    sampler.start()
    ms_total, info = timed(step, K)
    clocks = sampler.stop()
    launches = chol.launch_count()
    upd_ms, upd_flops, upd_launches = lap.profile_read()
    hist = lap.profile_history() if os.environ.get("CHOL_BENCH_HIST") else None
    lap.profile(False)
    info = combine_info(info)
    ms_per_step = ms_total / K
    gflops = N ** 3 / 3.0 / (ms_per_step * 1e-3) / 1e9

    # the same factorisation through the other device-pointer entry point (same kernels, separate driver loop)
    other = None
    if P == 1:
        step_dist()
        ms_o, _ = timed(step_dist, min(K, 2))
        other = {"api": "cholDistDpotrf (1 rank)", "gflops": N ** 3 / 3.0 / (ms_o / min(K, 2) * 1e-3) / 1e9}
        other["rel_diff"] = abs(other["gflops"] - gflops) / gflops

    # ---- untimed correctness at full size: || A x - L (L^T x) || / || A x || ----
    x = torch.rand(N, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    rows = torch.arange(N, device=dev)
    y_loc = torch.empty(nloc, dtype=torch.float64, device=dev)
    for s in range(0, nloc, 2048):
        e = min(nloc, s + 2048)
        Lc = torch.where(rows[None, :] >= cols_t[s:e, None], A[s:e], torch.zeros((), dtype=torch.float64, device=dev))
        y_loc[s:e] = Lc @ x  # (L^T x)[cols]
    z = torch.zeros(N, dtype=torch.float64, device=dev)
    ax = torch.zeros(N, dtype=torch.float64, device=dev)
    for s in range(0, nloc, 2048):
        e = min(nloc, s + 2048)
        Lc = torch.where(rows[None, :] >= cols_t[s:e, None], A[s:e], torch.zeros((), dtype=torch.float64, device=dev))
        z += y_loc[s:e] @ Lc
        ax += x[cols_t[s:e]] @ A0[s:e]
    del Lc
    if P > 1:
        dist.all_reduce(z)
        dist.all_reduce(ax)
    probe_resid = (torch.linalg.norm(z - ax) / torch.linalg.norm(ax)).item()

    # ---- untimed parity with the reference: leading 8192 x 8192 block, elementwise ----
    parity = None if args.no_parity else lead_block_parity(A, A0, cols, N, P, rank)

    # ---- e2e: host buffers, copies inside the timed region ----
    e2e_steps = max(1, min(K, args.e2e_steps))
    e2e_times = []
    e2e_note = None
    if args.e2e_steps == 0:  # records at sizes whose host matrix does not fit a sensible run (N = 131072: 137 GB)
        h2d_bytes = d2h_bytes = 0
        e2e_times, info2, e2e_api = [float("nan")], 0, "skipped (--e2e-steps 0)"
    elif P == 1:
        h2d_bytes = d2h_bytes = nloc * N * 8
        hostA = torch.empty(nloc, N, dtype=torch.float64, pin_memory=True)
        del A  # the host-pointer API owns its own device copy
        torch.cuda.empty_cache()
        mg = chol.MultiGPU([local_rank])
        if args.nb:
            mg.set_block_size(args.nb)
        for it in range(1 + e2e_steps):
            hostA.copy_(A0)  # restore the input on the host (untimed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info2 = mg.potrf("d", "L", N, hostA.data_ptr(), N)  # scatter (H2D) + factor + gather (D2H)
            dt = time.perf_counter() - t0
            if it >= 1:
                e2e_times.append(dt)
        mg.close()
        e2e_api = "cuMultiGPUDpotrf(host pointer), 1 GPU"
    elif args.e2e_api == "shards":
        h2d_bytes = d2h_bytes = nloc * N * 8
        hostA = torch.empty(nloc, N, dtype=torch.float64, pin_memory=True)
        for it in range(1 + e2e_steps):
            hostA.copy_(A0)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            A.copy_(hostA, non_blocking=True)
            info2 = d.potrf("d", N, A.data_ptr(), N)
            hostA.copy_(A, non_blocking=True)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            if it >= 1:
                e2e_times.append(dt)
        e2e_api = "pinned host shard -> cholDistDpotrf -> pinned host shard, %d ranks" % P
    else:
        # the drop-in API itself: ONE process (rank 0) drives all P GPUs through cuMultiGPUDpotrf on a page-locked
        # N x N host matrix (one worker thread per GPU, NCCL inside the library); the other ranks have released
        # their device memory and wait.  34 GB up + 34 GB down over PCIe inside the timed region.
        h2d_bytes = d2h_bytes = N * N * 8
        del A, A0
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        # the waiting ranks must not spin in an NCCL kernel on the GPUs rank 0 is about to use: wait on the CPU (gloo)
        gl = dist.new_group(backend="gloo")
        dist.barrier(group=gl)
        info2 = 0
        if rank == 0:
            hostA = torch.empty(N, N, dtype=torch.float64, pin_memory=True)  # row j = matrix column j (symmetric input)

            def fill_host():  # the same matrix family, regenerated on rank 0's GPU chunk by chunk (untimed)
                for s0 in range(0, N, 4096):
                    e0_ = min(N, s0 + 4096)
                    blk = Bt[:, s0:e0_].T @ Bt
                    blk[torch.arange(e0_ - s0, device=dev), torch.arange(s0, e0_, device=dev)] += float(N)
                    hostA[s0:e0_].copy_(blk)
                    del blk
                torch.cuda.synchronize()
                torch.cuda.empty_cache()

            mg = chol.MultiGPU(list(range(P)))
            if args.nb:
                mg.set_block_size(args.nb)
            for it in range(1 + e2e_steps):
                fill_host()
                if it == 0:
                    A0s = hostA[:2048, :2048].clone()
                t0 = time.perf_counter()
                info2 = mg.potrf("d", "L", N, hostA.data_ptr(), N)
                dt = time.perf_counter() - t0
                if it >= 1:
                    e2e_times.append(dt)
            mg.close()
            Lh = torch.tril(hostA[:2048, :2048].T)  # spot check of what came back: leading block residual
            e2e_note = {"lead2048_resid": (torch.linalg.norm(Lh @ Lh.T - A0s) / torch.linalg.norm(A0s)).item()}
        dist.barrier(group=gl)
        if rank != 0:
            e2e_times = [0.0]
        e2e_api = "cuMultiGPUDpotrf(host pointer), ONE process driving %d GPUs (rank 0; the other ranks idle)" % P
    info2 = combine_info(info2)
    e2e_s = max_over_ranks(sum(e2e_times) / len(e2e_times))
    e2e_gflops = N ** 3 / 3.0 / e2e_s / 1e9
    d.close()

    if rank != 0:
        if P > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (trailing-update DMMA tiles) ----
    achieved_tf = (upd_flops / (upd_ms * 1e-3) / 1e12) if upd_ms > 0 else None
    prof = dram_profile()
    roofline = {"bound": "tensor", "kernel": "chol::gemm_dmma_kernel<false,0> (trailing update A22 -= W*W^T)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if achieved_tf else None,
                # DRAM bytes (read + written) of ONE launch of this kernel from this round's committed ncu --set full capture
                "traffic": prof.get("dram_bytes") if prof else None,
                "traffic_launch": prof.get("launch") if prof else None,
                "algorithmic_bytes_of_that_launch": prof.get("algorithmic_bytes") if prof else None,
                "ncu_tensor_pipe_active_pct": prof.get("tensor_pipe_active_pct") if prof else None,
                "traffic_source": "profiles/r02_ncu_dmma_update.json (ncu capture of the shipped kernel, not this run)" if prof else None,
                "launches": upd_launches, "ms_total": upd_ms,
                "share_of_step": upd_ms / ms_total if ms_total else None,
                "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) peak measured in this run by chol_probe_fp64; "
                               "MEASURED_PEAKS.json has no FP64 entry (bf16 %.0f TF/s, HBM %.0f GB/s)"
                               % (measured_peaks().get("bf16_tflops", 0), measured_peaks().get("hbm_gbs", 0)),
                "whole_factorisation_frac_of_peak": gflops / 1e3 / (peak_tf * P) if peak_tf else None}

    # ---- the reference's CPU path on this box's host cores (bounded sample) ----
    cpu = None
    if P == 1 and not args.no_cpu_baseline:
        if parity is not None:  # the parity run above IS a timed reference factorisation of the workload's leading block
            m = min(LEAD, N)
            cpu = {"value": m ** 3 / 3.0 / parity[2] / 1e9, "unit": "GFLOP/s", "cores": parity[4], "kind": parity[3],
                   "sample": "reference CPU dpotrf (lower) of the leading %dx%d principal block of this run's input, one run, %.1f s"
                             % (m, m, parity[2])}
        else:
            cores = use_all_host_cores()
            sec, kind, _ = time_cpu_dpotrf(args.cpu_n, 1, 0)
            cpu = {"value": args.cpu_n ** 3 / 3.0 / sec / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                   "sample": "reference CPU dpotrf (lower) of a %dx%d SPD matrix of the same family, one run, %.1f s"
                             % (args.cpu_n, args.cpu_n, sec)}

    line = {"metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": P, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "dpotrf lower N=%d, FP64 DMMA trailing update (BASELINE configs[1]/[4] family at the metric's N)" % N,
                       "N": N, "uplo": "L", "nb": nb, "layout": "1-D block-cyclic block columns over %d rank(s)" % P,
                       "timed_api": timed_api, "other_api": other,
                       "generator": generator,
                       "parallelism": "bc%d" % P, "l2": "inputs (%.1f GB per rank) are larger than L2" % (nloc * N * 8 / 1e9),
                       "restore_in_timed_region": True, "info": info, "probe_resid": probe_resid,
                       "lead_parity_maxrel": parity[0] if parity else None, "lead_parity_bound": parity[1] if parity else None,
                       "lead_parity": ("leading %d x %d block of the factor vs the reference's CPU dpotrf (%s), elementwise, "
                                       "max|L - L_ref| / max|L_ref|" % (min(LEAD, N), min(LEAD, N), parity[3])) if parity else None,
                       # library switches in effect (empty = the defaults the committed numbers were measured with)
                       "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("CHOL_") and k != "CHOL_BENCH_HIST"}},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "s_per_step": e2e_s, "steps": e2e_steps, "api": e2e_api, "info": info2, "check": e2e_note},
            "roofline": roofline}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if hist:
        sys.stderr.write("update-kernel launches (ms, TFLOP/s): " + " ".join("%.3f/%.1f" % (m, f / m / 1e9) for m, f in hist[:len(hist) // K]) + "\n")
    print(json.dumps(line), flush=True)
    if P > 1:
        dist.destroy_process_group()
    bad = []
    if info != 0 or info2 != 0:
        bad.append("info %d / %d" % (info, info2))
    if parity is not None and not parity[0] <= parity[1]:
        bad.append("leading-block parity %.3e > %.3e" % (parity[0], parity[1]))
    if not probe_resid <= 1e-12:
        bad.append("residual probe %.3e" % probe_resid)
    if bad:
        sys.stderr.write("bench.py: FAILED correctness: " + "; ".join(bad) + "\n")
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", dest="n", type=int, default=65536, help="matrix order N (the metric's N = 65536)")
    ap.add_argument("--block", dest="nb", type=int, default=0, help="outer block size (0 = library default)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-order", dest="cpu_n", type=int, default=6144, help="order of the CPU-baseline sample")
    ap.add_argument("--ref-order", dest="ref_n", type=int, default=0, help="order of the per-step sample of --impl reference (0 = 8192 if it fits a few minutes, else 4096)")
    ap.add_argument("--skip-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--skip-parity", dest="no_parity", action="store_true", help="skip the leading-block comparison with the CPU reference")
    ap.add_argument("--e2e-api", default="inprocess", choices=["inprocess", "shards"],
                    help="N > 1: e2e through cuMultiGPUDpotrf in ONE process (default) or per-rank shards around cholDistDpotrf")
    args = ap.parse_args()  # option names avoid every prefix of torch.distributed.run's own options
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

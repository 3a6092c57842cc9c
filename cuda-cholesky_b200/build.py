"""Build libcholesky_b200.so in-tree: nvcc (sm_100a) for csrc/*.cu, gcc (C99) for host/*.c.

  python cuda-cholesky_b200/build.py          # incremental
The shared library is the drop-in boundary (include/*.h); it travels to the GPU box with the
gpurun snapshot (git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcholesky_b200.so")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA, "bin", "nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CC_FLAGS = ["-std=c99", "-O2", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter", "-D_POSIX_C_SOURCE=200809L",
            "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include")]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src, *extra))


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return p.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    csrc, host = os.path.join(HERE, "csrc"), os.path.join(HERE, "host")
    cu = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    cc = sorted(f for f in os.listdir(host) if f.endswith(".c"))
    hdrs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".h")]
    hdrs += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    jobs, objs = [], []
    for f in cu:
        src, o = os.path.join(csrc, f), os.path.join(OBJ, f[:-3] + ".o")
        objs.append(o)
        if force or _newer(src, o, hdrs):
            jobs.append(([NVCC, *NVCC_FLAGS, "-c", src, "-o", o], os.path.join(OBJ, f[:-3] + ".ptxas.log")))
    for f in cc:
        src, o = os.path.join(host, f), os.path.join(OBJ, "host_" + f[:-2] + ".o")
        objs.append(o)
        if force or _newer(src, o, hdrs):
            jobs.append((["gcc", *CC_FLAGS, "-c", src, "-o", o], None))
    with ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(lambda j: _run(*j), jobs):
            if verbose and out.strip():
                print(out)
    if jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB, *objs, "-L" + os.path.join(CUDA, "lib64"), "-L" + os.path.join(CUDA, "lib64", "stubs"),
              "-lcuda", "-lpthread", "-lm"] + _nccl_link())
    return LIB


def _nccl_link():
    """NCCL for the multi-GPU path: link by soname so the copy torch already loaded is reused."""
    for d in ("/usr/lib/x86_64-linux-gnu", os.path.join(CUDA, "lib64")):
        if os.path.exists(os.path.join(d, "libnccl.so.2")) or os.path.exists(os.path.join(d, "libnccl.so")):
            return []  # resolved lazily with dlopen in host/multigpu.c - no link-time dependency
    return []


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

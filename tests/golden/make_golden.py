"""Generate tests/golden/* from the REFERENCE's own code (container only: needs /root/reference
and `make -C oracle ref`).  Committed outputs let the oracle be checked where the reference
cannot travel.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402

ROOT = H.ROOT


def latmc_golden():
    """4. the reference's own test inputs: {s,d,c,z}latmc(n, c = 2) after srand(0) (what every reference
    LAPACK test factors, test/lapack/dpotrf.c:56) from the reference's generator, and the reference's
    CPU potrf of them, both triangles."""
    exe = os.path.join(H.ORACLE_DIR, "_ref", "pin_latmc")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-w", "-I/root/reference/test/lapack/util",
                           os.path.join(H.ORACLE_DIR, "pin_latmc.c"), "-o", exe, "-lm"])
    REF = H.CpuLib("ref")
    d = {}
    for p in "sdcz":
        for n in (2, 64):
            raw = subprocess.check_output([exe, p, str(n), "2.0", "0"])
            A = np.frombuffer(raw, dtype=H.PREC[p]["np"]).reshape(n, n).T
            d["%s_%d_A" % (p, n)] = np.ascontiguousarray(A)
            for uplo in "LU":
                F = H.fmat(n, n, H.PREC[p]["np"])
                F[:, :] = A
                assert REF.potrf(p, uplo, F) == 0
                d["%s_%d_%s" % (p, n, uplo)] = np.ascontiguousarray(H.tri(F, uplo))
    np.savez_compressed(os.path.join(HERE, "latmc.npz"), **d)


def dsfmt_kinds_golden(exe=None):
    """1b. every output form of the reference's dSFMT-19937 (get, getOpenClose, getOpenOpen, getCloseClose; getCloseOpen again)
    as bit patterns, from the start of the stream and after an odd-sized first call (which drops half a word)."""
    if exe is None:
        exe = os.path.join(H.ORACLE_DIR, "_ref", "pin_dsfmt")
        subprocess.check_call(["gcc", "-O2", "-msse2", "-w", "-I/root/reference/rng", os.path.join(H.ORACLE_DIR, "pin_dsfmt.c"), "-o", exe])
    d = {"seed": H.SEED, "source": "reference rng/dsfmt_19937.c: get / getCloseOpen / getOpenClose / getOpenOpen / getCloseClose",
         "n": 1001, "skip": 7, "kinds": {}}
    for name, kind in H.DSFMT_KINDS.items():
        if kind == 1:
            first = [np.float64(x).view(np.uint64).item() for x in map(float, subprocess.check_output([exe, str(H.SEED), "1001", "1"]).decode().split())]
            after = [np.float64(x).view(np.uint64).item() for x in map(float, subprocess.check_output([exe, str(H.SEED), "1001", "1", "7"]).decode().split())]
        else:
            first = [int(x, 16) for x in subprocess.check_output([exe, str(H.SEED), "1001", str(kind)]).decode().split()]
            after = [int(x, 16) for x in subprocess.check_output([exe, str(H.SEED), "1001", str(kind), "7"]).decode().split()]
        d["kinds"][name] = {"first": ["%016x" % v for v in first], "after_skip": ["%016x" % v for v in after]}
    with open(os.path.join(HERE, "dsfmt19937_kinds.json"), "w") as f:
        json.dump(d, f)


def main():
    subprocess.check_call(["make", "-C", H.ORACLE_DIR, "ref", "oracle"])
    if "--only-latmc" in sys.argv:
        latmc_golden()
        return
    if "--only-dsfmt-kinds" in sys.argv:
        dsfmt_kinds_golden()
        return
    # 1. dSFMT-19937: compile the reference's rng/dsfmt_19937.c behind oracle/pin_dsfmt.c
    exe = os.path.join(H.ORACLE_DIR, "_ref", "pin_dsfmt")
    subprocess.check_call(["gcc", "-O2", "-msse2", "-w", "-I/root/reference/rng",
                           os.path.join(H.ORACLE_DIR, "pin_dsfmt.c"), "-o", exe])
    out = subprocess.check_output([exe, str(H.SEED), "1000"]).decode().split()
    with open(os.path.join(HERE, "dsfmt19937_seed20121018.json"), "w") as f:
        json.dump({"seed": H.SEED, "source": "reference rng/dsfmt_19937.c getCloseOpen",
                   "close_open": [float(x) for x in out]}, f)
    dsfmt_kinds_golden(exe)
    # 2. potrf factors of the seeded SPD inputs, from the reference's CPU {s,d,c,z}potrf
    REF = H.CpuLib("ref")
    for p in "sdcz":
        d = {}
        for uplo in "LU":
            for n in (1, 5, 31, 32, 33, 100, 257):
                A = H.spd_matrix(p, n, ld=n + 3)
                assert REF.potrf(p, uplo, A) == 0
                d["%s_%d" % (uplo, n)] = np.ascontiguousarray(H.tri(A, uplo))
        np.savez_compressed(os.path.join(HERE, "potrf_%s.npz" % p), **d)
    # 3. potri + logdet of the n=100 case (d only, both uplo)
    d = {}
    for uplo in "LU":
        A = H.spd_matrix("d", 100)
        REF.potrf("d", uplo, A)
        d["logdet_" + uplo] = np.array(REF.logdet("d", A, A.strides[1] // 8 + 1, 100))
        REF.potri("d", uplo, A)
        d["potri_" + uplo] = np.ascontiguousarray(H.tri(A, uplo))
    np.savez_compressed(os.path.join(HERE, "potri_logdet_d.npz"), **d)
    latmc_golden()
    print("golden written to", HERE)


if __name__ == "__main__":
    main()

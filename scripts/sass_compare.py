"""sass_compare.py OLD.o NEW.o — per kernel, is the SASS of two objects the same instruction stream?

Used when a variant is added WITHOUT hardware access: the default kernels must stay bit-identical to
the build that passed the GPU parity tests (adding code next to a kernel can change its layout -
it did for the FP32 engine, which is why its deeper-k variant lives in its own translation unit).
    git show <verified commit>:cuda-cholesky_b200/csrc/x.cu > /tmp/r1/x.cu   (plus the headers)
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -c /tmp/r1/x.cu -o /tmp/r1/x.o
    python scripts/sass_compare.py /tmp/r1/x.o cuda-cholesky_b200/build/x.o
"""
import subprocess, re, sys
def funcs(obj):
    out = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
    d, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []; continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append(re.sub(r"\s+", " ", m.group(1)).strip())
    return d
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
for k in sorted(a):
    if k not in b:
        print("MISSING in new:", k); continue
    print("%-9s %6d instr  %s" % ("same" if a[k] == b[k] else "DIFFERENT", len(a[k]), k[:90]))
for k in sorted(set(b) - set(a)):
    print("new only  %6d instr  %s" % (len(b[k]), k[:90]))

/*
 * pin_latmc.c — TEST INFRASTRUCTURE ONLY; builds only in the container where /root/reference
 * exists.  Compiles the reference's own test-matrix generators test/lapack/util/{s,d,c,z}latmc.c
 * (static functions meant to be #included by its tests, e.g. test/lapack/dpotrf.c:56) and dumps
 * the n x n matrix latmc(n, c) after srand(seed) as raw column-major bytes on stdout, so that
 * tests/harness.py::latmc can be pinned to it (tests/golden/latmc.npz).
 *   gcc -std=c99 -O2 -I/root/reference/test/lapack/util oracle/pin_latmc.c -o oracle/_ref/pin_latmc -lm
 *   pin_latmc <s|d|c|z> <n> <c> <seed>
 */
#include <complex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "slatmc.c"
#include "dlatmc.c"
#include "clatmc.c"
#include "zlatmc.c"

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  const char p = argv[1][0];
  const size_t n = strtoull(argv[2], NULL, 10);
  const double c = atof(argv[3]);
  srand((unsigned)strtoul(argv[4], NULL, 10));
  const size_t es = p == 's' ? 4 : (p == 'z' ? 16 : 8);
  void *A = malloc(n * n * es);
  int rc = 1;
  if (p == 's') rc = slatmc(n, (float)c, A, n);
  if (p == 'd') rc = dlatmc(n, c, A, n);
  if (p == 'c') rc = clatmc(n, (float)c, A, n);
  if (p == 'z') rc = zlatmc(n, c, A, n);
  if (rc != 0) return 3;
  fwrite(A, es, n * n, stdout);
  return 0;
}

/*
 * pin_dsfmt.c — TEST INFRASTRUCTURE ONLY; builds only in the container where /root/reference
 * exists.  Compiles the reference's own rng/dsfmt_19937.c (which #includes rng/dsfmt.c) behind
 * a few stub typedefs (the reference's rng/ tree is unfinished and has no header for them,
 * SURVEY.md §2.5) and prints the first N doubles of getCloseOpen for a seed, so that
 * oracle/dsfmt_oracle.c can be pinned to the reference's generator bit for bit.
 *   gcc -O2 -msse2 -I/root/reference/rng oracle/pin_dsfmt.c -o oracle/_ref/pin_dsfmt
 *   pin_dsfmt SEED N [KIND [SKIP]]   KIND 1 getCloseOpen (default, printed as %.17g), 2 getOpenClose, 3 getOpenOpen,
 *                                    4 getCloseClose, 0 get - kinds 0, 2, 3, 4 are printed as 16 hex digits (the
 *                                    bit patterns); SKIP values are drawn with the same call first and thrown away.
 * The reference's `get` (rng/dsfmt.c:127-146) indexes its result wrongly after the loop and loses the last pair;
 * for kind 0 only the values its loop stored, 2 * ((N - 1) / 2) of them, are printed.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef struct { size_t n, inc; double *data; } vectord;
static inline void vectordSet(vectord *v, size_t i, double x) { v->data[i * v->inc] = x; }
typedef struct {
  const char *name; size_t size; uint64_t min, max;
  void (*set)(uint64_t, void *);
  void (*get)(uint64_t *, size_t, void *);
  void (*getOpenOpen)(vectord *, void *);
  void (*getOpenClose)(vectord *, void *);
  void (*getCloseOpen)(vectord *, void *);
  void (*getCloseClose)(vectord *, void *);
} rng64_t;
#include "dsfmt_19937.c"
int main(int argc, char **argv) {
  uint64_t seed = argc > 1 ? strtoull(argv[1], NULL, 10) : 20121018u;
  size_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 16;
  int kind = argc > 3 ? atoi(argv[3]) : 1;
  size_t skip = argc > 4 ? strtoull(argv[4], NULL, 10) : 0;
  void *st = malloc(type.size);
  double *out = calloc((n > skip ? n : skip) + 2, sizeof(double));
  type.set(seed, st);
  for (int pass = skip ? 0 : 1; pass < 2; pass++) {
    vectord v = {pass ? n : skip, 1, out};
    switch (kind) {
      case 0: type.get((uint64_t *)out, v.n, st); break;
      case 1: type.getCloseOpen(&v, st); break;
      case 2: type.getOpenClose(&v, st); break;
      case 3: type.getOpenOpen(&v, st); break;
      default: type.getCloseClose(&v, st); break;
    }
  }
  if (kind == 1) {
    for (size_t i = 0; i < n; i++) printf("%.17g\n", out[i]);
  } else {
    const size_t m = kind == 0 ? 2 * ((n - 1) / 2) : n;
    for (size_t i = 0; i < m; i++) printf("%016llx\n", (unsigned long long)((uint64_t *)out)[i]);
  }
  return 0;
}

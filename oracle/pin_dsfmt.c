/*
 * pin_dsfmt.c — TEST INFRASTRUCTURE ONLY; builds only in the container where /root/reference
 * exists.  Compiles the reference's own rng/dsfmt_19937.c (which #includes rng/dsfmt.c) behind
 * a few stub typedefs (the reference's rng/ tree is unfinished and has no header for them,
 * SURVEY.md §2.5) and prints the first N doubles of getCloseOpen for a seed, so that
 * oracle/dsfmt_oracle.c can be pinned to the reference's generator bit for bit.
 *   gcc -O2 -msse2 -I/root/reference/rng oracle/pin_dsfmt.c -o oracle/_ref/pin_dsfmt
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
  void *st = malloc(type.size);
  double *out = malloc(n * sizeof(double));
  vectord v = {n, 1, out};
  type.set(seed, st);
  type.getCloseOpen(&v, st);
  for (size_t i = 0; i < n; i++) printf("%.17g\n", out[i]);
  return 0;
}

/* TEST INFRASTRUCTURE ONLY - z-precision instantiation of the CPU oracle (see chol_oracle_impl.h). */
#include <complex.h>
#define SCALAR double complex
#define REAL double
#define IS_COMPLEX 1
#define X(name) oracle_z##name
#include "chol_oracle_impl.h"

/* TEST INFRASTRUCTURE ONLY - s-precision instantiation of the CPU oracle (see chol_oracle_impl.h). */
#include <complex.h>
#define SCALAR float
#define REAL float
#define IS_COMPLEX 0
#define X(name) oracle_s##name
#include "chol_oracle_impl.h"

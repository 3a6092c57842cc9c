/* TEST INFRASTRUCTURE ONLY - d-precision instantiation of the CPU oracle (see chol_oracle_impl.h). */
#include <complex.h>
#define SCALAR double
#define REAL double
#define IS_COMPLEX 0
#define X(name) oracle_d##name
#include "chol_oracle_impl.h"

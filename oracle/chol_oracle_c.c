/* TEST INFRASTRUCTURE ONLY - c-precision instantiation of the CPU oracle (see chol_oracle_impl.h). */
#include <complex.h>
#define SCALAR float complex
#define REAL float
#define IS_COMPLEX 1
#define X(name) oracle_c##name
#include "chol_oracle_impl.h"

/*
 * error.h — error-reporting convention of the library.
 * Replaces reference include/error.h:1-79 (same hook type, same early-return macros).
 * Difference: the reference declares its own `cuGetErrorString(CUresult)` (include/error.h:40),
 * which clashes with the CUDA driver's symbol since CUDA 6; here it is cuCholGetErrorString.
 */
#ifndef CHOL_B200_ERROR_H
#define CHOL_B200_ERROR_H

#include <stdbool.h>
#include <string.h>
#include <cuda.h>

#define STRINGx(x) #x
#define STRING(x) STRINGx(x)

#ifdef __cplusplus
extern "C" {
#endif

/* (call text, calling function, file, line, error code, error -> string) */
typedef void (*errorHandler_t)(const char *, const char *, const char *, int, int, const char *(*)(int));

/* Global hook invoked before an error is returned; may be set to NULL (reference
 * multigpu/error.c:4-11 prints to stderr by default, so does this library). */
extern errorHandler_t errorHandler;

/* Human-readable text for a CUresult (reference include/error.h:40, renamed). */
const char *cuCholGetErrorString(CUresult);

#define CU_ERROR_CHECK(call)                                                                     \
  do {                                                                                           \
    CUresult __error__;                                                                          \
    if ((__error__ = (call)) != CUDA_SUCCESS) {                                                  \
      if (errorHandler != NULL)                                                                  \
        errorHandler(STRING(call), __func__, __FILE__, __LINE__, __error__,                      \
                     (const char *(*)(int))cuCholGetErrorString);                                \
      return __error__;                                                                          \
    }                                                                                            \
  } while (false)

#define ERROR_CHECK(call)                                                                        \
  do {                                                                                           \
    int __error__;                                                                               \
    if ((__error__ = (call)) != 0) {                                                             \
      if (errorHandler != NULL)                                                                  \
        errorHandler(STRING(call), __func__, __FILE__, __LINE__, __error__,                      \
                     (const char *(*)(int))strerror);                                            \
      return CUDA_ERROR_OPERATING_SYSTEM;                                                        \
    }                                                                                            \
  } while (false)

#ifdef __cplusplus
}
#endif
#endif

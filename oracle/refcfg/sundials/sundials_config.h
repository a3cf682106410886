/* TEST INFRASTRUCTURE (oracle/_ref build) -- hand-written configuration header that stands in
 * for the file the SUNDIALS build system would generate from include/sundials/sundials_config.in.
 * Settings follow PelePhysics ThirdParty/GNUmakefile:169-196 (double precision, 32-bit index,
 * no MPI/OpenMP/CUDA, no fused kernels, no profiling). Not part of the product. */
#ifndef _SUNDIALS_CONFIG_H
#define _SUNDIALS_CONFIG_H
#include "sundials/sundials_export.h"
#ifndef SUNDIALS_DEPRECATED_MSG
#define SUNDIALS_DEPRECATED_MSG(msg) __attribute__((__deprecated__(msg)))
#endif
#ifndef SUNDIALS_DEPRECATED_EXPORT_MSG
#define SUNDIALS_DEPRECATED_EXPORT_MSG(msg) SUNDIALS_EXPORT SUNDIALS_DEPRECATED_MSG(msg)
#endif
#ifndef SUNDIALS_DEPRECATED_NO_EXPORT_MSG
#define SUNDIALS_DEPRECATED_NO_EXPORT_MSG(msg) SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED_MSG(msg)
#endif
#define SUNDIALS_VERSION "6.2.0"
#define SUNDIALS_VERSION_MAJOR 6
#define SUNDIALS_VERSION_MINOR 2
#define SUNDIALS_VERSION_PATCH 0
#define SUNDIALS_VERSION_LABEL ""
#define SUNDIALS_GIT_VERSION ""
#define SUNDIALS_DOUBLE_PRECISION 1
#define SUNDIALS_INT32_T 1
#define SUNDIALS_INDEX_TYPE int32_t
#define SUNDIALS_HAVE_POSIX_TIMERS
#define SUNDIALS_LOGGING_LEVEL 0
#define SUNDIALS_COMPILER_HAS_SNPRINTF_AND_VA_COPY
#define SUNDIALS_MPI_ENABLED 0
#define SUNDIALS_CVODE 1
#define SUNDIALS_NVECTOR_SERIAL 1
#define SUNDIALS_SUNMATRIX_DENSE 1
#define SUNDIALS_SUNMATRIX_SPARSE 1
#define SUNDIALS_SUNMATRIX_BAND 1
#define SUNDIALS_SUNLINSOL_DENSE 1
#define SUNDIALS_SUNLINSOL_SPGMR 1
#define SUNDIALS_SUNNONLINSOL_NEWTON 1
#define SUNDIALS_SUNNONLINSOL_FIXEDPOINT 1
#ifndef SUNDIALS_CXX_INLINE
#define SUNDIALS_CXX_INLINE inline
#endif
#ifndef SUNDIALS_C_INLINE
#ifndef __STDC_VERSION__
#define SUNDIALS_C_INLINE
#else
#define SUNDIALS_C_INLINE inline
#endif
#endif
#ifdef __cplusplus
#define SUNDIALS_INLINE SUNDIALS_CXX_INLINE
#else
#define SUNDIALS_INLINE SUNDIALS_C_INLINE
#endif
#define SUNDIALS_STATIC_INLINE static SUNDIALS_INLINE
#endif

// sm/macros.h -- drop-in for the reference's include/macros.h:1-33.
// Only the names user code can see are kept; the CPU tuning knobs
// (CHUNK_SIZE = OpenMP grain, PRAGMA_UNROLL) have no meaning on the device
// path and are defined for source compatibility only.
#pragma once

#if defined(_MSC_VER)
#define likely(x) (x)
#define unlikely(x) (x)
#define ALWAYS_INLINE __forceinline
#else
#define likely(x) __builtin_expect(!!(x), 1)
#define unlikely(x) __builtin_expect(!!(x), 0)
#define ALWAYS_INLINE __attribute__((always_inline)) inline
#endif

#define CHUNK_SIZE 1024 /* reference macros.h:16; unused by the GPU launcher */
#define PRAGMA_UNROLL(n)

/*
 * lpmp_synthetic.h -- specification of the counter-based synthetic potential generator used by
 * lpmp_grid_fill_synthetic ("host oracle and device generate identical values without a
 * transfer").  Shared verbatim by the CUDA library and the CPU oracle; plain integer arithmetic, so both
 * produce bit-identical values.  u01 has 24 random bits, hence is exact in float and in double.
 */
#ifndef LPMP_SYNTHETIC_H
#define LPMP_SYNTHETIC_H
#include <stdint.h>

#if defined(__CUDACC__)
#define LPMP_HD __host__ __device__ __forceinline__
#else
#define LPMP_HD static inline
#endif

enum { LPMP_SYNTH_STREAM_UNARY = 0, LPMP_SYNTH_STREAM_HORIZONTAL = 1, LPMP_SYNTH_STREAM_VERTICAL = 2 };

/* splitmix64 finaliser over (seed, stream, counter) */
LPMP_HD uint64_t lpmp_synth_bits(uint64_t seed, uint64_t stream, uint64_t counter) {
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (stream + 1) + 0xD1B54A32D192ED03ull * counter;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

/* uniform in [0,1) with 24 bits */
LPMP_HD double lpmp_synth_u01(uint64_t seed, uint64_t stream, uint64_t counter) {
  return (double)(lpmp_synth_bits(seed, stream, counter) >> 40) * (1.0 / 16777216.0);
}

/* uniform in [lo,hi), evaluated in double and then rounded to the storage type by the caller */
LPMP_HD double lpmp_synth_param(uint64_t seed, uint64_t stream, uint64_t counter, double lo, double hi) {
  return lo + (hi - lo) * lpmp_synth_u01(seed, stream, counter);
}

#endif /* LPMP_SYNTHETIC_H */

// Device-side building blocks shared by all kernels of the LP_MP hot path (sm_100a).
//
// Arithmetic contract (DESIGN.md "bit parity"): every + and * is a single IEEE rounding (the library is built
// with -fmad=false), min is exact and may be re-associated freely, sums into theta / a / b are performed in the
// reference's order.  This makes the state (theta, a, b, labels) bit-identical to the CPU oracle in both f32 and
// f64.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lpmp_b200.h"

namespace lpmp {

template <typename Real> struct RealTraits;
template <> struct RealTraits<float> {
  __host__ __device__ static float inf() { return HUGE_VALF; }
};
template <> struct RealTraits<double> {
  __host__ __device__ static double inf() { return HUGE_VAL; }
};

template <typename Real> __device__ __forceinline__ Real rmin(Real a, Real b) { return b < a ? b : a; }
template <> __device__ __forceinline__ float rmin<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double rmin<double>(double a, double b) { return fmin(a, b); }

// ---- sub-warp groups: G consecutive lanes (G a power of two <= 32) cooperate on one factor -------------------
__device__ __forceinline__ unsigned group_mask(int G) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned base = lane & ~(unsigned)(G - 1);
  return (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << base;
}

template <typename Real>
__device__ __forceinline__ Real group_min(Real v, int G, unsigned mask) {
  for (int o = G >> 1; o > 0; o >>= 1) v = rmin(v, __shfl_xor_sync(mask, v, o));
  return v;
}

// (value, index) argmin with ties resolved to the LOWER index (std::min_element, test/test_model.hxx:27-33)
template <typename Real>
__device__ __forceinline__ void group_argmin(Real& v, uint32_t& idx, int G, unsigned mask) {
  for (int o = G >> 1; o > 0; o >>= 1) {
    const Real ov = __shfl_xor_sync(mask, v, o);
    const uint32_t oi = __shfl_xor_sync(mask, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// two smallest values (with multiplicity) and one minimiser: vector::two_min, include/vector.hxx:348-443
template <typename Real>
struct TwoMin {
  Real m1, m2;
  uint32_t i1;
};
template <typename Real>
__device__ __forceinline__ void two_min_insert(TwoMin<Real>& t, Real v, uint32_t i) {
  if (v < t.m1) { t.m2 = t.m1; t.m1 = v; t.i1 = i; } else if (v < t.m2) { t.m2 = v; }
}
template <typename Real>
__device__ __forceinline__ void group_two_min(TwoMin<Real>& t, int G, unsigned mask) {
  for (int o = G >> 1; o > 0; o >>= 1) {
    const Real om1 = __shfl_xor_sync(mask, t.m1, o);
    const Real om2 = __shfl_xor_sync(mask, t.m2, o);
    const uint32_t oi1 = __shfl_xor_sync(mask, t.i1, o);
    // merged second minimum = min(larger of the two minima, both second minima)
    const Real hi = t.m1 < om1 ? om1 : t.m1;
    Real s = rmin(t.m2, om2);
    s = rmin(s, hi);
    if (om1 < t.m1 || (om1 == t.m1 && oi1 < t.i1)) { t.m1 = om1; t.i1 = oi1; }
    t.m2 = s;
  }
}

// ---- pairwise parameter access (mirrors Oracle::cost, oracle/lpmp_oracle.cpp) ---------------------------------
__host__ __device__ __forceinline__ uint32_t pairwise_param_count(int pw, uint32_t L) {
  return pw == LPMP_PAIRWISE_DENSE ? L * L : (pw == LPMP_PAIRWISE_POTTS ? 1u : 2u);
}

template <typename Real>
__device__ __forceinline__ Real pairwise_cost(int pw, const Real* __restrict__ p, uint32_t L, uint32_t x, uint32_t y) {
  if (pw == LPMP_PAIRWISE_DENSE) return p[(uint64_t)x * L + y];
  if (pw == LPMP_PAIRWISE_POTTS) return x == y ? Real(0) : p[0];
  const uint32_t d = x > y ? x - y : y - x;
  return rmin(Real(p[0] * Real(d)), p[1]);
}

// h(x) = min_y [ c(x,y) + other(y) ] for the labels x = lane, lane+G, ... of this group (x on `side`).
// `oth` is the other side's vector staged in shared memory (L entries, visible to the whole group); results go to
// hs[x] (shared, one writer per x).  Arithmetic of matrix::min1(v)/min2(v), include/vector.hxx:850-859,889-913;
// Potts through two_min (:348-443); truncated linear: all terms with slope*d >= trunc collapse to trunc + min.
template <typename Real>
__device__ __forceinline__ void group_min_plus(int pw, const Real* __restrict__ p, uint32_t L, uint32_t side,
                                               const Real* oth, Real* hs, int G, int gl, unsigned mask) {
  const Real INF = RealTraits<Real>::inf();
  if (pw == LPMP_PAIRWISE_POTTS) {
    TwoMin<Real> t{INF, INF, 0u};
    for (uint32_t y = gl; y < L; y += G) two_min_insert(t, oth[y], y);
    group_two_min(t, G, mask);
    const Real lambda = p[0];
    for (uint32_t x = gl; x < L; x += G) hs[x] = rmin(oth[x], Real(lambda + (x == t.i1 ? t.m2 : t.m1)));
    return;
  }
  if (pw == LPMP_PAIRWISE_TRUNCLIN && p[0] >= Real(0) && p[1] >= Real(0)) {
    const Real slope = p[0], trunc = p[1];
    Real g = INF;
    for (uint32_t y = gl; y < L; y += G) g = rmin(g, oth[y]);
    g = group_min(g, G, mask);
    const Real far = trunc + g;
    for (uint32_t x = gl; x < L; x += G) {
      Real m = rmin(far, oth[x]);
      for (uint32_t d = 1; d < L; ++d) {
        const Real cd = slope * Real(d);
        if (!(cd < trunc)) break;
        if (x >= d) m = rmin(m, Real(cd + oth[x - d]));
        if (x + d < L) m = rmin(m, Real(cd + oth[x + d]));
      }
      hs[x] = m;
    }
    return;
  }
  for (uint32_t x = gl; x < L; x += G) {
    Real m = INF;
    for (uint32_t y = 0; y < L; ++y) {
      const Real c = side == 0 ? pairwise_cost(pw, p, L, x, y) : pairwise_cost(pw, p, L, y, x);
      m = rmin(m, Real(c + oth[y]));
    }
    hs[x] = m;
  }
}

// sum over the CTA, valid in thread 0 (fixed tree: deterministic)
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (int)((blockDim.x + 31) >> 5) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

// smallest power of two >= min(L, 32)
__host__ __device__ inline int group_size_for(uint32_t L) {
  int g = 1;
  while (g < 32 && (uint32_t)g < L) g <<= 1;
  return g;
}

}  // namespace lpmp

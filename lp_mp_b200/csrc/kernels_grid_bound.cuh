// Lower bound of a grid LP in one streaming sweep over the node records (sm_100a).
//
// Replaces LP::LowerBound (/root/reference/include/LP_MP.h:1507-1518) = sum_i UnarySimplexFactor::LowerBound
// (min theta_i, test/simplex.cpp:8-11) + sum_e PairwiseSimplexFactor::LowerBound (min_x [a(x) + min_y (c(x,y) + b(y))],
// test/simplex.cpp:52-64) for the record layout of kernels_grid.cuh.  A lane group (G = L/4 lanes, 4 labels per lane)
// owns one node: its theta and the two edges it is the first variable of (R, D), whose second-variable vectors
// (bL of the right neighbour, bU of the neighbour below) are the only reads outside the node's own record.  Every
// record and every parameter is read exactly once: algorithmic bytes = s*(5*N*L + 4*N*P), HBM-bound.
// Partial sums are per CTA in double, combined by final_sum_kernel in a fixed order (deterministic bound).
#pragma once
#include "kernels_grid.cuh"

namespace lpmp {

template <typename Real>
struct GridBoundArgs {
  uint32_t H, W;
  const Real* rec;       // [H][W][5][L]
  const Real* par;       // [H][W][4][P]
  const Real* below_b;   // stripes: [W][L] second-variable side of the edges below the last row, or null
  double* partial;       // [gridDim.x]
};

template <typename Real> __device__ __forceinline__ void ldg4(Real (&v)[4], const Real* p);
template <> __device__ __forceinline__ void ldg4<float>(float (&v)[4], const float* p) {
  const float4 q = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void ldg4<double>(double (&v)[4], const double* p) {
  const double2 q0 = __ldg(reinterpret_cast<const double2*>(p)), q1 = __ldg(reinterpret_cast<const double2*>(p) + 1);
  v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
}

constexpr int kGridBoundWarps = 8;

template <typename Real, int PW, int L>
constexpr size_t grid_bound_smem() {
  return PW == LPMP_PAIRWISE_TRUNCLIN ? (size_t)kGridBoundWarps * 2 * (32 / (L / 4)) * 3 * L * sizeof(Real) : 0;
}

template <typename Real, int PW, int L, int HOPS>
__global__ void __launch_bounds__(32 * kGridBoundWarps) grid_bound_kernel(const GridBoundArgs<Real> A) {
  constexpr int G = L / 4, TG = 32 / G, P = ParamCount<PW>::value, REC = REC_VECS * L, PREC = PAR_EDGES * P;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "L/4 must be a power of two <= 32");
  extern __shared__ __align__(16) unsigned char bound_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane / G, gl = lane % G;
  uint32_t stage_a = 0, stage_b = 0;
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    Real* sc = reinterpret_cast<Real*>(bound_smem) + (size_t)warp * 2 * TG * 3 * L;
    for (int i = lane; i < 2 * TG * 3 * L; i += 32) sc[i] = RealTraits<Real>::inf();
    stage_a = smem_u32(sc) + (uint32_t)((2 * t) * 3 * L + L + 4 * gl) * S;
    stage_b = stage_a + 3 * L * S;
    __syncwarp();
  }
  const uint64_t N = (uint64_t)A.H * A.W;
  const uint64_t stride = (uint64_t)gridDim.x * kGridBoundWarps * TG;
  double acc = 0.0;
  for (uint64_t base = ((uint64_t)blockIdx.x * kGridBoundWarps + warp) * TG; base < N; base += stride) {
    const bool valid = base + t < N;
    const uint64_t i = valid ? base + t : N - 1;          // lanes past the end shadow the last node (shuffles stay full-warp)
    const uint32_t r = (uint32_t)(i / A.W), c = (uint32_t)(i % A.W);
    const Real* rec = A.rec + i * REC + 4 * gl;
    const Real* par = A.par + i * PREC;
    const bool hasR = c + 1 < A.W, hasD_in = r + 1 < A.H, hasD = hasD_in || A.below_b != nullptr;
    Real th[4], aR[4], aD[4], bR[4], bD[4], pR[P], pD[P], hR[4], hD[4];
    ldg4(th, rec);
    ldg4(aR, rec + REC_AR * L);
    ldg4(aD, rec + REC_AD * L);
#pragma unroll
    for (int k = 0; k < 4; ++k) { bR[k] = Real(0); bD[k] = Real(0); }
    if (hasR) ldg4(bR, rec + REC + REC_BL * L);
    if (hasD_in) ldg4(bD, rec + (uint64_t)A.W * REC + REC_BU * L);
    else if (hasD) ldg4(bD, A.below_b + (uint64_t)c * L + 4 * gl);
#pragma unroll
    for (int p = 0; p < P; ++p) { pR[p] = __ldg(par + PAR_R * P + p); pD[p] = __ldg(par + PAR_D * P + p); }
    lane_min_plus2<Real, PW, L, HOPS>(pR, bR, hR, pD, bD, hD, stage_a, stage_b, gl);
    Real mT = rmin(rmin(th[0], th[1]), rmin(th[2], th[3]));
    Real mR = rmin(rmin(Real(hR[0] + aR[0]), Real(hR[1] + aR[1])), rmin(Real(hR[2] + aR[2]), Real(hR[3] + aR[3])));
    Real mD = rmin(rmin(Real(hD[0] + aD[0]), Real(hD[1] + aD[1])), rmin(Real(hD[2] + aD[2]), Real(hD[3] + aD[3])));
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
      mT = rmin(mT, __shfl_xor_sync(FULL, mT, o));
      mR = rmin(mR, __shfl_xor_sync(FULL, mR, o));
      mD = rmin(mD, __shfl_xor_sync(FULL, mD, o));
    }
    if (valid && gl == 0) acc += (double)mT + (hasR ? (double)mR : 0.0) + (hasD ? (double)mD : 0.0);
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

}  // namespace lpmp

// Host-side launcher of the wavefront pass kernel (kernels_grid.cuh).  The kernel is instantiated per
// (Real, pairwise family, anisotropic?) in wavefront_inst_*.cu so that the translation units compile in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_grid.cuh"

namespace lpmp {

struct WavefrontTuning {
  int max_ctas = 0;     // 0 = one CTA (= one warp = one band at a time) per resident slot
  int ctas_per_sm = 0;  // 0 = as many as fit
};

// tile shape of the kernel instance for (Real, pw, L): rows per band, columns per tile, dynamic shared memory
struct WavefrontShape { int T, K; size_t smem; };

template <typename Real, int PW>
inline bool wavefront_shape_pw(uint32_t L, WavefrontShape& s) {
  switch (L) {
    case 8: s = {WavefrontCfg<Real, PW, 8>::T, WavefrontCfg<Real, PW, 8>::K, WavefrontCfg<Real, PW, 8>::SMEM_BYTES}; return true;
    case 16: s = {WavefrontCfg<Real, PW, 16>::T, WavefrontCfg<Real, PW, 16>::K, WavefrontCfg<Real, PW, 16>::SMEM_BYTES}; return true;
    case 32: s = {WavefrontCfg<Real, PW, 32>::T, WavefrontCfg<Real, PW, 32>::K, WavefrontCfg<Real, PW, 32>::SMEM_BYTES}; return true;
    case 64: s = {WavefrontCfg<Real, PW, 64>::T, WavefrontCfg<Real, PW, 64>::K, WavefrontCfg<Real, PW, 64>::SMEM_BYTES}; return true;
    default: return false;
  }
}
// false if no kernel instance exists for this shape (dense tables, other label counts: level schedule instead)
template <typename Real>
inline bool wavefront_shape(int pw, uint32_t L, WavefrontShape& s) {
  if (pw == LPMP_PAIRWISE_POTTS) return wavefront_shape_pw<Real, LPMP_PAIRWISE_POTTS>(L, s);
  if (pw == LPMP_PAIRWISE_TRUNCLIN) return wavefront_shape_pw<Real, LPMP_PAIRWISE_TRUNCLIN>(L, s);
  return false;
}

// maps: [0] records, box T rows (tile stores, anisotropic loads); [1] records, box T+1 rows (uniform-mode loads);
// [2] parameter records, box T rows.  hops: label-window half width / 4 of the truncated-linear kernel (0 = all).
template <typename Real, int PW, bool ANISO>
cudaError_t launch_family(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int hops, const WavefrontTuning& tune, int num_sms,
                          cudaStream_t stream);

}  // namespace lpmp

// Host-side launcher of the wavefront pass kernel (kernels_grid.cuh).  The kernel is instantiated per
// (Real, pairwise family, anisotropic?) in wavefront_inst_*.cu so that the translation units compile in parallel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_grid.cuh"
#include "kernels_grid_bound.cuh"
#include "kernels_grid_dense.cuh"

namespace lpmp {

struct WavefrontTuning {
  int max_ctas = 0;     // 0 = one CTA (= warps_per_cta bands at a time) per resident slot
  int ctas_per_sm = 0;  // 0 = as many as fit
  int tile_cols = 0;    // 0 = automatic, else 8 or 4
  int lean = -1;        // lean variant for the anisotropic passes (compute warp + two mover warps per band): 1 on, 0 off,
                        // -1 automatic (on where it fits; profiles/r01c_lean.txt)
  int cluster = 0;      // lean variant: CTAs per thread-block cluster (consecutive bands, hand-off through distributed shared
                        // memory); 0 = automatic (8), 1 = no clusters
  int level_group = 0;  // level kernel: lanes per node (power of two below the default pow2 >= L); 0 = default
  int warps_per_cta = 0;  // consecutive bands per CTA (<= 8 and what the shared memory holds); 0 or 1 = one band per CTA
  int split_uni = -1;   // ... and its instances for the uniform / damped-uniform modes: 0 off (one-warp kernel), otherwise on where they apply
  int split = -1;       // edge-split variant for the anisotropic passes (kernels_grid_split.cuh: one row per warp): 1 on, 0 off, -1 automatic
  int fused_bound = -1; // split variant, backward passes: accumulate LP::LowerBound on the fly: 1 on, 0 off, -1 automatic (bandwidth-bound grids)
  int rows = -1;        // one-label-per-lane variant for the anisotropic passes (kernels_grid_rows.cuh): 1 on, 0 off, -1 automatic
};

// tile shapes of the kernel instances for (Real, pw, L): rows per band; per tile width (8 / 4 columns) whether the
// instance exists and its dynamic shared memory
struct WavefrontShape {
  int T;
  bool fits8, fits4;
  size_t smem8, smem4;
  int default_k() const { return fits8 ? 8 : 4; }
};

template <typename Real, int PW, int L>
inline WavefrontShape wavefront_shape_l() {
  using C8 = WavefrontCfg<Real, PW, L, 8>;
  using C4 = WavefrontCfg<Real, PW, L, 4>;
  return WavefrontShape{C8::T, C8::FITS, C4::FITS, C8::SMEM_BYTES, C4::SMEM_BYTES};
}
template <typename Real, int PW>
inline bool wavefront_shape_pw(uint32_t L, WavefrontShape& s) {
  switch (L) {
    case 8: s = wavefront_shape_l<Real, PW, 8>(); break;
    case 16: s = wavefront_shape_l<Real, PW, 16>(); break;
    case 32: s = wavefront_shape_l<Real, PW, 32>(); break;
    case 64: s = wavefront_shape_l<Real, PW, 64>(); break;
    default: return false;
  }
  return s.fits8 || s.fits4;
}
// false if no kernel instance exists for this shape (dense tables, other label counts: level schedule instead)
template <typename Real>
inline bool wavefront_shape(int pw, uint32_t L, WavefrontShape& s) {
  if (pw == LPMP_PAIRWISE_POTTS) return wavefront_shape_pw<Real, LPMP_PAIRWISE_POTTS>(L, s);
  if (pw == LPMP_PAIRWISE_TRUNCLIN) return wavefront_shape_pw<Real, LPMP_PAIRWISE_TRUNCLIN>(L, s);
  return false;
}

// maps: [0] records, box T rows (tile stores, anisotropic loads); [1] records, box T+1 rows (uniform-mode loads);
// [2] parameter records, box T rows -- all with `ksel` columns per box.  hops: label-window half width / 4 of the
// truncated-linear kernel (0 = all distances).  ksel = 4 exists for the anisotropic kernels (and for the uniform
// ones where 8 does not fit).
template <typename Real, int PW, bool ANISO>
cudaError_t launch_family(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int hops, int ksel, const WavefrontTuning& tune, int num_sms,
                          cudaStream_t stream);

// one-label-per-lane variant (kernels_grid_rows.cuh): anisotropic passes, L in {8, 16, 32}.  maps: [0] records, [1] parameter
// records, boxes of rows_band_rows(L) / 4 rows x 8 columns.  win: truncated linear: label window (host: slope * (win + 1) >=
// trunc for every edge, win <= 8); Potts: 1 = no negative lambda seen, 0 = general.
inline int rows_band_rows(uint32_t L) { return L == 8 || L == 16 || L == 32 ? 4 * (32 / (int)L) : 0; }
template <typename Real>
inline bool rows_supported(int pw, uint32_t L) {
  if (pw != LPMP_PAIRWISE_POTTS && pw != LPMP_PAIRWISE_TRUNCLIN) return false;
  return rows_band_rows(L) != 0;
}
template <typename Real, int PW>
cudaError_t launch_rows(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream);

// edge-split variant (kernels_grid_split.cuh): anisotropic passes, bands of 8 rows, L in {16, 32} where the shared memory holds it
// (truncated linear: L = 16).  maps: [0] records, [1] parameter records, boxes of 1 row x 8 columns.  win as for launch_rows.
constexpr int kSplitBandRows = 8;
template <typename Real, int PW> bool split_fits(uint32_t L);
template <typename Real>
inline bool split_supported(int pw, uint32_t L) {
  if (pw == LPMP_PAIRWISE_POTTS) return split_fits<Real, LPMP_PAIRWISE_POTTS>(L);
  if (pw == LPMP_PAIRWISE_TRUNCLIN) return split_fits<Real, LPMP_PAIRWISE_TRUNCLIN>(L);
  return false;
}
template <typename Real, int PW>
cudaError_t launch_split(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream);

// ... and its instances for the uniform / damped-uniform modes (grids with W % 8 == 0, W >= 16, no stripes)
template <typename Real, int PW> bool split_uni_fits(uint32_t L);
template <typename Real>
inline bool split_uni_supported(int pw, uint32_t L) {
  if (pw == LPMP_PAIRWISE_POTTS) return split_uni_fits<Real, LPMP_PAIRWISE_POTTS>(L);
  if (pw == LPMP_PAIRWISE_TRUNCLIN) return split_uni_fits<Real, LPMP_PAIRWISE_TRUNCLIN>(L);
  return false;
}
template <typename Real, int PW>
cudaError_t launch_split_uni(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream);

// dense-table grids (kernels_grid_dense.cuh): anisotropic passes for L in {16, 32, 64}
template <typename Real>
inline bool dense_wavefront_supported(uint32_t L) {
  switch (L) {
    case 16: return DenseCfg<Real, 16>::FITS;
    case 32: return DenseCfg<Real, 32>::FITS;
    case 64: return DenseCfg<Real, 64>::FITS;
    default: return false;
  }
}
template <typename Real>
cudaError_t launch_dense_pass(const DenseGridArgs<Real>& args, uint32_t L, const WavefrontTuning& tune, int num_sms, cudaStream_t stream);

// one streaming sweep over the node records (kernels_grid_bound.cuh); same (pw, L) coverage as the wavefront kernels.
// Writes one partial sum per CTA into args.partial and returns the number of CTAs through *ctas.
template <typename Real, int PW>
cudaError_t launch_grid_bound(const GridBoundArgs<Real>& args, uint32_t L, int hops, int max_ctas, int* ctas, cudaStream_t stream);

}  // namespace lpmp

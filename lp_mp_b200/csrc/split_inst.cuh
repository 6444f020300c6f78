// Included by split_inst_*.cu with LPMP_INST_REAL and LPMP_INST_PW defined: instances of wavefront_split_kernel
// (kernels_grid_split.cuh) and their launcher.
#include "wavefront_launch.h"
#include "kernels_grid_split.cuh"

namespace lpmp {

template <typename Real, int PW, int L, int WIN>
static cudaError_t launch_split_one(const CUtensorMap* maps, const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  using Cfg = SplitCfg<Real, PW, L>;
  if constexpr (!Cfg::FITS) {
    return cudaErrorInvalidValue;
  } else {
    auto kernel = args.forward ? wavefront_split_kernel<Real, PW, L, WIN, true> : wavefront_split_kernel<Real, PW, L, WIN, false>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64][2] = {};
    if (!configured[dev & 63][args.forward ? 1 : 0]) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
      configured[dev & 63][args.forward ? 1 : 0] = true;
    }
    int ctas = (int)args.num_bands;
    if (ctas > num_sms) ctas = num_sms;   // one CTA per SM; persistent: the ticket loop covers the remaining bands
    if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
    kernel<<<ctas, Cfg::THREADS, Cfg::SMEM, stream>>>(args);
    return cudaGetLastError();
  }
}

template <typename Real, int PW, int L>
static cudaError_t launch_split_win(const CUtensorMap* maps, const GridArgs<Real>& args, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    if constexpr (L == 16) {
      if (win <= 3) return launch_split_one<Real, PW, L, 3>(maps, args, tune, num_sms, stream);
      if (win <= 4) return launch_split_one<Real, PW, L, 4>(maps, args, tune, num_sms, stream);
      if (win <= 8) return launch_split_one<Real, PW, L, 8>(maps, args, tune, num_sms, stream);
    }
    return cudaErrorInvalidValue;
  } else {
    return win ? launch_split_one<Real, PW, L, 1>(maps, args, tune, num_sms, stream) : launch_split_one<Real, PW, L, 0>(maps, args, tune, num_sms, stream);
  }
}

template <typename Real, int PW>
cudaError_t launch_split(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  switch (L) {
    case 16: return launch_split_win<Real, PW, 16>(maps, args, win, tune, num_sms, stream);
    case 32: return launch_split_win<Real, PW, 32>(maps, args, win, tune, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename Real, int PW>
bool split_fits(uint32_t L) {
  switch (L) {
    case 16: return SplitCfg<Real, PW, 16>::FITS;
    case 32: return SplitCfg<Real, PW, 32>::FITS;
    default: return false;
  }
}

template cudaError_t launch_split<LPMP_INST_REAL, LPMP_INST_PW>(const CUtensorMap*, const GridArgs<LPMP_INST_REAL>&, uint32_t, int, const WavefrontTuning&, int, cudaStream_t);
template bool split_fits<LPMP_INST_REAL, LPMP_INST_PW>(uint32_t);

}  // namespace lpmp

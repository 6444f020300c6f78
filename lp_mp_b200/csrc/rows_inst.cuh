// Included by rows_inst_*.cu with LPMP_INST_REAL and LPMP_INST_PW defined: instances of wavefront_rows_kernel
// (kernels_grid_rows.cuh) and their launcher.
#include "wavefront_launch.h"
#include "kernels_grid_rows.cuh"

namespace lpmp {

template <typename Real, int PW, int L, int WIN>
static cudaError_t launch_rows_one(const CUtensorMap* maps, const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  using Cfg = RowsCfg<Real, PW, L>;
  if constexpr (!Cfg::FITS) {
    return cudaErrorInvalidValue;
  } else {
    auto kernel = args.forward ? wavefront_rows_kernel<Real, PW, L, WIN, true> : wavefront_rows_kernel<Real, PW, L, WIN, false>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64][2] = {};
    if (!configured[dev & 63][args.forward ? 1 : 0]) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
      configured[dev & 63][args.forward ? 1 : 0] = true;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::THREADS, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
    int ctas = (int)args.num_bands;
    if (ctas > per_sm * num_sms) ctas = per_sm * num_sms;   // persistent: the ticket loop covers the remaining bands
    if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
    kernel<<<ctas, Cfg::THREADS, Cfg::SMEM, stream>>>(maps[0], maps[1], args);
    return cudaGetLastError();
  }
}

template <typename Real, int PW, int L>
static cudaError_t launch_rows_win(const CUtensorMap* maps, const GridArgs<Real>& args, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    if (win <= 3) return launch_rows_one<Real, PW, L, 3>(maps, args, tune, num_sms, stream);
    if (win <= 4) return launch_rows_one<Real, PW, L, 4>(maps, args, tune, num_sms, stream);
    if constexpr (L >= 16) { if (win <= 8) return launch_rows_one<Real, PW, L, 8>(maps, args, tune, num_sms, stream); }
    return cudaErrorInvalidValue;
  } else {
    return win ? launch_rows_one<Real, PW, L, 1>(maps, args, tune, num_sms, stream) : launch_rows_one<Real, PW, L, 0>(maps, args, tune, num_sms, stream);
  }
}

template <typename Real, int PW>
cudaError_t launch_rows(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  switch (L) {
    case 8: return launch_rows_win<Real, PW, 8>(maps, args, win, tune, num_sms, stream);
    case 16: return launch_rows_win<Real, PW, 16>(maps, args, win, tune, num_sms, stream);
    case 32: return launch_rows_win<Real, PW, 32>(maps, args, win, tune, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template cudaError_t launch_rows<LPMP_INST_REAL, LPMP_INST_PW>(const CUtensorMap*, const GridArgs<LPMP_INST_REAL>&, uint32_t, int, const WavefrontTuning&, int, cudaStream_t);

}  // namespace lpmp

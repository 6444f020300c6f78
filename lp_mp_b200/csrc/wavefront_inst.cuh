// Included by wavefront_inst_*.cu with LPMP_INST_REAL and LPMP_INST_PW defined.
#include "wavefront_launch.h"

namespace lpmp {

template <typename Real, int PW, int L, int VPL, int STAGES>
static cudaError_t launch_one(const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  using Cfg = WavefrontCfg<Real, PW, L, VPL>;
  auto kernel = wavefront_pass_kernel<Real, PW, L, VPL, STAGES>;
  const int warps = tune.warps < 1 ? 1 : (tune.warps > 4 ? 4 : tune.warps);
  const size_t smem = (size_t)warps * (STAGES * Cfg::STAGE_REALS + Cfg::SCRATCH_REALS) * sizeof(Real);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  int ctas = (int)((args.num_bands + warps - 1) / warps);
  const int resident = per_sm * num_sms;
  if (ctas > resident) ctas = resident;   // persistent: the ticket loop covers the remaining bands
  if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
  kernel<<<ctas, warps * 32, smem, stream>>>(args);
  return cudaGetLastError();
}

template <typename Real, int PW, int L, int VPL>
static cudaError_t launch_stages(const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  if (tune.stages == 4) return launch_one<Real, PW, L, VPL, 4>(args, tune, num_sms, stream);
  if (tune.stages == 8) return launch_one<Real, PW, L, VPL, 8>(args, tune, num_sms, stream);
  return cudaErrorInvalidValue;
}

template <typename Real, int PW>
cudaError_t launch_family(const GridArgs<Real>& args, uint32_t L, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  const int v = tune.vpl;
  if (L == 8 && v == 4) return launch_stages<Real, PW, 8, 4>(args, tune, num_sms, stream);
  if (L == 16 && v == 4) return launch_stages<Real, PW, 16, 4>(args, tune, num_sms, stream);
  if (L == 16 && v == 2) return launch_stages<Real, PW, 16, 2>(args, tune, num_sms, stream);
  if (L == 32 && v == 4) return launch_stages<Real, PW, 32, 4>(args, tune, num_sms, stream);
  if (L == 32 && v == 2) return launch_stages<Real, PW, 32, 2>(args, tune, num_sms, stream);
  if (L == 64 && v == 4) return launch_stages<Real, PW, 64, 4>(args, tune, num_sms, stream);
  return cudaErrorInvalidValue;
}

template cudaError_t launch_family<LPMP_INST_REAL, LPMP_INST_PW>(const GridArgs<LPMP_INST_REAL>&, uint32_t, const WavefrontTuning&, int, cudaStream_t);

}  // namespace lpmp

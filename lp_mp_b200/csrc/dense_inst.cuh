// Included by dense_inst_*.cu with LPMP_INST_REAL defined: the dense-table row-dataflow pass kernels.
#include "wavefront_launch.h"

namespace lpmp {

template <typename Real, int L>
static cudaError_t launch_dense_one(const DenseGridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  using Cfg = DenseCfg<Real, L>;
  if constexpr (!Cfg::FITS) {
    return cudaErrorInvalidValue;
  } else {
    auto kernel = dense_row_pass_kernel<Real, L>;
    const size_t smem = Cfg::SMEM_BYTES;
    static bool configured[64] = {};   // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[dev & 63] = true;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Cfg::NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
    int ctas = (int)args.H;
    const int resident = per_sm * num_sms;   // persistent: every CTA must be resident (rows wait for the row above)
    if (ctas > resident) ctas = resident;
    if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
    kernel<<<ctas, Cfg::NT, smem, stream>>>(args);
    return cudaGetLastError();
  }
}

template <typename Real>
cudaError_t launch_dense_pass(const DenseGridArgs<Real>& args, uint32_t L, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  switch (L) {
    case 16: return launch_dense_one<Real, 16>(args, tune, num_sms, stream);
    case 32: return launch_dense_one<Real, 32>(args, tune, num_sms, stream);
    case 64: return launch_dense_one<Real, 64>(args, tune, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template cudaError_t launch_dense_pass<LPMP_INST_REAL>(const DenseGridArgs<LPMP_INST_REAL>&, uint32_t, const WavefrontTuning&, int, cudaStream_t);

}  // namespace lpmp

// Included by split_inst_*.cu / split_uni_inst_*.cu with LPMP_INST_REAL and LPMP_INST_PW defined: instances of wavefront_split_kernel
// (kernels_grid_split.cuh) and their launcher; with LPMP_INST_UNI the instances for the uniform / damped-uniform modes.
#include "wavefront_launch.h"
#include "kernels_grid_split.cuh"

namespace lpmp {

template <typename Real, int PW, int L, int WIN>
static cudaError_t launch_split_one(const CUtensorMap* maps, const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
#ifdef LPMP_INST_UNI
  using Cfg = SplitCfg<Real, PW, L, true>;
#else
  using Cfg = SplitCfg<Real, PW, L>;
#endif
  if constexpr (!Cfg::FITS) {
    return cudaErrorInvalidValue;
  } else {
#ifdef LPMP_INST_UNI
    if (args.lb_rows || args.W % Cfg::K != 0 || args.W < 2 * Cfg::K) return cudaErrorInvalidValue;
    // stripes: whole bands only, and the ghost row next to the last virtual row holds the neighbouring stripe's vectors
    if ((args.in_halo || args.out_halo) && args.H % Cfg::T != 0) return cudaErrorInvalidValue;
    if (args.out_halo && args.later_boundary == nullptr) return cudaErrorInvalidValue;
    const int variant = (args.primal ? 2 : 0) + (args.forward ? 1 : 0);
    auto kernel = args.primal ? (args.forward ? wavefront_split_kernel<Real, PW, L, WIN, true, false, true, true> : wavefront_split_kernel<Real, PW, L, WIN, false, false, true, true>)
                              : (args.forward ? wavefront_split_kernel<Real, PW, L, WIN, true, false, false, true> : wavefront_split_kernel<Real, PW, L, WIN, false, false, false, true>);
#else
    if (args.forward && args.lb_rows) return cudaErrorInvalidValue;   // the fused bound belongs to the backward pass
    if (args.primal && args.lb_rows) return cudaErrorInvalidValue;    // rounding passes leave the bound to the sweep
    const int variant = args.primal ? (args.forward ? 3 : 4) : (args.forward ? 1 : (args.lb_rows ? 2 : 0));
    auto kernel = args.primal ? (args.forward ? wavefront_split_kernel<Real, PW, L, WIN, true, false, true> : wavefront_split_kernel<Real, PW, L, WIN, false, false, true>)
                  : args.forward ? wavefront_split_kernel<Real, PW, L, WIN, true, false, false>
                                 : (args.lb_rows ? wavefront_split_kernel<Real, PW, L, WIN, false, true, false> : wavefront_split_kernel<Real, PW, L, WIN, false, false, false>);
#endif
    int dev = 0;
    cudaGetDevice(&dev);
    static bool configured[64][5] = {};
    if (!configured[dev & 63][variant]) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
      configured[dev & 63][variant] = true;
    }
    int ctas = (int)args.num_bands;
    if (ctas > num_sms) ctas = num_sms;   // one CTA per SM; persistent: the ticket loop covers the remaining bands
    if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
    // thread-block clusters: the CTAs of a cluster take consecutive bands and hand off through distributed shared memory
    // (8 where every band has its own CTA -- the pass is a latency chain; 4 for persistent launches: 37 x 4 CTAs fill the 148 SMs,
    // 18 x 8 do not -- measured C2 0.47 / 0.49 ms, C3 4.17 / 3.73 ms per pass with clusters of 8 / 4)
    int cs = tune.cluster > 0 ? tune.cluster : ((int)args.num_bands > num_sms ? 4 : 8);
    if (cs > 8) cs = 8;
    while (cs > 1 && ((cs & (cs - 1)) != 0 || cs > ctas)) --cs;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (; cs > 1; cs >>= 1) {   // the largest cluster size the device can co-schedule
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)cs);
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) == cudaSuccess && max_clusters >= 1) {
        int clusters = (ctas + cs - 1) / cs;
        if (clusters > max_clusters) clusters = max_clusters;
        if (clusters * cs > num_sms) clusters = num_sms / cs;
        if (tune.max_ctas > 0 && clusters * cs > tune.max_ctas) clusters = tune.max_ctas / cs;
        // a latency-bound launch (every band could have its own SM) whose bands do not all fit with this cluster size: band
        // `resident` can only start when the CTA of band 0 has finished its W columns, long after the chain has reached it when
        // the grid is much wider than resident * 8 rows (1024 x 4096 stripes: 120 of 128 bands resident with clusters of 8 cost
        // 0.7 ms per pass) -- take the next smaller cluster size, which co-schedules more CTAs
        const int resident = clusters * cs;
        if (tune.cluster <= 0 && cs > 2 && (int)args.num_bands <= num_sms && resident < (int)args.num_bands &&
            (long long)args.W > 2LL * resident * kSplitBandRows) { (void)cudaGetLastError(); continue; }
        if (clusters >= 1) { ctas = clusters * cs; break; }
      }
      (void)cudaGetLastError();
    }
    if (cs <= 1) {
      kernel<<<ctas, Cfg::THREADS, Cfg::SMEM, stream>>>(args);
      return cudaGetLastError();
    }
    cfg.gridDim = dim3((unsigned)ctas);
    return cudaLaunchKernelEx(&cfg, kernel, args);
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

#ifdef LPMP_INST_UNI
#define LPMP_SPLIT_LAUNCH launch_split_uni
#define LPMP_SPLIT_FITS split_uni_fits
#else
#define LPMP_SPLIT_LAUNCH launch_split
#define LPMP_SPLIT_FITS split_fits
#endif

template <typename Real, int PW>
cudaError_t LPMP_SPLIT_LAUNCH(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int win, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  switch (L) {
    case 16: return launch_split_win<Real, PW, 16>(maps, args, win, tune, num_sms, stream);
    case 32: return launch_split_win<Real, PW, 32>(maps, args, win, tune, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename Real, int PW>
bool LPMP_SPLIT_FITS(uint32_t L) {
#ifdef LPMP_INST_UNI
  switch (L) {
    case 16: return SplitCfg<Real, PW, 16, true>::FITS;
    case 32: return SplitCfg<Real, PW, 32, true>::FITS;
    default: return false;
  }
#else
  switch (L) {
    case 16: return SplitCfg<Real, PW, 16>::FITS;
    case 32: return SplitCfg<Real, PW, 32>::FITS;
    default: return false;
  }
#endif
}

template cudaError_t LPMP_SPLIT_LAUNCH<LPMP_INST_REAL, LPMP_INST_PW>(const CUtensorMap*, const GridArgs<LPMP_INST_REAL>&, uint32_t, int, const WavefrontTuning&, int, cudaStream_t);
template bool LPMP_SPLIT_FITS<LPMP_INST_REAL, LPMP_INST_PW>(uint32_t);

}  // namespace lpmp

// Included by wavefront_inst_*.cu with LPMP_INST_REAL, LPMP_INST_PW and LPMP_INST_ANISO defined.
#include "wavefront_launch.h"
#include "kernels_grid_lean.cuh"

namespace lpmp {

template <typename Real, int PW, int L, int HOPS, bool ANISO, int KSEL>
static cudaError_t launch_one(const CUtensorMap* maps, const GridArgs<Real>& args, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  using Cfg = WavefrontCfg<Real, PW, L, KSEL>;
  // the narrow tile exists for the anisotropic kernels, and for the uniform ones only where the wide tile does not fit
  constexpr bool EXISTS = Cfg::FITS;
  if constexpr (!EXISTS) {
    return cudaErrorInvalidValue;
  } else {
    int dev = 0;
    cudaGetDevice(&dev);
    int smem_cap = 0;
    cudaError_t e = cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if constexpr (ANISO && LeanCfg<Real, PW, L, KSEL>::LEAN_FITS) {
      if (tune.lean != 0 && tune.warps_per_cta <= 1) {
        // lean variant: compute warp + in-mover + out-mover + tile warp per band (kernels_grid_lean.cuh)
        using LCfg = LeanCfg<Real, PW, L, KSEL>;
        auto lkernel = wavefront_lean_kernel<Real, PW, L, HOPS, KSEL>;
        static bool lconfigured[64] = {};
        if (!lconfigured[dev & 63]) {
          e = cudaFuncSetAttribute(lkernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LCfg::LEAN_SMEM);
          if (e != cudaSuccess) return e;
          lconfigured[dev & 63] = true;
        }
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lkernel, 128, LCfg::LEAN_SMEM);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
        int ctas = (int)args.num_bands;
        if (ctas > per_sm * num_sms) ctas = per_sm * num_sms;   // persistent: the ticket loop covers the remaining bands
        if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
        // thread-block clusters: the CTAs of a cluster take consecutive bands and hand off through distributed shared memory
        int cs = tune.cluster > 0 ? tune.cluster : 8;
        if (cs > 8) cs = 8;
        while (cs > 1 && ((cs & (cs - 1)) != 0 || cs > ctas || (uint32_t)cs > args.num_bands)) --cs;
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = LCfg::LEAN_SMEM;
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        for (; cs > 1; cs >>= 1) {   // the largest cluster size the device can co-schedule
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.gridDim = dim3((unsigned)cs);
          int max_clusters = 0;
          if (cudaOccupancyMaxActiveClusters(&max_clusters, lkernel, &cfg) == cudaSuccess && max_clusters >= 1) {
            int clusters = ctas / cs;
            if (tune.ctas_per_sm > 0 && clusters > per_sm * num_sms / cs) clusters = per_sm * num_sms / cs;
            if (clusters > max_clusters) clusters = max_clusters;
            const int want = (int)((args.num_bands + cs - 1) / cs);
            if (clusters > want) clusters = want;
            if (clusters >= 1) { ctas = clusters * cs; break; }
          }
          (void)cudaGetLastError();
        }
        if (cs <= 1) {
          lkernel<<<ctas, 128, LCfg::LEAN_SMEM, stream>>>(maps[0], maps[2], args);
          return cudaGetLastError();
        }
        cfg.gridDim = dim3((unsigned)ctas);
        return cudaLaunchKernelEx(&cfg, lkernel, maps[0], maps[2], args);
      }
    }
    // warps (= consecutive bands) per CTA: at most what the shared memory holds and MAX_WARPS; the hand-off between the
    // bands of a CTA stays in shared memory
    int nw_max = 1;
    while (nw_max < Cfg::MAX_WARPS && Cfg::cta_smem(nw_max + 1) <= (size_t)smem_cap) ++nw_max;
    // measured (profiles/r01b_superbands.txt): the ring kernel's longer step outweighs the cheaper hand-off on C2 / C3,
    // so one band per CTA stays the default and super-bands are opt-in ("wavefront_warps_per_cta")
    const int nw = tune.warps_per_cta > 1 ? (tune.warps_per_cta < nw_max ? tune.warps_per_cta : nw_max) : 1;
    auto kernel = nw > 1 ? wavefront_pass_kernel<Real, PW, L, HOPS, ANISO, KSEL, true> : wavefront_pass_kernel<Real, PW, L, HOPS, ANISO, KSEL, false>;
    static bool configured[64][2] = {};   // the attribute is per device and kernel
    if (!configured[dev & 63][nw > 1]) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
      if (e != cudaSuccess) return e;
      configured[dev & 63][nw > 1] = true;
    }
    const size_t smem = Cfg::cta_smem(nw);
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * nw, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (tune.ctas_per_sm > 0 && per_sm > tune.ctas_per_sm) per_sm = tune.ctas_per_sm;
    int ctas = (int)((args.num_bands + nw - 1) / nw);
    const int resident = per_sm * num_sms;
    if (ctas > resident) ctas = resident;   // persistent: the ticket loop covers the remaining super-bands
    if (tune.max_ctas > 0 && ctas > tune.max_ctas) ctas = tune.max_ctas;
    kernel<<<ctas, 32 * nw, smem, stream>>>(maps[0], maps[1], maps[2], args);
    return cudaGetLastError();
  }
}

template <typename Real, int PW, int L, int HOPS, bool ANISO>
static cudaError_t launch_k(const CUtensorMap* maps, const GridArgs<Real>& args, int ksel, const WavefrontTuning& tune, int num_sms, cudaStream_t stream) {
  if (ksel == 8) return launch_one<Real, PW, L, HOPS, ANISO, 8>(maps, args, tune, num_sms, stream);
  if (ksel == 4) return launch_one<Real, PW, L, HOPS, ANISO, 4>(maps, args, tune, num_sms, stream);
  return cudaErrorInvalidValue;
}

template <typename Real, int PW, int L, bool ANISO>
static cudaError_t launch_hops(const CUtensorMap* maps, const GridArgs<Real>& args, int hops, int ksel, const WavefrontTuning& tune, int num_sms,
                               cudaStream_t stream) {
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    if (hops == 1) return launch_k<Real, PW, L, 1, ANISO>(maps, args, ksel, tune, num_sms, stream);
    if (hops == 2 && L >= 16) return launch_k<Real, PW, L, 2, ANISO>(maps, args, ksel, tune, num_sms, stream);
  } else if constexpr (ANISO) {
    // Potts, anisotropic passes: the instance without the repulsive (two smallest entries) branch when no lambda is negative
    if (args.potts_nonneg) return launch_k<Real, PW, L, 1, ANISO>(maps, args, ksel, tune, num_sms, stream);
  }
  return launch_k<Real, PW, L, 0, ANISO>(maps, args, ksel, tune, num_sms, stream);
}

template <typename Real, int PW, bool ANISO>
cudaError_t launch_family(const CUtensorMap* maps, const GridArgs<Real>& args, uint32_t L, int hops, int ksel, const WavefrontTuning& tune, int num_sms,
                          cudaStream_t stream) {
  switch (L) {
    case 8: return launch_hops<Real, PW, 8, ANISO>(maps, args, hops, ksel, tune, num_sms, stream);
    case 16: return launch_hops<Real, PW, 16, ANISO>(maps, args, hops, ksel, tune, num_sms, stream);
    case 32: return launch_hops<Real, PW, 32, ANISO>(maps, args, hops, ksel, tune, num_sms, stream);
    case 64: return launch_hops<Real, PW, 64, ANISO>(maps, args, hops, ksel, tune, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

#if LPMP_INST_ANISO
template <typename Real, int PW, int L, int HOPS>
static cudaError_t launch_bound_one(const GridBoundArgs<Real>& args, int max_ctas, int* ctas_out, cudaStream_t stream) {
  auto kernel = grid_bound_kernel<Real, PW, L, HOPS>;
  constexpr size_t smem = grid_bound_smem<Real, PW, L>();
  static bool configured[64] = {};   // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63] && smem > 40 * 1024) {   // static shared memory (block_sum) counts towards the 48 KiB default limit
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev & 63] = true;
  }
  constexpr int nodes_per_cta = kGridBoundWarps * (32 / (L / 4));
  const uint64_t N = (uint64_t)args.H * args.W;
  uint64_t ctas = (N + nodes_per_cta - 1) / nodes_per_cta;
  if (ctas > (uint64_t)max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  *ctas_out = (int)ctas;
  kernel<<<(int)ctas, 32 * kGridBoundWarps, smem, stream>>>(args);
  return cudaGetLastError();
}
template <typename Real, int PW, int L>
static cudaError_t launch_bound_hops(const GridBoundArgs<Real>& args, int hops, int max_ctas, int* ctas, cudaStream_t stream) {
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    if (hops == 1) return launch_bound_one<Real, PW, L, 1>(args, max_ctas, ctas, stream);
    if (hops == 2 && L >= 16) return launch_bound_one<Real, PW, L, 2>(args, max_ctas, ctas, stream);
  }
  return launch_bound_one<Real, PW, L, 0>(args, max_ctas, ctas, stream);
}
template <typename Real, int PW>
cudaError_t launch_grid_bound(const GridBoundArgs<Real>& args, uint32_t L, int hops, int max_ctas, int* ctas, cudaStream_t stream) {
  switch (L) {
    case 8: return launch_bound_hops<Real, PW, 8>(args, hops, max_ctas, ctas, stream);
    case 16: return launch_bound_hops<Real, PW, 16>(args, hops, max_ctas, ctas, stream);
    case 32: return launch_bound_hops<Real, PW, 32>(args, hops, max_ctas, ctas, stream);
    case 64: return launch_bound_hops<Real, PW, 64>(args, hops, max_ctas, ctas, stream);
    default: return cudaErrorInvalidValue;
  }
}
template cudaError_t launch_grid_bound<LPMP_INST_REAL, LPMP_INST_PW>(const GridBoundArgs<LPMP_INST_REAL>&, uint32_t, int, int, int*, cudaStream_t);
#endif

template cudaError_t launch_family<LPMP_INST_REAL, LPMP_INST_PW, LPMP_INST_ANISO>(const CUtensorMap*, const GridArgs<LPMP_INST_REAL>&, uint32_t, int, int,
                                                                                  const WavefrontTuning&, int, cudaStream_t);

}  // namespace lpmp

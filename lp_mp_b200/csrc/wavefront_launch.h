// Host-side launcher of the wavefront pass kernel (kernels_grid.cuh).  The kernel is instantiated per
// (Real, pairwise family) in wavefront_inst_*.cu so that the translation units compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_grid.cuh"

namespace lpmp {

struct WavefrontTuning {
  int vpl = 4;     // labels per lane
  int warps = 4;   // bands (warps) per CTA
  int stages = 8;  // cp.async ring depth (steps of prefetch)
  int max_ctas = 0;  // 0 = as many as there are bands / warps
};

// true if a kernel instance exists for this shape
bool wavefront_supported(int pw, uint32_t L, int vpl, int stages);
// rows per band for this shape
int wavefront_rows_per_band(uint32_t L, int vpl);

template <typename Real>
cudaError_t launch_wavefront(const GridArgs<Real>& args, int pw, uint32_t L, const WavefrontTuning& tune, int num_sms,
                             cudaStream_t stream);

}  // namespace lpmp

// repro: TMA 4-D tile store with out-of-bounds parts
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int c2, int c3, int do_load, int do_store, int sc2, int sc3) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  double* buf = reinterpret_cast<double*>(sm + 128);
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), buf_a = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  if (do_load) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(8u * 8 * 40 * 8));
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];"
        ::"r"(buf_a), "l"(&tm), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < 8 * 8 * 40; i += 32) buf[i] += 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (do_store && threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1,%2,%3,%4}], [%5];" ::"l"(&tm), "r"(0), "r"(0), "r"(sc2), "r"(sc3), "r"(buf_a) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
#include <cstdlib>
int main(int argc, char** argv) {
  const int H = 12, W = 10, L = 8;
  double* d; cudaMalloc(&d, sizeof(double) * H * W * 5 * L); cudaMemset(d, 0, sizeof(double) * H * W * 5 * L);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap tm;
  const cuuint64_t dims[4] = {L, 5, W, H};
  const cuuint64_t strides[3] = {L * 8, L * 5 * 8, (cuuint64_t)W * L * 5 * 8};
  const cuuint32_t box[4] = {L, 5, 8, 8}; const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int cases[][6] = {{-8,0,1,0,0,0}, {0,-4,1,0,0,0}, {-3,-1,1,0,0,0}, {0,0,1,1,-8,0}, {0,0,1,1,0,-4}, {0,0,1,1,8,8}};
  int which = argc > 1 ? atoi(argv[1]) : 0;
  {
    auto& c = cases[which];
    k<<<1, 32, 100 * 1024>>>(tm, c[0], c[1], c[2], c[3], c[4], c[5]);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case load(%d,%d)=%d store(%d,%d)=%d -> %s\n", c[0], c[1], c[2], c[4], c[5], c[3], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  std::vector<double> h(H * W * 5 * L); cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
  double s = 0; for (double v : h) s += v; printf("sum %f\n", s);
  return 0;
}

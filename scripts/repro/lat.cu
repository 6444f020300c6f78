// Latencies that bound one node update of the wavefront chain (cycles, one warp alone on its scheduler):
//   shuffle, CREDUX.MIN.F32 (redux.sync.min.f32, sm_100a), LDS, and a shared-memory hand-off between two warps of a CTA
//   (producer: st.volatile.shared of a 16-byte vector; consumer: spin on ld.volatile.shared until it is not the sentinel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_shfl(float* o, int iters, long long* t) {
  float v = (float)threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1 + (i & 7))) + 1.0f;
  long long t1 = clock64();
  o[threadIdx.x] = v;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_shfl_only(float* o, int iters, long long* t) {
  float v = (float)threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1);
  long long t1 = clock64();
  o[threadIdx.x] = v;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_redux(float* o, int iters, long long* t) {
  float v = (float)threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    v = r + (float)threadIdx.x;
  }
  long long t1 = clock64();
  o[threadIdx.x] = v;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
// two independent reductions (lower / upper half-warp via +inf masking) per iteration
__global__ void k_redux2(float* o, int iters, long long* t) {
  float v = (float)threadIdx.x;
  const bool lo = threadIdx.x < 16;
  const float INF = __int_as_float(0x7f800000);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float ra, rb;
    const float va = lo ? v : INF, vb = lo ? INF : v;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(ra) : "f"(va));
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(rb) : "f"(vb));
    v = (lo ? ra : rb) + (float)threadIdx.x;
  }
  long long t1 = clock64();
  o[threadIdx.x] = v;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_lds(float* o, int iters, long long* t) {
  __shared__ unsigned idx[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i * 33 + 7) & 1023;
  __syncwarp();
  unsigned p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) p = *(volatile unsigned*)&idx[p];
  long long t1 = clock64();
  o[threadIdx.x] = (float)p;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
// ping-pong of a 16-byte self-validating slot between warp 0 and warp `peer` of one CTA (all 32 lanes poll, like a node group)
__global__ void k_pingpong(float* o, int iters, int peer, long long* t) {
  __shared__ __align__(16) float slot_a[32 * 4], slot_b[32 * 4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w == 0) for (int k = 0; k < 4; ++k) { slot_a[lane * 4 + k] = __int_as_float(-1); slot_b[lane * 4 + k] = __int_as_float(-1); }
  __syncthreads();
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot_a + lane * 4), b = (unsigned)__cvta_generic_to_shared(slot_b + lane * 4);
  float x = (float)lane;
  long long t0 = clock64();
  if (w == 0) {
    for (int i = 0; i < iters; ++i) {
      asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(a), "f"(x) : "memory");
      float r0, r1, r2, r3;
      bool miss;
      do {
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3) : "r"(b) : "memory");
        miss = __float_as_int(r0) == -1;
      } while (__any_sync(0xffffffffu, miss));
      asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(b), "r"(-1) : "memory");
      x = r0 + 1.0f;
    }
  } else if (w == peer) {
    for (int i = 0; i < iters; ++i) {
      float r0, r1, r2, r3;
      bool miss;
      do {
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3) : "r"(a) : "memory");
        miss = __float_as_int(r0) == -1;
      } while (__any_sync(0xffffffffu, miss));
      asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(-1) : "memory");
      x = r0 + 1.0f;
      asm volatile("st.volatile.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(b), "f"(x) : "memory");
    }
  }
  long long t1 = clock64();
  if (w == 0 || w == peer) o[threadIdx.x] = x;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
// same ping-pong with a scalar (one label per lane) slot
__global__ void k_pingpong1(float* o, int iters, int peer, long long* t) {
  __shared__ float slot_a[32], slot_b[32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w == 0) { slot_a[lane] = __int_as_float(-1); slot_b[lane] = __int_as_float(-1); }
  __syncthreads();
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot_a + lane), b = (unsigned)__cvta_generic_to_shared(slot_b + lane);
  float x = (float)lane;
  long long t0 = clock64();
  if (w == 0) {
    for (int i = 0; i < iters; ++i) {
      asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
      float r0;
      bool miss;
      do {
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(r0) : "r"(b) : "memory");
        miss = __float_as_int(r0) == -1;
      } while (__any_sync(0xffffffffu, miss));
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(b), "r"(-1) : "memory");
      x = r0 + 1.0f;
    }
  } else if (w == peer) {
    for (int i = 0; i < iters; ++i) {
      float r0;
      bool miss;
      do {
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(r0) : "r"(a) : "memory");
        miss = __float_as_int(r0) == -1;
      } while (__any_sync(0xffffffffu, miss));
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(-1) : "memory");
      x = r0 + 1.0f;
      asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(b), "f"(x) : "memory");
    }
  }
  long long t1 = clock64();
  if (w == 0 || w == peer) o[threadIdx.x] = x;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}
int main() {
  float* o; long long* t; long long h;
  cudaMalloc(&o, 4096); cudaMalloc(&t, 64);
  const int iters = 20000;
  auto rep = [&](const char* name, double per) { cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost); printf("%-44s %.1f cycles\n", name, (double)h / iters / per); };
  for (int r = 0; r < 2; ++r) {
    k_shfl<<<1, 32>>>(o, iters, t); rep("shfl.bfly + fmnmx + fadd chain", 1);
    k_shfl_only<<<1, 32>>>(o, iters, t); rep("shfl.bfly chain", 1);
    k_redux<<<1, 32>>>(o, iters, t); rep("credux.min.f32 + fadd chain", 1);
    k_redux2<<<1, 32>>>(o, iters, t); rep("2 x credux (half-warp masks) + sel + fadd", 1);
    k_lds<<<1, 32>>>(o, iters, t); rep("lds dependent chain", 1);
    for (int peer : {1, 4, 2}) {
      k_pingpong<<<1, 32 * 8>>>(o, iters, peer, t); printf("peer warp %d: ", peer); rep("smem v4 hand-off one way (incl. re-arm)", 2);
      k_pingpong1<<<1, 32 * 8>>>(o, iters, peer, t); printf("peer warp %d: ", peer); rep("smem scalar hand-off one way (incl. re-arm)", 2);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

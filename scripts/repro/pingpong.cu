// Hand-off latency between two CTAs through L2 (what one band-to-band hand-off of the wavefront kernel costs).
//   ping-pong: CTA 0 stores i, CTA 1 polls for i and answers, CTA 0 polls for the answer: time / (2*iters) = one-way latency
//   poll RT  : one thread re-loads the same L2-resident word with a dependent chain: time / iters = one poll round trip
#include <cstdio>
#include <cuda_runtime.h>
template <int SCOPE> __device__ __forceinline__ unsigned ld(const unsigned* p) {
  unsigned v;
  if (SCOPE == 0) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (SCOPE == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <int SCOPE> __device__ __forceinline__ void st(unsigned* p, unsigned v) {
  if (SCOPE == 0) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else if (SCOPE == 1) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  else asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int SCOPE> __global__ void pingpong(unsigned* a, unsigned* b, int iters, long long* out) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  if (blockIdx.x == 0) {
    for (int i = 1; i <= iters; ++i) { st<SCOPE>(a, i); while (ld<SCOPE>(b) != (unsigned)i) {} }
    out[0] = clock64() - t0;
  } else if (blockIdx.x == gridDim.x - 1) {
    for (int i = 1; i <= iters; ++i) { while (ld<SCOPE>(a) != (unsigned)i) {} st<SCOPE>(b, i); }
  }
}
template <int SCOPE> __global__ void pollrt(unsigned* a, int iters, long long* out) {
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) acc += ld<SCOPE>(a + (acc & 1u));
  out[0] = clock64() - t0; out[1] = acc;
}
int main() {
  unsigned* buf; long long* out; long long h[2];
  cudaMalloc(&buf, 4096); cudaMalloc(&out, 16); cudaMemset(buf, 0, 4096);
  const int iters = 20000;
  const char* names[3] = {"relaxed.sys", "relaxed.gpu", "volatile"};
  for (int grid : {2, 148}) for (int s = 0; s < 3; ++s) {
    cudaMemset(buf, 0, 4096);
    if (s == 0) pingpong<0><<<grid, 32>>>(buf, buf + 64, iters, out);
    if (s == 1) pingpong<1><<<grid, 32>>>(buf, buf + 64, iters, out);
    if (s == 2) pingpong<2><<<grid, 32>>>(buf, buf + 64, iters, out);
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("pingpong %-12s CTA0<->CTA%d one-way %.0f cycles\n", names[s], grid - 1, (double)h[0] / (2.0 * iters));
  }
  for (int s = 0; s < 3; ++s) {
    cudaMemset(buf, 0, 4096);
    if (s == 0) pollrt<0><<<1, 1>>>(buf, iters, out);
    if (s == 1) pollrt<1><<<1, 1>>>(buf, iters, out);
    if (s == 2) pollrt<2><<<1, 1>>>(buf, iters, out);
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("poll RT  %-12s %.0f cycles\n", names[s], (double)h[0] / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

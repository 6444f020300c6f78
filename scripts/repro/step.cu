// One compute warp of the split wavefront kernel in isolation: cycles per steady step (truncated linear, L = 16, WIN = 3, f32)
// for several orderings of the same instruction mix.  Answers "where do the 370 cycles of a step go" without the rest of the
// kernel (tile warp, in-mover, neighbours).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o step step.cu
//   V0: the step as kernels_grid_split.cuh issues it (poll + loads at the top, per-lane sentinel branch)
//   V1: poll word and parameters of the NEXT step loaded at the top of this step, per-lane sentinel branch
//   V2: V1, sentinel test as a vote at the end of the previous step (uniform branch, no BSSY / BRA.DIV before the shuffles)
//   V3: V2 without the second CREDUX          (diagnostic, wrong result)
//   V4: V2 without the xor-16 shuffle         (diagnostic, wrong result)
//   V5: V2 without the record stores          (diagnostic)
//   V6: V2 with shuffle minima instead of CREDUX (diagnostic)
//   V7: V2 with the whole tile's record fields preloaded into registers
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int L = 16, K = 8, REC = 5 * L, PREC = 4 * 2, NSLOT = 8, RING = 32, WIN = 3;
constexpr uint32_t S = 4, REC_B = REC * S, PREC_B = PREC * S, SLOT_REC_B = K * REC_B, SLOT_PAR_B = K * PREC_B, LB = L * S;

__device__ __forceinline__ float credux_min(float v) { float r; asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ bool is_sent(float v) { return __float_as_int(v) == -1; }
__device__ __forceinline__ float lds(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float ldsv(uint32_t a) { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void lds2(float& a0, float& a1, uint32_t a) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a0), "=f"(a1) : "r"(a)); }
__device__ __forceinline__ void sts_if(bool on, uint32_t a, float v) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.volatile.shared.f32 [%0], %1; }" ::"r"(a), "f"(v), "r"((uint32_t)on) : "memory"); }

template <int V>
__device__ __forceinline__ float half_min(float v, bool lo) {
  const float INF = __int_as_float(0x7f800000);
  if (V == 6) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  const float a0 = credux_min(lo ? v : INF);
  if (V == 3) return a0;
  const float a1 = credux_min(lo ? INF : v);
  return lo ? a0 : a1;
}

template <int V>
__device__ __forceinline__ float min_plus(float slope, float trunc, float o, bool lo) {
  float up[WIN], dn[WIN];
#pragma unroll
  for (int d = 1; d <= WIN; ++d) { up[d - 1] = __shfl_up_sync(0xffffffffu, o, d, 16); dn[d - 1] = __shfl_down_sync(0xffffffffu, o, d, 16); }
  float r = fminf(o, trunc + half_min<V>(o, lo));
#pragma unroll
  for (int d = 1; d <= WIN; ++d) { const float c = fminf(slope * (float)d, trunc); r = fminf(r, fminf(c + up[d - 1], c + dn[d - 1])); }
  return r;
}

template <int V>
__global__ void __launch_bounds__(256, 1) k_step(float* out, int ntiles, long long* t, int self_first_i, int out_col_i, int fence) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t WB = 2 * RING * LB + NSLOT * (SLOT_REC_B + SLOT_PAR_B);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + (threadIdx.x >> 5) * WB;
  const uint32_t ring_in0 = base, ring_out0 = base + RING * LB, tiles = base + 2 * RING * LB, pars = tiles + NSLOT * SLOT_REC_B;
  const int lane = threadIdx.x & 31;
  float* f = reinterpret_cast<float*>(smem + (threadIdx.x >> 5) * WB);
  for (int i = lane; i < (int)((2 * RING * LB + NSLOT * (SLOT_REC_B + SLOT_PAR_B)) / 4); i += 32) f[i] = 0.001f * (float)((i * 37) % 101);
  for (int i = lane; i < NSLOT * K * PREC; i += 32) f[(pars - base) / 4 + i] = (i & 1) ? 1.0f : 0.25f;
  __syncthreads();
  const bool hi = lane >= 16, lo = !hi;
  const int xl = lane & 15;
  const bool self_first = (self_first_i != 0) != hi, out_col_exists = out_col_i != 0, out_ring = hi && out_col_exists;
  const uint32_t ring_in = ring_in0 + xl * S, ring_out = ring_out0 + xl * S, dummy = ring_out0 + (RING - 1) * LB + xl * S;
  const uint32_t f_th = tiles + xl * S, f_in = f_th + (hi ? 4 : 3) * LB, f_out = f_th + (hi ? 2 : 1) * LB, f_par = pars + (hi ? 2 : 0) * S;
  float oth = 0.f, pvn = ldsv(ring_in), parn0, parn1;
  lds2(parn0, parn1, f_par);
  long long t0 = clock64();
#pragma unroll 1
  for (int j = 0; j < ntiles; ++j) {
    const uint32_t slot = (uint32_t)j & (NSLOT - 1);
    const uint32_t Tth = f_th + slot * SLOT_REC_B, Tin = f_in + slot * SLOT_REC_B, Tout = f_out + slot * SLOT_REC_B, Tpar = f_par + slot * SLOT_PAR_B;
    const uint32_t c0 = ((uint32_t)(j * K) & (RING - 1)) * LB;
    const uint32_t Rin = ring_in + c0, Rout = ring_out + c0;
    const uint32_t slotn = (uint32_t)(j + 1) & (NSLOT - 1);
    const uint32_t Tparn = f_par + slotn * SLOT_PAR_B, Rinn = ring_in + (((uint32_t)((j + 1) * K) & (RING - 1)) * LB);
    float Pth[K], Poi[K], Poo[K];
    if (V == 7) {
#pragma unroll
      for (int jj = 0; jj < K; ++jj) { Pth[jj] = lds(Tth + jj * REC_B); Poi[jj] = lds(Tin + jj * REC_B); Poo[jj] = lds(Tout + jj * REC_B); }
    }
#pragma unroll
    for (int jj = 0; jj < K; ++jj) {
      const uint32_t ro = jj * REC_B, po = jj * PREC_B, rin = Rin + jj * LB;
      float pv, p0, p1, th, oi, oo;
      if (V == 0) {
        pv = ldsv(rin);
        lds2(p0, p1, Tpar + po);
        th = lds(Tth + ro); oi = lds(Tin + ro); oo = lds(Tout + ro);
        if (is_sent(pv)) { do { pv = ldsv(rin); } while (is_sent(pv)); }
      } else {
        pv = pvn; p0 = parn0; p1 = parn1;
        if (V == 1) { if (is_sent(pv)) { do { pv = ldsv(rin); } while (is_sent(pv)); } }
        // next step's poll word and parameters
        pvn = ldsv(jj + 1 < K ? rin + LB : Rinn);
        lds2(parn0, parn1, jj + 1 < K ? Tpar + po + PREC_B : Tparn);
        if (V == 7) { th = Pth[jj]; oi = Poi[jj]; oo = Poo[jj]; }
        else { th = lds(Tth + ro); oi = lds(Tin + ro); oo = lds(Tout + ro); }
      }
      sts_if(hi, dummy, __int_as_float(-1));
      oth = hi ? pv : oth;
      const float h = min_plus<V>(p0, p1, oth, lo);
      const float m = h + oi;
      oi = oi - m;
      const float mo = V == 4 ? m : __shfl_xor_sync(0xffffffffu, m, 16);
      const float v = (th + (self_first ? m : mo)) + (self_first ? mo : m);
      const float d = 0.5f * v;
      const float d_oc = out_col_exists ? d : 0.f;
      oo = oo + (hi ? d_oc : d);
      th = (v - d_oc) - d;
      sts_if(out_ring, Rout + jj * LB, oo);
      if (V != 5) {
        sts_if(!hi, Tth + ro, th);
        sts_if(true, Tin + ro, oi);
        sts_if(true, Tout + ro, oo);
      }
      oth = oo;
      if (V >= 2) {
        if (__any_sync(0xffffffffu, is_sent(pvn))) {
          const uint32_t a = jj + 1 < K ? rin + LB : Rinn;
          do { pvn = ldsv(a); } while (__any_sync(0xffffffffu, is_sent(pvn)));
        }
      }
    }
    if (fence) {   // publication of a finished tile, as the kernel does it
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ring_out0 + (RING - 2) * LB), "r"(j) : "memory");
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = oth;
  if (threadIdx.x == 0) t[0] = t1 - t0;
}

template <int V>
void run(const char* name, float* o, long long* t, int nwarps = 1, int fence = 0) {
  const int ntiles = 512;
  const size_t smem = (size_t)nwarps * (2 * RING * LB + NSLOT * (SLOT_REC_B + SLOT_PAR_B));
  cudaFuncSetAttribute(k_step<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long best = 1LL << 60;
  for (int rep = 0; rep < 3; ++rep) {
    k_step<V><<<1, 32 * nwarps, smem>>>(o, ntiles, t, 1, 1, fence);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, t, sizeof(c), cudaMemcpyDeviceToHost);
    if (c < best) best = c;
  }
  float h[32];
  cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-60s warps %d fence %d %7.1f cycles / step   (check %.6f %.6f)  %s\n", name, nwarps, fence, (double)best / (ntiles * K), h[0], h[17], cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* o; long long* t;
  cudaMalloc(&o, 4096); cudaMalloc(&t, 64);
  run<0>("V0 step as issued today", o, t);
  run<1>("V1 poll + parameters one step ahead, per-lane branch", o, t);
  run<2>("V2 V1 + sentinel vote at the end of the previous step", o, t);
  run<3>("V3 V2 - second CREDUX (diagnostic)", o, t);
  run<4>("V4 V2 - xor-16 shuffle (diagnostic)", o, t);
  run<5>("V5 V2 - record stores (diagnostic)", o, t);
  run<6>("V6 V2 with shuffle minima (diagnostic)", o, t);
  run<7>("V7 V2 + tile fields preloaded into registers", o, t);
  run<0>("V0", o, t, 8, 0);
  run<2>("V2", o, t, 8, 0);
  run<0>("V0", o, t, 1, 1);
  run<2>("V2", o, t, 1, 1);
  run<0>("V0", o, t, 8, 1);
  run<2>("V2", o, t, 8, 1);
  run<0>("V0", o, t, 4, 1);
  run<2>("V2", o, t, 4, 1);
  return 0;
}

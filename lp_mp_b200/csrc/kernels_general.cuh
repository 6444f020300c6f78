// General-graph kernels: level-scheduled node update (covers the graph-colouring schedule), lower bound,
// primal evaluation, single-factor entry points.
//
// Replaces, for the SRMP simplex configuration:
//   LP::ComputePass(it...) + FactorContainer::UpdateFactor / UpdateFactorPrimal
//        (/root/reference/include/LP_MP.h:981-1005,1592-1602; include/factors_messages.hxx:2256-2261,2332-2373)
//   LP::LowerBound, LP::EvaluatePrimal (include/LP_MP.h:1507-1536)
// One sub-warp group of G lanes (G = pow2 >= min(L,32)) owns one node (or edge); lanes stride over labels.
#pragma once
#include "device_common.cuh"

namespace lpmp {

// topology of the edge slots: explicit endpoint arrays (general graphs) or index math (grids)
struct GraphTopo {
  const uint32_t* ei;
  const uint32_t* ej;
  uint64_t num_slots;
  __device__ __forceinline__ uint32_t second_label(const uint32_t* labels, uint64_t e) const { return labels[ej[e]]; }
  __device__ __forceinline__ bool valid(uint64_t) const { return true; }
  __device__ __forceinline__ uint32_t first(uint64_t e) const { return ei[e]; }
  __device__ __forceinline__ uint32_t second(uint64_t e) const { return ej[e]; }
  template <typename St> __device__ __forceinline__ auto b_of(const St& st, uint64_t e) const { return st.b(e); }
};
// grid slots: [0, H*W) horizontal edge (r,c)-(r,c+1) at slot r*W+c; [H*W, 2*H*W) vertical edge (r,c)-(r+1,c).
// A stripe's last row owns the edges to the stripe below; their b side is passed in `below_b` ([W][L]).
template <typename Real>
struct GridTopo {
  uint32_t H, W;
  uint64_t num_slots;
  const Real* below_b;
  const uint32_t* below_labels = nullptr;   // [W] labels of the first row of the stripe below (primal evaluation of a stripe)
  __device__ __forceinline__ uint32_t second_label(const uint32_t* labels, uint64_t e) const {
    const uint64_t HW = (uint64_t)H * W;
    if (e >= HW && (e - HW) / W + 1 >= H) return below_labels[(e - HW) % W];
    return labels[second(e)];
  }
  __device__ __forceinline__ bool valid(uint64_t e) const {
    const uint64_t HW = (uint64_t)H * W;
    if (e < HW) return (e % W) + 1 < W;
    return (e - HW) / W + 1 < H || below_b != nullptr;
  }
  __device__ __forceinline__ uint32_t first(uint64_t e) const { const uint64_t HW = (uint64_t)H * W; return (uint32_t)(e < HW ? e : e - HW); }
  __device__ __forceinline__ uint32_t second(uint64_t e) const { const uint64_t HW = (uint64_t)H * W; return (uint32_t)(e < HW ? e + 1 : e - HW + W); }
  template <typename St> __device__ __forceinline__ const Real* b_of(const St& st, uint64_t e) const {
    const uint64_t HW = (uint64_t)H * W;
    if (e >= HW && (e - HW) / W + 1 >= H) return below_b + ((e - HW) % W) * (uint64_t)st.L_;
    return st.b(e);
  }
};

// Where the vectors of node i / edge slot e live.  General graphs: theta[N][L], a[E][L], b[E][L], params[E][P].
// Grids: the node records of kernels_grid.cuh (rec[N][5][L], par[N][4][P]); slots [0, N) are the horizontal edges
// (first = node e), slots [N, 2N) the vertical ones (first = node e - N).
template <typename Real>
struct Store {
  Real* theta; uint64_t theta_stride;
  Real* a0; Real* a1; Real* b0; Real* b1;
  const Real* p0; const Real* p1;
  uint64_t split, estride, pstride;
  uint32_t L_;
  __device__ __forceinline__ Real* th(uint64_t i) const { return theta + i * theta_stride; }
  __device__ __forceinline__ Real* a(uint64_t e) const { return e < split ? a0 + e * estride : a1 + (e - split) * estride; }
  __device__ __forceinline__ Real* b(uint64_t e) const { return e < split ? b0 + e * estride : b1 + (e - split) * estride; }
  __device__ __forceinline__ const Real* par(uint64_t e) const { return e < split ? p0 + e * pstride : p1 + (e - split) * pstride; }
};

template <typename Real>
struct LevelArgs {
  Store<Real> st;
  const uint32_t* nodes;    // nodes of this level
  uint32_t num_nodes;
  const uint64_t* inc_off;  // [N+1]
  const uint32_t* inc_edge; // incidence entries in the reference's per-node message order
  const uint8_t* inc_flags; // bit0: side (0 = node is first variable), bit1: other endpoint is later (forward sense)
  uint32_t* labels;         // [N]
  uint32_t L;
  int pw, repam, forward, primal;
  int residual;             // update_factor_residual instead of UpdateFactor in the plain passes
};

// One node update = receive, (argmin), snapshot, send.  Shared memory per group: th[L], t[L], oth[L], hs[L].
template <typename Real>
__global__ void __launch_bounds__(256) level_update_kernel(LevelArgs<Real> A, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real* smem = reinterpret_cast<Real*>(smem_raw);
  const int groups_per_cta = blockDim.x / G;
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned mask = group_mask(G);
  const uint32_t L = A.L;
  Real* th = smem + (size_t)g * 4 * L;
  Real* t = th + L;
  Real* oth = t + L;
  Real* hs = oth + L;
  const bool aniso = A.repam == LPMP_REPAM_ANISOTROPIC || A.repam == LPMP_REPAM_ANISOTROPIC2;

  for (uint32_t k = blockIdx.x * groups_per_cta + g; k < A.num_nodes; k += gridDim.x * groups_per_cta) {
    const uint32_t i = A.nodes[k];
    const uint64_t k0 = A.inc_off[i], k1 = A.inc_off[i + 1];
    Real* theta_i = A.st.th(i);
    for (uint32_t x = gl; x < L; x += G) th[x] = theta_i[x];
    uint32_t n_in = 0, n_out = 0;
    // ReceiveMessages (include/factors_messages.hxx:2440-2471)
    for (uint64_t q = k0; q < k1; ++q) {
      const uint32_t e = A.inc_edge[q];
      const uint32_t fl = A.inc_flags[q];
      const uint32_t side = fl & 1u, later = (fl >> 1) & 1u;
      if (later) ++n_out; else ++n_in;
      const bool earlier = A.forward ? !later : later;
      if (aniso && !earlier) continue;
      Real* own = side == 0 ? A.st.a(e) : A.st.b(e);
      const Real* other = side == 0 ? A.st.b(e) : A.st.a(e);
      __syncwarp(mask);
      for (uint32_t x = gl; x < L; x += G) oth[x] = other[x];
      __syncwarp(mask);
      group_min_plus<Real>(A.pw, A.st.par(e), L, side, oth, hs, G, gl, mask);
      for (uint32_t x = gl; x < L; x += G) {
        const Real ow = own[x];
        const Real m = hs[x] + ow;
        th[x] = th[x] + m;   // RepamLeft(+)   :498-500
        own[x] = ow - m;     // RepamRight(-)
      }
    }
    if (A.primal) {  // MaximizePotentialAndComputePrimal: first minimum
      Real v = RealTraits<Real>::inf();
      uint32_t idx = 0xffffffffu;
      for (uint32_t x = gl; x < L; x += G) if (th[x] < v) { v = th[x]; idx = x; }
      group_argmin(v, idx, G, mask);
      if (gl == 0) A.labels[i] = idx;
    }
    if (k1 > k0) {
      const uint32_t deg = n_in + n_out;
      double wd;
      if (A.repam == LPMP_REPAM_ANISOTROPIC) wd = 1.0 / (double)(n_in > n_out ? n_in : n_out);   // include/LP_MP.h:1397
      else if (A.repam == LPMP_REPAM_ANISOTROPIC2) wd = 1.0 / (double)(A.forward ? n_out : n_in);   // :1134
      else if (A.repam == LPMP_REPAM_UNIFORM) wd = 1.0 / (double)deg;         // :1437, leave_weight 0
      else wd = 1.0 / (double)(deg + 1);                                      // :1437, leave_weight 1
      const Real w = (Real)wd;
      for (uint32_t x = gl; x < L; x += G) t[x] = th[x];  // tmp_factor snapshot, :2801-2805
      // CallSendMessages (:2698-2744)
      for (uint64_t q = k0; q < k1; ++q) {
        const uint32_t e = A.inc_edge[q];
        const uint32_t fl = A.inc_flags[q];
        const uint32_t side = fl & 1u, later = (fl >> 1) & 1u;
        const bool is_later = A.forward ? later : !later;
        if (aniso && !is_later) continue;
        Real* own = side == 0 ? A.st.a(e) : A.st.b(e);
        for (uint32_t x = gl; x < L; x += G) {
          const Real d = w * t[x];
          own[x] = own[x] + d;  // RepamRight(+) first, :501-503
          th[x] = th[x] - d;    // RepamLeft(-)
        }
      }
      if (A.residual && !A.primal) {   // send_messages_residual (include/factors_messages.hxx:2957-3007)
        Real r = Real(0);
        for (uint64_t q = k0; q < k1; ++q) {
          const uint32_t e = A.inc_edge[q];
          const uint32_t fl = A.inc_flags[q];
          const uint32_t side = fl & 1u, later = (fl >> 1) & 1u;
          const bool is_later = A.forward ? later : !later;
          if (aniso && !is_later) continue;
          r = r + w;
          Real* own = side == 0 ? A.st.a(e) : A.st.b(e);
          for (uint32_t x = gl; x < L; x += G) {
            const Real d = r * th[x];
            own[x] = own[x] + d;
            th[x] = th[x] - d;
          }
        }
      }
    }
    for (uint32_t x = gl; x < L; x += G) theta_i[x] = th[x];
    __syncwarp(mask);
  }
}

// ---- lower bound: partial sums per CTA, deterministic second stage -------------------------------------------
template <typename Real, typename Topo>
struct BoundArgs {
  Store<Real> st;
  Topo topo;
  uint64_t num_nodes;
  uint32_t L;
  int pw;
  double* partial;  // [gridDim.x]
};

// LP::LowerBound (include/LP_MP.h:1507-1518): sum_i min theta_i + sum_e min_x [a_e(x) + min_y (c_e(x,y) + b_e(y))]
template <typename Real, typename Topo>
__global__ void __launch_bounds__(256) lower_bound_kernel(BoundArgs<Real, Topo> A, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real* smem = reinterpret_cast<Real*>(smem_raw);
  const int groups_per_cta = blockDim.x / G;
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned mask = group_mask(G);
  const uint32_t L = A.L;
  Real* oth = smem + (size_t)g * 2 * L;
  Real* hs = oth + L;
  const Real INF = RealTraits<Real>::inf();
  double acc = 0.0;
  const uint64_t total = A.num_nodes + A.topo.num_slots;
  for (uint64_t k = (uint64_t)blockIdx.x * groups_per_cta + g; k < total; k += (uint64_t)gridDim.x * groups_per_cta) {
    Real m = INF;
    if (k < A.num_nodes) {
      const Real* th = A.st.th(k);
      for (uint32_t x = gl; x < L; x += G) m = rmin(m, th[x]);
    } else {
      const uint64_t e = k - A.num_nodes;
      if (!A.topo.valid(e)) continue;
      __syncwarp(mask);
      const Real* eb = A.topo.b_of(A.st, e);
      const Real* ea = A.st.a(e);
      for (uint32_t x = gl; x < L; x += G) oth[x] = eb[x];
      __syncwarp(mask);
      group_min_plus<Real>(A.pw, A.st.par(e), L, 0u, oth, hs, G, gl, mask);
      for (uint32_t x = gl; x < L; x += G) m = rmin(m, Real(hs[x] + ea[x]));
    }
    m = group_min(m, G, mask);
    if (gl == 0) acc += (double)m;
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

// LP::EvaluatePrimal (include/LP_MP.h:1520-1536): sum_i theta_i(x_i) + sum_e [(c_e(x_i,x_j) + a_e(x_i)) + b_e(x_j)]
template <typename Real, typename Topo>
__global__ void __launch_bounds__(256) primal_cost_kernel(BoundArgs<Real, Topo> A, const uint32_t* __restrict__ labels) {
  const uint32_t L = A.L;
  double acc = 0.0;
  const uint64_t total = A.num_nodes + A.topo.num_slots;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
    if (k < A.num_nodes) {
      acc += (double)A.st.th(k)[labels[k]];
    } else {
      const uint64_t e = k - A.num_nodes;
      if (!A.topo.valid(e)) continue;
      const uint32_t x = labels[A.topo.first(e)], y = A.topo.second_label(labels, e);
      const Real c = pairwise_cost<Real>(A.pw, A.st.par(e), L, x, y);
      acc += (double)Real(Real(c + A.st.a(e)[x]) + A.topo.b_of(A.st, e)[y]);
    }
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

// Stripe boundary push (per-pass halo coupling through peer memory): this stripe's side of the boundary edges, src[c * src_stride
// + x], into the neighbouring stripe's in-buffer dst[c * L + x] -- self-validating slots: a slot is written only once it holds
// the sentinel again (the consumer of the previous pass has taken its vector), so consecutive pushes can never overtake the
// consumer.  One thread per real; bounded wait (sticky abort flag in control[1]).
template <typename Real>
__global__ void push_boundary_kernel(Real* dst, const Real* __restrict__ src, uint32_t W, uint32_t L, uint64_t src_stride, unsigned int* control, bool reverse) {
  const uint64_t total = (uint64_t)W * L;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
    // slot k = (q, x) of the in-buffer: q counts columns in the PROCESSING order of the consuming pass (backward: q = W - 1 - c)
    const uint64_t c = reverse ? (uint64_t)W - 1 - k / L : k / L;
    const Real v = src[c * src_stride + (k % L)];
    unsigned spins = 0;
    for (;;) {
      Real cur;
      if constexpr (sizeof(Real) == 4) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(cur) : "l"(dst + k) : "memory");
      else asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(cur) : "l"(dst + k) : "memory");
      bool free_slot;
      if constexpr (sizeof(Real) == 4) free_slot = __float_as_uint(cur) == 0xffffffffu;
      else free_slot = __double_as_longlong(cur) == -1ll;
      if (free_slot) break;
      if ((++spins & 0xffu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 4000000000ull) { if (control) atomicExch(control + 1, 1u); break; }
      }
    }
    if constexpr (sizeof(Real) == 4) asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst + k), "f"(v) : "memory");
    else asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + k), "d"(v) : "memory");
  }
}

// Uniform-mode passes of a stripe on the split kernel: the vectors the neighbouring stripe pushed into in-buffer `in` ([W][L], slot
// protocol) are parked in the stripe's ghost row -- dst[c * dst_stride + x], the in-side slot of the records of row H (forward) or
// row -1 (backward) -- where the pass kernel's tile warp finds them like any other row below a band.  Waits for every slot
// (the push is issued before the neighbour's pass, so this never waits on a pass), re-arms it; bounded wait.
template <typename Real>
__global__ void later_gather_kernel(Real* __restrict__ dst, Real* in, uint32_t W, uint32_t L, uint64_t dst_stride, unsigned int* control, bool reverse) {
  const uint64_t total = (uint64_t)W * L;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (uint64_t)gridDim.x * blockDim.x) {
    Real cur;
    unsigned spins = 0;
    for (;;) {
      if constexpr (sizeof(Real) == 4) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(cur) : "l"(in + k) : "memory");
      else asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(cur) : "l"(in + k) : "memory");
      bool empty;
      if constexpr (sizeof(Real) == 4) empty = __float_as_uint(cur) == 0xffffffffu;
      else empty = __double_as_longlong(cur) == -1ll;
      if (!empty) break;
      if ((++spins & 0xffu) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 4000000000ull) { if (control) atomicExch(control + 1, 1u); cur = Real(0); break; }
      }
    }
    dst[(reverse ? (uint64_t)W - 1 - k / L : k / L) * dst_stride + (k % L)] = cur;   // in-buffers are in processing order, records in column order
    if constexpr (sizeof(Real) == 4) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(in + k), "r"(0xffffffffu) : "memory");
    else asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(in + k), "l"(-1ll) : "memory");
  }
}

// fixed-order final sum (deterministic bound): one warp, sequential chunks then a fixed tree; `extra` (m values, or null) is a
// second array of partial sums (the stripe-boundary edges next to the per-row terms of a fused bound)
__global__ void final_sum_kernel(const double* __restrict__ partial, int n, double constant, double* out, const double* __restrict__ extra = nullptr, int m = 0) {
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += 32) v += partial[k];
  for (int k = threadIdx.x; k < m; k += 32) v += extra[k];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (threadIdx.x == 0) *out = constant + v;
}

// LP::LowerBound terms of the `count` edges slot0 .. slot0 + count - 1 whose b side lives elsewhere (`ext_b`: [count][L]): the edges
// from a stripe's last row to the stripe below.  Same arithmetic as the edge branch of lower_bound_kernel.
template <typename Real>
__global__ void __launch_bounds__(256) boundary_bound_kernel(Store<Real> st, uint64_t slot0, uint32_t count, uint32_t L, int pw, const Real* __restrict__ ext_b,
                                                             double* __restrict__ partial, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real* smem = reinterpret_cast<Real*>(smem_raw);
  const int groups_per_cta = blockDim.x / G;
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned mask = group_mask(G);
  Real* oth = smem + (size_t)g * 2 * L;
  Real* hs = oth + L;
  double acc = 0.0;
  for (uint32_t k = blockIdx.x * groups_per_cta + g; k < count; k += gridDim.x * groups_per_cta) {
    const uint64_t e = slot0 + k;
    __syncwarp(mask);
    const Real* eb = ext_b + (uint64_t)k * L;
    const Real* ea = st.a(e);
    for (uint32_t x = gl; x < L; x += G) oth[x] = eb[x];
    __syncwarp(mask);
    group_min_plus<Real>(pw, st.par(e), L, 0u, oth, hs, G, gl, mask);
    Real m = RealTraits<Real>::inf();
    for (uint32_t x = gl; x < L; x += G) m = rmin(m, Real(hs[x] + ea[x]));
    m = group_min(m, G, mask);
    if (gl == 0) acc += (double)m;
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// ---- single-factor entry points --------------------------------------------------------------------------------
// n independent pairwise factors L1 x L2 (dense) or L x L (Potts / truncated linear); one thread per output entry.
template <typename Real>
__global__ void minmarginal_kernel(int pw, uint64_t n, uint32_t L1, uint32_t L2, const Real* __restrict__ p0,
                                   const Real* __restrict__ p1, const Real* __restrict__ msg1,
                                   const Real* __restrict__ msg2, int side, Real* __restrict__ out) {
  const uint32_t Lo = side == 0 ? L1 : L2, Li = side == 0 ? L2 : L1;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * Lo) return;
  const uint64_t k = idx / Lo;
  const uint32_t x = (uint32_t)(idx % Lo);
  const Real INF = RealTraits<Real>::inf();
  Real m = INF;
  for (uint32_t y = 0; y < Li; ++y) {
    const uint32_t x1 = side == 0 ? x : y, x2 = side == 0 ? y : x;
    Real c;
    if (pw == LPMP_PAIRWISE_DENSE) c = p0[(k * L1 + x1) * L2 + x2];
    else if (pw == LPMP_PAIRWISE_POTTS) c = x1 == x2 ? Real(0) : p0[k];
    else { const uint32_t d = x1 > x2 ? x1 - x2 : x2 - x1; c = rmin(Real(p0[k] * Real(d)), p1[k]); }
    const Real o = side == 0 ? (msg2 ? msg2[k * L2 + y] : Real(0)) : (msg1 ? msg1[k * L1 + y] : Real(0));
    m = rmin(m, Real(c + o));
  }
  const Real own = side == 0 ? (msg1 ? msg1[k * L1 + x] : Real(0)) : (msg2 ? msg2[k * L2 + x] : Real(0));
  out[idx] = m + own;
}

template <typename Real>
__global__ void row_min_kernel(uint64_t n, uint32_t L, const Real* __restrict__ v, Real* __restrict__ mn, uint32_t* __restrict__ arg) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t best = 0;
  Real bv = v[k * L];
  for (uint32_t x = 1; x < L; ++x) { const Real c = v[k * L + x]; if (c < bv) { bv = c; best = x; } }
  if (mn) mn[k] = bv;
  if (arg) arg[k] = best;
}

}  // namespace lpmp

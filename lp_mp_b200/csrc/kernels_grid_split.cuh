// Wavefront pass, "split" variant for the anisotropic (SRMP) passes: ONE ROW PER WARP, the two receive edges of a node split
// over the two half-warps.
//
// Replaces LP::ComputePass over forward/backwardUpdateOrdering_ (/root/reference/include/LP_MP.h:981-1005,1592-1602) and
// FactorContainer::UpdateFactor (include/factors_messages.hxx:2256-2261); same storage, TMA tiles and self-validating hand-off
// slots as the other wavefront kernels (kernels_grid.cuh), same sequence of IEEE operations per node (bit-identical state).
//
// Why: a pass in the reference order is a dependency chain of H + W - 1 node updates; its duration is W * tau_x + H * tau_y
// (tau_x / tau_y: latency of one update as seen by the row / column neighbour), and on one in-order warp tau is essentially the
// LENGTH OF THE INSTRUCTION STREAM of a step (measured ~3.5 cycles per instruction: dependent issue).  So a step must be as few
// instructions as possible:
//   * lanes 0-15 min-marginalise the edge to the ROW neighbour, lanes 16-31 the edge to the COLUMN neighbour -- one min-plus
//     instruction stream serves both (LPL = L/16 labels per lane); the two min-marginals meet through one xor-16 shuffle;
//   * the group minimum of a half-warp is redux.sync.min.f32 (CREDUX.MIN.F32, sm_100a; 19 cycles against 4 dependent
//     shuffle+min rounds, scripts/repro/lat.cu), the truncated-linear label window is 2*WIN independent 16-lane shuffles;
//   * steady steps are unrolled per tile of K = 8 columns: every shared-memory address is a per-tile base + an immediate, one
//     publication of the step counter per tile, no vote in the hand-off poll (every lane watches the same slot);
//   * the row neighbour's vector stays in the registers of the lower half-warp; the column neighbour's vector comes from the
//     warp above through a ring of self-validating slots in shared memory (next CTA: L2 hand-off slots, in-mover warp).
// CTA = NW = 8 compute warps (band of 8 consecutive rows) + one in-mover warp + one tile warp (all TMA traffic of the CTA).
#pragma once
#include "kernels_grid.cuh"
#include "kernels_grid_rows.cuh"

namespace lpmp {

// UNI: the instance for the uniform / damped-uniform modes (a node also receives from -- and sends to -- its two LATER neighbours):
// one more ring of tiles holds the records of the first row of the next band (whose in-side vectors the band's last row reads).
template <typename Real, int PW, int L, bool UNI = false>
struct SplitCfg {
  static constexpr int LPL = L / 16;                       // labels per lane (16 lanes per edge)
  static constexpr int NW = 8;                             // compute warps = rows per CTA
  static constexpr int T = NW;
  static constexpr int K = 8, LOGK = 3;                    // columns per tile
  static constexpr int P = ParamCount<PW>::value;
  static constexpr int REC = REC_VECS * L, PREC = PAR_EDGES * P;
  static constexpr int SLOT_REC = K * REC, SLOT_PAR = K * PREC;   // reals per tile (one row)
  static constexpr int RING = 32;                          // columns per hand-off ring (power of two)
  static constexpr size_t SLOT_BYTES = (size_t)(SLOT_REC + SLOT_PAR) * sizeof(Real);
  static constexpr size_t SLOT_REC_BYTES = (size_t)SLOT_REC * sizeof(Real);
  static constexpr size_t CTRL_OFF = 640;                  // [0,640) mbarriers ([NW + 1][NSLOT]), [640,768) band queue {super-band, iteration + 1} x 16,
  static constexpr size_t CTR_OFF = 768;                   // [768,1024) step counters: +4 w: compute warp w, +64: warp 0 of the next CTA of the cluster
  static constexpr size_t RINGS_OFF = 1024;
  static constexpr size_t TILES_OFF = (RINGS_OFF + (size_t)NW * RING * L * sizeof(Real) + 127) / 128 * 128;
  static constexpr int NSLOT = TILES_OFF + 8 * NW * SLOT_BYTES + (UNI ? 8 * SLOT_REC_BYTES : 0) <= 216 * 1024 ? 8 : 4;   // tiles per ring (power of two)
  static constexpr size_t WARP_TILE_BYTES = (size_t)NSLOT * SLOT_BYTES;
  static constexpr size_t HALO_OFF = TILES_OFF + NW * WARP_TILE_BYTES;          // UNI: [NSLOT] record tiles of the row below the band
  static constexpr size_t SMEM = HALO_OFF + (UNI ? (size_t)NSLOT * SLOT_REC_BYTES : 0);
  static constexpr int THREADS = 32 * (NW + 2);
  static constexpr bool FITS = (L == 16 || L == 32) && SMEM <= (UNI ? 226 : 220) * 1024 && (NW + 1) * NSLOT * 8 <= 640 && (PW == LPMP_PAIRWISE_POTTS || LPL == 1);
};

// ---- thread-block cluster / distributed shared memory primitives ----
__device__ __forceinline__ uint32_t sp_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t sp_cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t sp_mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void sp_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_n_if(bool on, uint32_t a, const float (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared::cluster.f32 [%0], %1; }" ::"r"(a), "f"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_cluster_n_if(bool on, uint32_t a, const float (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.shared::cluster.v2.f32 [%0], {%1,%2}; }" ::"r"(a), "f"(v[0]), "f"(v[1]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_cluster_n_if(bool on, uint32_t a, const double (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared::cluster.f64 [%0], %1; }" ::"r"(a), "d"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_cluster_n_if(bool on, uint32_t a, const double (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.shared::cluster.v2.f64 [%0], {%1,%2}; }" ::"r"(a), "d"(v[0]), "d"(v[1]), "r"((uint32_t)on) : "memory"); }

// minimum over the 16 lanes of each half-warp
template <typename Real>
__device__ __forceinline__ Real half_min(Real v) {
  if constexpr (std::is_same<Real, float>::value) {
    const bool lo = (threadIdx.x & 16u) == 0u;
    const float INF = RealTraits<float>::inf();
    const float a0 = credux_min(lo ? v : INF), a1 = credux_min(lo ? INF : v);
    return lo ? a0 : a1;
  } else {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = rmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
}

// h(x) = min_y [c(x,y) + o(y)] for the LPL labels x = (lane & 15) * LPL + k of this lane, o spread over the 16 lanes of the
// half-warp.  Truncated linear (LPL = 1; slope, trunc >= 0; host: slope * (WIN + 1) >= trunc): far terms collapse to trunc + min(o),
// near terms are 2 * WIN shuffles inside the half-warp (a shuffle that leaves it returns the lane's own value: c_d + o(x) >= o(x)
// is harmless).  Potts: WIN = 1: all lambda >= 0 (one minimum); WIN = 0: the two smallest entries.
template <typename Real, int PW, int LPL, int WIN>
__device__ __forceinline__ void split_min_plus(const Real* p, const Real (&o)[LPL], Real (&h)[LPL]) {
  constexpr unsigned FULL = 0xffffffffu;
  Real lm = o[0];
#pragma unroll
  for (int k = 1; k < LPL; ++k) lm = rmin(lm, o[k]);
  if constexpr (PW == LPMP_PAIRWISE_POTTS) {
    const Real lam = p[0];
    if (WIN == 0 && __any_sync(FULL, lam < Real(0))) {
      const Real INF = RealTraits<Real>::inf();
      const unsigned lane = threadIdx.x & 31u;
      const unsigned segmask = 0xffffu << (lane & 16u);
      const Real g = half_min<Real>(lm);
      const unsigned il = (unsigned)__ffs((int)(__ballot_sync(FULL, lm == g) & segmask)) - 1u;   // first lane holding the minimum
      int kf = 0;                                                                                    // ... and its first label there
#pragma unroll
      for (int k = LPL - 1; k >= 0; --k) if (o[k] == g) kf = k;
      const int k1 = __shfl_sync(FULL, kf, (int)il);
      Real sm = INF;
#pragma unroll
      for (int k = 0; k < LPL; ++k) sm = rmin(sm, lane == il && k == k1 ? INF : o[k]);
      const Real g2 = half_min<Real>(sm);
#pragma unroll
      for (int k = 0; k < LPL; ++k) h[k] = rmin(o[k], Real(lam + (lane == il && k == k1 ? g2 : g)));
    } else {
      const Real f = lam + half_min<Real>(lm);
#pragma unroll
      for (int k = 0; k < LPL; ++k) h[k] = rmin(o[k], f);
    }
  } else {
    static_assert(PW == LPMP_PAIRWISE_POTTS || LPL == 1, "truncated linear: one label per lane");
    const Real slope = p[0], trunc = p[1];
    Real up[WIN], dn[WIN];
#pragma unroll
    for (int d = 1; d <= WIN; ++d) {
      up[d - 1] = __shfl_up_sync(FULL, o[0], d, 16);
      dn[d - 1] = __shfl_down_sync(FULL, o[0], d, 16);
    }
    Real r = rmin(o[0], Real(trunc + half_min<Real>(lm)));
#pragma unroll
    for (int d = 1; d <= WIN; ++d) {
      const Real c = rmin(Real(slope * Real(d)), trunc);
      r = rmin(r, rmin(Real(c + up[d - 1]), Real(c + dn[d - 1])));
    }
    h[0] = r;
  }
}

// LPL consecutive reals of this lane (shared memory, 32-bit shared address)
template <typename Real, int N> __device__ __forceinline__ void lds_n(Real (&v)[N], uint32_t a);
template <> __device__ __forceinline__ void lds_n<float, 1>(float (&v)[1], uint32_t a) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a)); }
template <> __device__ __forceinline__ void lds_n<float, 2>(float (&v)[2], uint32_t a) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a)); }
template <> __device__ __forceinline__ void lds_n<double, 1>(double (&v)[1], uint32_t a) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(a)); }
template <> __device__ __forceinline__ void lds_n<double, 2>(double (&v)[2], uint32_t a) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a)); }
template <typename Real, int N> __device__ __forceinline__ void lds_n_volatile(Real (&v)[N], uint32_t a);
template <> __device__ __forceinline__ void lds_n_volatile<float, 1>(float (&v)[1], uint32_t a) { asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a) : "memory"); }
template <> __device__ __forceinline__ void lds_n_volatile<float, 2>(float (&v)[2], uint32_t a) { asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a) : "memory"); }
template <> __device__ __forceinline__ void lds_n_volatile<double, 1>(double (&v)[1], uint32_t a) { asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v[0]) : "r"(a) : "memory"); }
template <> __device__ __forceinline__ void lds_n_volatile<double, 2>(double (&v)[2], uint32_t a) { asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a) : "memory"); }
__device__ __forceinline__ void sts_n_if(bool on, uint32_t a, const float (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.volatile.shared.f32 [%0], %1; }" ::"r"(a), "f"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void sts_n_if(bool on, uint32_t a, const float (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.volatile.shared.v2.f32 [%0], {%1,%2}; }" ::"r"(a), "f"(v[0]), "f"(v[1]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void sts_n_if(bool on, uint32_t a, const double (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.volatile.shared.f64 [%0], %1; }" ::"r"(a), "d"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void sts_n_if(bool on, uint32_t a, const double (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.volatile.shared.v2.f64 [%0], {%1,%2}; }" ::"r"(a), "d"(v[0]), "d"(v[1]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void sts_n(uint32_t a, const float (&v)[1]) { asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(a), "f"(v[0]) : "memory"); }
__device__ __forceinline__ void sts_n(uint32_t a, const float (&v)[2]) { asm volatile("st.volatile.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v[0]), "f"(v[1]) : "memory"); }
__device__ __forceinline__ void sts_n(uint32_t a, const double (&v)[1]) { asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v[0]) : "memory"); }
__device__ __forceinline__ void sts_n(uint32_t a, const double (&v)[2]) { asm volatile("st.volatile.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v[0]), "d"(v[1]) : "memory"); }
// generic address of a shared::cluster address (a ring of the next CTA of the cluster): one generic store then serves both kinds
// of hand-off that leave the CTA -- distributed shared memory and the L2 / peer-memory slots
__device__ __forceinline__ void* sp_generic_of_cluster(uint32_t a) {
  uint64_t g;
  asm volatile("{ .reg .u64 t; cvt.u64.u32 t, %1; cvta.shared::cluster.u64 %0, t; }" : "=l"(g) : "r"(a));
  return reinterpret_cast<void*>(g);
}
__device__ __forceinline__ void st_generic_n_if(bool on, float* p, const float (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.relaxed.sys.f32 [%0], %1; }" ::"l"(p), "f"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_generic_n_if(bool on, float* p, const float (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.relaxed.sys.v2.f32 [%0], {%1,%2}; }" ::"l"(p), "f"(v[0]), "f"(v[1]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_generic_n_if(bool on, double* p, const double (&v)[1]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.relaxed.sys.f64 [%0], %1; }" ::"l"(p), "d"(v[0]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void st_generic_n_if(bool on, double* p, const double (&v)[2]) { asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.relaxed.sys.v2.f64 [%0], {%1,%2}; }" ::"l"(p), "d"(v[0]), "d"(v[1]), "r"((uint32_t)on) : "memory"); }
__device__ __forceinline__ void stg_n_relaxed_if(bool on, float* p, const float (&v)[1]) { stg1_relaxed_if(on, p, v[0]); }
__device__ __forceinline__ void stg_n_relaxed_if(bool on, double* p, const double (&v)[1]) { stg1_relaxed_if(on, p, v[0]); }
__device__ __forceinline__ void stg_n_relaxed_if(bool on, float* p, const float (&v)[2]) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.relaxed.sys.global.v2.f32 [%0], {%1,%2}; }" ::"l"(p), "f"(v[0]), "f"(v[1]), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void stg_n_relaxed_if(bool on, double* p, const double (&v)[2]) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p st.relaxed.sys.global.v2.f64 [%0], {%1,%2}; }" ::"l"(p), "d"(v[0]), "d"(v[1]), "r"((uint32_t)on) : "memory");
}

// minimum over the whole warp of a value that both half-warps hold (theta)
template <typename Real>
__device__ __forceinline__ Real both_halves_min(Real v) {
  if constexpr (std::is_same<Real, float>::value) return credux_min(v);
  else return half_min<Real>(v);
}

// LBF (backward passes only): LP::LowerBound (/root/reference/include/LP_MP.h:1507-1518) accumulated on the fly.  In a backward pass
// node i is the LAST writer of theta_i and of both vectors of its edges to the right / lower neighbour (that neighbour was updated
// earlier in the pass, its side b is the vector this step min-marginalises), and h(x) = min_y [c(x,y) + b(y)] is in registers: the
// node's term min theta_i and the two edge terms min_x [a(x) + h(x)] cost three group minima per step, off the dependency chain, and
// a pass leaves one double per grid row in A.lb_rows -- the separate sweep over all records (kernels_grid_bound.cuh) disappears.
// PRM: rounding passes (LP::ComputePassAndPrimal, include/LP_MP.h:1592-1602) on the fast steps: the first minimiser of theta after
// the receives (include/factors_messages.hxx:2363,2382-2389) is worked out after the stores of the step, off the dependency chain.
template <typename Real, int PW, int L, int WIN, bool FWD, bool LBF = false, bool PRM = false, bool UNI = false>
__global__ void __launch_bounds__(SplitCfg<Real, PW, L, UNI>::THREADS, 1) wavefront_split_kernel(const GridArgs<Real> A) {
  static_assert(!(LBF && FWD), "the fused bound belongs to the backward pass");
  static_assert(!(LBF && UNI), "the fused bound needs the anisotropic backward pass (every edge final after its later node)");
  using Cfg = SplitCfg<Real, PW, L, UNI>;
  constexpr int LPL = Cfg::LPL, NW = Cfg::NW, T = Cfg::T, K = Cfg::K, LOGK = Cfg::LOGK, P = Cfg::P, REC = Cfg::REC, PREC = Cfg::PREC, NSLOT = Cfg::NSLOT, RING = Cfg::RING;
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  constexpr uint32_t REC_B = REC * S, PREC_B = PREC * S, SLOT_REC_B = Cfg::SLOT_REC * S, SLOT_PAR_B = Cfg::SLOT_PAR * S, LB = L * S;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  constexpr bool fwd = FWD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bars = smem_base;                         // [NW][NSLOT] mbarriers
  const uint32_t queue = smem_base + (uint32_t)Cfg::CTRL_OFF;    // band queue: {super-band, iteration + 1} x 16
  const uint32_t ctrs = smem_base + (uint32_t)Cfg::CTR_OFF;      // +4 w: steps completed by compute warp w (monotone over the kernel); +64: pushed by the next CTA's warp 0
  const uint32_t rings = smem_base + (uint32_t)Cfg::RINGS_OFF;   // [NW][RING][L]: ring w feeds compute warp w with the vectors of the row above
  const uint32_t tiles = smem_base + (uint32_t)Cfg::TILES_OFF;   // per compute warp: [NSLOT] record tiles, then [NSLOT] parameter tiles
  const int lane = threadIdx.x & 31, wid = (int)(threadIdx.x >> 5);
  const uint32_t H = A.H, W = pin(A.W);
  const int nchunks = ((int)W + K - 1) >> LOGK;
  const int nb = (int)A.num_bands;
  const int qoff = fwd ? 0 : nchunks * K - (int)W;         // padding columns / rows in front of virtual column / row 0
  const int roff = fwd ? 0 : nb * T - (int)H;              // (tiles are anchored at row 0 / column 0 in both directions)
  const int nsteps = (int)W;
  constexpr int kxor = fwd ? 0 : K - 1;
  // ring slot of a column: its tile-space index q' = q + qoff modulo RING (tiles and ring slots wrap together)

  if (threadIdx.x == 0) {
    for (int i = 0; i < (NW + 1) * NSLOT; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    Real* rg = reinterpret_cast<Real*>(smem_raw + Cfg::RINGS_OFF);
    for (int i = (int)threadIdx.x; i < NW * RING * L; i += Cfg::THREADS) rg[i] = sentinel<Real>();
    if (threadIdx.x < (Cfg::RINGS_OFF - Cfg::CTRL_OFF) / 4) reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::CTRL_OFF)[threadIdx.x] = 0u;   // band queue + step counters
  }
  // CTAs of a thread-block cluster work on CONSECUTIVE bands (rank r on band sb * csize + r): the last row of band r hands its
  // vectors straight into ring 0 of rank r + 1 through distributed shared memory (~0.2 us instead of ~0.8 us through the L2
  // slots and the in-mover); only the boundary between two clusters goes through L2
  const uint32_t crank = sp_cluster_ctarank(), csize = sp_cluster_nctarank();
  const unsigned num_super = (A.num_bands + csize - 1u) / csize;
  if (csize > 1) sp_cluster_sync();   // every ring of the cluster is armed before a neighbour may write into it
  else __syncthreads();

  // bounded waits: a lost partner must not hang the device (sticky abort flag, checked by the host after the pass)
  const unsigned long long t_launch = globaltimer_ns();
  auto timed_out = [&](unsigned& spins, unsigned) {
    if ((++spins & 0x3ffu) != 0) return false;
    return wait_expired(A.control, t_launch);
  };
  auto warp_live = [&](unsigned kb, int w) {               // does compute warp w own a grid row in virtual band kb?
    const int u = (int)kb * T + w - roff;
    return kb < A.num_bands && u >= 0 && u < (int)H;
  };

  // UNI: is there a row below virtual band kb whose records the halo ring holds?  Inside the grid, or -- stripes (whole bands only,
  // A.later_boundary set by the host) -- the ghost row next to the stripe's last virtual row (later_gather_kernel filled it)
  auto halo_live = [&](unsigned kb) {
    return warp_live(kb, NW) || (A.later_boundary != nullptr && kb + 1u == A.num_bands && (int)kb * T + NW - roff == (int)H);
  };

  uint32_t step_base = 0;                 // steps of earlier bands: the published step counters are monotone
  uint32_t chunk_base[NW];                // per compute warp: tiles of earlier bands (slot / mbarrier-phase sequence continues)
#pragma unroll
  for (int w = 0; w < NW; ++w) chunk_base[w] = 0;
  uint32_t chunk_base_halo = 0;           // UNI: ... and of the ring of the row below the band
  const uint32_t halo_tiles = smem_base + (uint32_t)Cfg::HALO_OFF;

  for (uint32_t iter = 0;; ++iter) {
    // rank 0 claims the next super-band and publishes it to every CTA of the cluster (a rank is at most csize - 1 iterations behind)
    const uint32_t qslot = queue + 8 * (iter & 15u);
    if (threadIdx.x == 0) {
      if (crank == 0) {
        const unsigned sb0 = atomicAdd(A.control, 1u);
        for (uint32_t rk = 0; rk < csize; ++rk)
          asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(sp_mapa(qslot, rk)), "r"(sb0), "r"(iter + 1u) : "memory");
      }
      unsigned sbq, it1, spins = 0;
      do {
        asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sbq), "=r"(it1) : "r"(qslot) : "memory");
      } while (it1 != iter + 1u && !timed_out(spins, 1u << 27));
      if (it1 != iter + 1u) sts_volatile_u32(qslot, 0xffffffffu);   // timed out: leave
    }
    __syncthreads();
    const unsigned sb = lds_volatile_u32(qslot);
    if (sb >= num_super) break;
    const unsigned kb = sb * csize + crank;                // may be >= num_bands in the last super-band: no live warp then
    const int t_in = kb == 0 ? roff : 0;                   // first grid row of the band (virtual row inside the band = compute warp)
    const bool band_in = kb > 0 || A.in_halo;              // row t_in has a column neighbour in the previous band / stripe

    if (wid == NW + 1) {
      // ================= tile warp: bulk loads as slots become free, bulk stores as the compute warps complete tiles =================
      // LANE w serves compute warp w (its own ring of NSLOT tiles, its own bulk groups): the bookkeeping of all rings is one SIMT
      // instruction stream -- a single thread serving every ring in turn was the bottleneck of the whole CTA (profiles/r02_split_tilewarp.txt)
      constexpr int DROP = NSLOT >= 8 ? 2 : 1;   // tiles of look-ahead given up so that a load never waits for the store just issued
      // UNI: lane NW serves the ring of the row below the band (records only, loads only; a slot is free again when the band's last
      // row -- its only reader -- has left the tile)
      const bool halo_lane = UNI && lane == NW;
      const int w = lane < NW ? lane : NW - 1;
      bool open = (lane < NW && warp_live(kb, w)) || (halo_lane && halo_live(kb));
      uint32_t cb = halo_lane ? chunk_base_halo : 0u;
#pragma unroll
      for (int k = 0; k < NW; ++k) cb = (k == w && !halo_lane) ? chunk_base[k] : cb;
      const int wr = halo_lane ? NW : w;       // row of the band this lane moves
      const int row0 = (int)((fwd ? kb : (unsigned)nb - 1u - kb) * T) + (fwd ? wr : NW - 1 - wr);
      const uint32_t rec_ring = halo_lane ? halo_tiles : tiles + (uint32_t)w * (uint32_t)Cfg::WARP_TILE_BYTES, par_ring = rec_ring + NSLOT * SLOT_REC_B;
      const uint32_t ctr_w = ctrs + 4 * (uint32_t)w;
      const bool mine = open;
      int j_load = 0, j_store = 0;
      unsigned spins = 0;
      while (__any_sync(FULL, open)) {
        bool moved = false;
        // ---- stores: a tile is final when the last row of the warp has left it ----
        const int c = (int)(lds_volatile_u32(ctr_w) - step_base);
        int need = (j_store + 1) * K - qoff;
        if (need > nsteps) need = nsteps;
        if (open && halo_lane && c >= need) { ++j_store; moved = true; }   // (nothing to store: the reader has left the tile)
        if (open && !halo_lane && c >= need) {
          const uint32_t slot = (cb + (uint32_t)j_store) & (NSLOT - 1);
          const int col0 = fwd ? j_store * K : (nchunks - 1 - j_store) * K;
          const uint32_t ncols = (uint32_t)((int)W - col0 < K ? (int)W - col0 : K);   // a tile is one contiguous run of records
          bulk_store_1d(A.rec + ((uint64_t)row0 * W + (uint32_t)col0) * REC, rec_ring + slot * SLOT_REC_B, ncols * REC_B);
          tma_commit();
          ++j_store;
          moved = true;
        }
        // ---- loads: into a slot whose previous tile has left shared memory (never the slot of the store just issued) ----
        if (open && j_load < nchunks && j_load < j_store + NSLOT - DROP) {
          const uint32_t slot = (cb + (uint32_t)j_load) & (NSLOT - 1);
          if (j_load >= NSLOT && !halo_lane) tma_wait_read<DROP>();   // at least DROP younger stores of this lane were committed after the one that read the slot
          const int col0 = fwd ? j_load * K : (nchunks - 1 - j_load) * K;
          const uint32_t bar = bars + 8 * ((uint32_t)wr * NSLOT + slot);
          const uint32_t ncols = (uint32_t)((int)W - col0 < K ? (int)W - col0 : K);
          const uint64_t node0 = (uint64_t)row0 * W + (uint32_t)col0;
          mbar_expect_tx(bar, ncols * (REC_B + (halo_lane ? 0u : PREC_B)));
          bulk_load_1d(rec_ring + slot * SLOT_REC_B, A.rec + node0 * REC, ncols * REC_B, bar);
          if (!halo_lane) bulk_load_1d(par_ring + slot * SLOT_PAR_B, A.par + node0 * PREC, ncols * PREC_B, bar);
          ++j_load;
          moved = true;
        }
        open = open && j_store < nchunks;
        if (__any_sync(FULL, moved)) spins = 0;
        else { __nanosleep((A.lean_flags & 8) ? 1000 : 100); if (timed_out(spins, 1u << 26)) break; }
      }
      if (mine) tma_wait_read<0>();   // every tile of this band has left shared memory before the rings are reused
      __syncwarp();
    } else if (wid == NW) {
      // ====== in-mover: predecessor's hand-off slots (L2) -> ring of the band's first grid row (zeros if there is no predecessor) ======
      constexpr int G4 = L / 4, TG = 32 / G4;               // lanes per column (4 labels each), columns per load
      constexpr int NBATCH = TG >= 16 ? 1 : 16 / TG;        // >= 16 columns polled per L2 round trip
      const int g = lane / G4, gl = lane % G4;
      const int w_in = t_in;                                // the compute warp that owns row t_in
      const uint32_t ring = rings + (uint32_t)w_in * RING * LB;
      const uint32_t ctr_in = ctrs + 4 * (uint32_t)w_in;
      int q_in = 0;
      unsigned spins = 0;
      if (kb >= A.num_bands || (kb > 0 && crank > 0)) {
        // no band, or the previous band runs in the previous CTA of this cluster and writes into ring 0 itself
      } else if (!band_in) {
        Real zero[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) zero[k] = Real(0);
        while (q_in < (int)W) {
          const int c = (int)(lds_volatile_u32(ctr_in) - step_base);
          int limit = c + RING;
          if (limit > (int)W) limit = (int)W;
          if (q_in < limit) {
            for (int q = q_in + g; q < limit; q += TG) sts4(ring + (((uint32_t)(q + qoff) & (uint32_t)(RING - 1)) * L + (uint32_t)(gl * 4)) * S, zero);
            q_in = limit;
            spins = 0;
          } else { __nanosleep((A.lean_flags & 16) ? 2000 : 200); if (timed_out(spins, 1u << 26)) break; }
        }
      } else {
        const Real* const hand_in = (kb == 0 ? A.in_boundary : A.handoff + (uint64_t)kb * W * L) + gl * 4;
        while (q_in < (int)W) {
          const int c = (int)(lds_volatile_u32(ctr_in) - step_base);
          int limit = c + RING;                             // the consumer has consumed (and re-armed) c columns
          if (limit > (int)W) limit = (int)W;
          Real v[NBATCH][4];
#pragma unroll
          for (int b = 0; b < NBATCH; ++b) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[b][k] = sentinel<Real>();
            const int q = q_in + b * TG + g;
            ld_relaxed_if<HAND_BYTES>(q < limit, v[b], hand_in + (int64_t)q * L);
          }
          int done = 0;
          bool open = true;   // the ring fills in column order: stop at the first column that has not arrived
#pragma unroll
          for (int b = 0; b < NBATCH; ++b) {
            const int q = q_in + b * TG + g;
            const bool ok = q < limit && !(is_sentinel(v[b][0]) || is_sentinel(v[b][1]) || is_sentinel(v[b][2]) || is_sentinel(v[b][3]));
            const unsigned bad = ~__ballot_sync(FULL, ok);
            const int n_ok = !open ? 0 : bad ? (__ffs((int)bad) - 1) / G4 : TG;
            if (g < n_ok) {
              sts4(ring + (((uint32_t)(q + qoff) & (uint32_t)(RING - 1)) * L + (uint32_t)(gl * 4)) * S, v[b]);
              Real sent[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) sent[k] = sentinel<Real>();
              st_relaxed<HAND_BYTES>(const_cast<Real*>(hand_in) + (int64_t)q * L, sent);   // the slot is free again for the next pass
            }
            done += n_ok;
            open = open && n_ok == TG;
          }
          q_in += done;
          if (done > 0) spins = 0;
          else if (timed_out(spins, 1u << 24)) break;
        }
      }
    } else if (!warp_live(kb, wid)) {
      // no grid row in this band (padding rows of the first / last band): the step counter still moves on with the band
      if (lane == 0) {
        sts_volatile_u32(ctrs + 4 * (uint32_t)wid, step_base + (uint32_t)nsteps);
        if (wid == 0 && crank > 0) st_cluster_u32(sp_mapa(ctrs + 64, crank - 1), step_base + (uint32_t)nsteps);
      }
    } else {
      // =================================================== compute warps ===================================================
      const int w = wid;
      const bool hi = lane >= 16;                           // 0: the edge to the row neighbour, 1: the edge to the column neighbour
      const int xl = (lane & 15) * LPL;                     // first label of this lane
      const int u = (int)kb * T + w - roff;                 // virtual grid row of this warp
      const uint32_t r = fwd ? (uint32_t)u : H - 1 - (uint32_t)u;
      const bool in_col_exists = u > 0 || A.in_halo, out_col_exists = u + 1 < (int)H || A.out_halo;
      const bool grid_last_row = u + 1 == (int)H;
      // where the outgoing column vector (upper half-warp) goes: next warp (ring), next band (L2 hand-off slots), next stripe
      const bool to_next_cta = !grid_last_row && w + 1 == NW && crank + 1 < csize;   // the next band runs in the next CTA of the cluster
      const bool out_ring = hi && !grid_last_row && w + 1 < NW;
      const bool out_dsm = hi && to_next_cta;
      const bool out_glob = hi && (grid_last_row ? (A.out_halo && A.out_boundary != nullptr) : (w + 1 == NW && !to_next_cta));
      Real* const out_ptr = (grid_last_row ? A.out_boundary : A.handoff + (uint64_t)(kb + 1) * W * L) + xl;
      const uint32_t ring_in = rings + (uint32_t)w * RING * LB + (uint32_t)xl * S;
      const uint32_t ring_out = rings + (uint32_t)(w + 1 < NW ? w + 1 : w) * RING * LB + (uint32_t)xl * S;
      const uint32_t ring_dsm = sp_mapa(rings + (uint32_t)xl * S, to_next_cta ? crank + 1 : crank);   // ring 0 of the next CTA
      const bool warp_out_ring = !grid_last_row && (w + 1 < NW || to_next_cta);
      const uint32_t rec_ring = tiles + (uint32_t)w * (uint32_t)Cfg::WARP_TILE_BYTES, par_ring = rec_ring + NSLOT * SLOT_REC_B;
      constexpr uint32_t o_in_row = (fwd ? REC_BL : REC_AR) * LB, o_in_col = (fwd ? REC_BU : REC_AD) * LB;
      constexpr uint32_t o_out_row = (fwd ? REC_AR : REC_BL) * LB, o_out_col = (fwd ? REC_AD : REC_BU) * LB;
      constexpr uint32_t p_in_row = (fwd ? PAR_L : PAR_R) * P * S, p_in_col = (fwd ? PAR_U : PAR_D) * P * S;
      // per-lane record fields: theta (both halves), own side of this half's receive edge, own side of this half's send edge
      const uint32_t f_th = rec_ring + (uint32_t)xl * S;
      const uint32_t f_in = f_th + (hi ? o_in_col : o_in_row), f_out = f_th + (hi ? o_out_col : o_out_row);
      const uint32_t f_par = par_ring + (hi ? p_in_col : p_in_row);
      const bool recv_row_first = pin((uint32_t)(fwd ? !A.fifo : (A.fifo != 0))) != 0u;
      const bool self_first = hi != recv_row_first;         // this half's min-marginal is added to theta first
      const Real w_fast = (Real)(1.0 / 2.0);
      const uint32_t cbase = chunk_base[w];
      const uint32_t ctr = ctrs + 4 * (uint32_t)w;
      const uint32_t ctr_next = w + 1 < NW ? ctrs + 4 * (uint32_t)(w + 1) : ctrs + 64;   // the consumer of this row's vectors
      const bool push_progress = w == 0 && crank > 0;      // ... and this warp is the consumer of the previous CTA's last row
      const uint32_t ctr_push = sp_mapa(ctrs + 64, push_progress ? crank - 1 : crank);

      auto wait_chunk = [&](int j) {
        const uint32_t n = cbase + (uint32_t)j;
        const uint32_t bar = bars + 8 * ((uint32_t)w * NSLOT + (n & (NSLOT - 1)));
        const uint32_t parity = (n / NSLOT) & 1u;
        unsigned spins = 0;
        while (!mbar_try_wait(bar, parity)) { if (timed_out(spins, 1u << 24)) break; }
      };
      auto publish = [&](int steps_done) {
        fence_proxy_async();   // the tile warp's TMA store reads completed tiles through the async proxy
        __syncwarp();
        asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0; @p st.volatile.shared.u32 [%1], %2; }" ::"r"((uint32_t)lane), "r"(ctr), "r"(step_base + (uint32_t)steps_done) : "memory");
        if (push_progress && lane == 0) st_cluster_u32(ctr_push, step_base + (uint32_t)steps_done);
      };

      Real oth[LPL];            // lower half: the row neighbour's vector (carried from step to step); upper half: the column neighbour's
#pragma unroll
      for (int k = 0; k < LPL; ++k) oth[k] = Real(0);
      int free_upto = to_next_cta ? 0 : RING;   // columns the consumer's ring is known to have room for (another CTA: ask first)
      double lb_acc = 0.0;      // LBF: bound terms of this row (lane 0's order of additions is the one that counts)
      // the edge to the stripe below is final only when BOTH stripes have finished their pass: left to the host (api.cu)
      const bool lb_skip_col = u == 0 && A.in_halo;
      auto lb_terms = [&](const Real (&th_new)[LPL], const Real (&own_new)[LPL], const Real (&h)[LPL], const bool has_edge) {
        Real tm = th_new[0], em = own_new[0] + h[0];
#pragma unroll
        for (int k = 1; k < LPL; ++k) { tm = rmin(tm, th_new[k]); em = rmin(em, Real(own_new[k] + h[k])); }
        em = has_edge ? em : RealTraits<Real>::inf();
        const Real gt = both_halves_min<Real>(tm);
        Real ge = half_min<Real>(em);
        ge = (has_edge && !(hi && lb_skip_col)) ? ge : Real(0);
        const Real go = __shfl_xor_sync(FULL, ge, 16);
        lb_acc += ((double)gt + (double)ge) + (double)go;
      };
      // PRM: label of the node of step s from theta after the receives (both halves hold it; the lower half's ballot decides)
      auto label_of = [&](const Real (&tv)[LPL]) -> unsigned {     // (the same value in every lane)
        Real lm = tv[0];
#pragma unroll
        for (int k = 1; k < LPL; ++k) lm = rmin(lm, tv[k]);
        const Real g = half_min<Real>(lm);
        const unsigned il = (unsigned)__ffs((int)(__ballot_sync(FULL, lm == g) & 0xffffu)) - 1u;
        int kf = 0;
#pragma unroll
        for (int k = LPL - 1; k >= 0; --k) if (tv[k] == g) kf = k;
        const int k1 = __shfl_sync(FULL, kf, (int)il);
        return il * LPL + (unsigned)k1;
      };
      auto write_label = [&](const Real (&tv)[LPL], const int s) {
        const unsigned lab = label_of(tv);
        if (lane == 0) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)s : W - 1 - (uint32_t)s)] = lab;
      };
      // tiles: lane jj keeps the label of step s0 + jj (a select, no branch in the step); one store of K labels after the tile
      unsigned tile_labels = 0u;
      auto keep_label = [&](const Real (&tv)[LPL], const int jj) {
        const unsigned lab = label_of(tv);
        tile_labels = lane == jj ? lab : tile_labels;
      };
      auto write_tile_labels = [&](const int s0) {
        if (lane < K) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)(s0 + lane) : W - 1 - (uint32_t)(s0 + lane))] = tile_labels;
      };
      if constexpr (UNI) {
        // ============ uniform / damped-uniform modes (include/LP_MP.h:1420-1449, full receive mask): a node receives from all four
        // neighbours and sends to all four with weight 1 / deg (1 / (deg + 1)).  The two LATER neighbours have not been updated in
        // this pass: their in-side vectors are read from shared memory -- the next column of this row's own tile, and the same
        // column of the tile of the row below (the next warp's ring; the halo ring for the band's last row) -- and their
        // min-marginals are off the dependency chain.  Host contract: W % 8 == 0, W >= 16, stripes of whole bands; every row is tiles only:
        // FIRST / LAST mark the tiles of column 0 / W - 1 (no row edge on that side: nothing is sent there, one edge less in deg).
        const bool has_ic = in_col_exists, has_oc = out_col_exists;
        const int damp = A.repam == LPMP_REPAM_DAMPED_UNIFORM ? 1 : 0;
        const int deg_int = 2 + (int)has_ic + (int)has_oc;
        const Real w_int = (Real)(1.0 / (double)(deg_int + damp)), w_edge = (Real)(1.0 / (double)(deg_int - 1 + damp));
        // order of the additions into theta (message order of the node, grid coordinates: LIFO D, R, L, U; FIFO R, D, U, L): the
        // pair of the later edges first in a forward pass, the pair of the earlier edges first in a backward pass; inside a pair
        // the row edge first iff ... (earlier pair: recv_row_first; later pair: the opposite)
        const int srcB1 = (lane & 15) | (recv_row_first ? 0 : 16), srcA1 = (lane & 15) | (recv_row_first ? 16 : 0);
        constexpr uint32_t p_out_row = (fwd ? PAR_R : PAR_L) * P * S, p_out_col = (fwd ? PAR_D : PAR_U) * P * S;
        const uint32_t f_parA = par_ring + (hi ? p_out_col : p_out_row);
        // the row below: the next warp's ring of tiles, or the halo ring
        const bool below_halo = w + 1 == NW;
        uint32_t cb_below = chunk_base_halo;
#pragma unroll
        for (int k = 1; k < NW; ++k) cb_below = (k == w + 1) ? chunk_base[k] : cb_below;
        const uint32_t below_ring = below_halo ? halo_tiles : tiles + (uint32_t)(w + 1) * (uint32_t)Cfg::WARP_TILE_BYTES;
        const uint32_t below_bars = bars + 8 * (uint32_t)(w + 1) * NSLOT;
        auto wait_below = [&](int j) {
          const uint32_t n = cb_below + (uint32_t)j;
          const uint32_t bar = below_bars + 8 * (n & (NSLOT - 1));
          const uint32_t parity = (n / NSLOT) & 1u;
          unsigned spins = 0;
          while (!mbar_try_wait(bar, parity)) { if (timed_out(spins, 1u << 24)) break; }
        };
        Real* const out_gen = to_next_cta ? reinterpret_cast<Real*>(sp_generic_of_cluster(ring_dsm)) : out_ptr;
        const bool out_gen_on = out_dsm || out_glob;
        const bool ring_row = !grid_last_row && w + 1 < NW;

        auto uni_tile = [&](auto first_tag, auto last_tag, auto ring_tag, const int j) {
          constexpr bool FIRST = decltype(first_tag)::value, LAST = decltype(last_tag)::value, RINGOUT = decltype(ring_tag)::value;
          wait_chunk(j);
          if (!LAST) wait_chunk(j + 1);                     // the right neighbour of the tile's last column
          if (has_oc) wait_below(j);
          const int s0 = j << LOGK;
          const uint32_t slot = (cbase + (uint32_t)j) & (NSLOT - 1), slot_n = (cbase + (uint32_t)j + 1u) & (NSLOT - 1);
          const uint32_t col_first = fwd ? 0u : (uint32_t)(K - 1) * REC_B;   // first column of a tile in processing order
          const uint32_t toff = slot * SLOT_REC_B + col_first;
          const uint32_t Tth = f_th + toff, Tin = f_in + toff, Tout = f_out + toff;
          const uint32_t poff = slot * SLOT_PAR_B + (fwd ? 0u : (uint32_t)(K - 1) * PREC_B);
          const uint32_t TparB = f_par + poff, TparA = f_parA + poff;
          // the later neighbour's in-side vector: lower half: the next column of this row (own tile; the next tile for the last
          // column), upper half: the same column of the row below
          const uint32_t slot_b = (cb_below + (uint32_t)j) & (NSLOT - 1);
          const uint32_t own_next = rec_ring + (uint32_t)xl * S + o_in_row;
          const uint32_t below = (has_oc ? below_ring + slot_b * SLOT_REC_B : rec_ring + slot * SLOT_REC_B) + (uint32_t)xl * S + o_in_col + col_first;
          const uint32_t Tlat = hi ? below : own_next + toff + (fwd ? REC_B : 0u - REC_B);
          const uint32_t Tlat_last = hi ? below + (fwd ? (uint32_t)(K - 1) * REC_B : 0u - (uint32_t)(K - 1) * REC_B)
                                        : own_next + (LAST ? toff : slot_n * SLOT_REC_B + col_first);
          const uint32_t cq = (uint32_t)(j << LOGK) & (uint32_t)(RING - 1);
          const uint32_t Rin = ring_in + cq * LB, Rout = ring_out + cq * LB;
          Real* const Gout = out_gen + (to_next_cta ? (int64_t)cq * L : (int64_t)s0 * L);
          if (warp_out_ring && s0 + K - 1 >= free_upto) {   // room in the next warp's ring for the whole tile (warp-uniform, rare)
            unsigned spins = 0;
            do {
              free_upto = (int)(lds_volatile_u32(ctr_next) - step_base) + RING;
            } while (s0 + K - 1 >= free_upto && !timed_out(spins, 1u << 27));
          }
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            constexpr int SG = fwd ? 1 : -1;
            const bool e_ir = !(FIRST && jj == 0), e_or = !(LAST && jj == K - 1);   // (folded after unrolling)
            const uint32_t ro = (uint32_t)(SG * jj * (int)REC_B), po = (uint32_t)(SG * jj * (int)PREC_B);
            const uint32_t rin = Rin + (uint32_t)jj * LB;
            Real pv[LPL], parB[P], parA[P], th[LPL], oi[LPL], oo[LPL], ot[LPL];
            lds_n_volatile<Real, LPL>(pv, rin);             // the polled slot: every lane reads the words of its labels (no vote)
            lds_par<Real, P>(parB, TparB + po);
            lds_par<Real, P>(parA, TparA + po);
            lds_n<Real, LPL>(ot, jj + 1 < K ? Tlat + ro : Tlat_last);
            lds_n<Real, LPL>(th, Tth + ro);
            lds_n<Real, LPL>(oi, Tin + ro);
            lds_n<Real, LPL>(oo, Tout + ro);
            // later edges first: nothing of this depends on the row above or on the previous column
            const bool e_A = hi ? has_oc : e_or;
#pragma unroll
            for (int k = 0; k < LPL; ++k) ot[k] = e_A ? ot[k] : Real(0);
#pragma unroll
            for (int p = 0; p < P; ++p) parA[p] = e_A ? parA[p] : Real(0);
            Real hA[LPL], mA[LPL];
            split_min_plus<Real, PW, LPL, WIN>(parA, ot, hA);
#pragma unroll
            for (int k = 0; k < LPL; ++k) { mA[k] = e_A ? Real(hA[k] + oo[k]) : Real(0); oo[k] = e_A ? Real(oo[k] - mA[k]) : oo[k]; }
            if (is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) {
              unsigned spins = 0;
              do { lds_n_volatile<Real, LPL>(pv, rin); } while ((is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) && ((++spins & 0xffffu) != 0 || !wait_expired(A.control, t_launch)));
              if (is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) {
#pragma unroll
                for (int k = 0; k < LPL; ++k) pv[k] = Real(0);
              }
            }
            {
              Real sent[LPL];
#pragma unroll
              for (int k = 0; k < LPL; ++k) sent[k] = sentinel<Real>();
              sts_n(rin, sent);                             // the slot is free again (both halves: same words)
            }
#pragma unroll
            for (int k = 0; k < LPL; ++k) oth[k] = hi ? pv[k] : oth[k];
            Real hB[LPL], mB[LPL], tv[LPL];
            split_min_plus<Real, PW, LPL, WIN>(parB, oth, hB);
#pragma unroll
            for (int k = 0; k < LPL; ++k) { mB[k] = hB[k] + oi[k]; oi[k] = oi[k] - mB[k]; }
            const Real wq = e_ir && e_or ? w_int : w_edge;
#pragma unroll
            for (int k = 0; k < LPL; ++k) {
              const Real a1 = __shfl_sync(FULL, mA[k], srcA1), a2 = __shfl_sync(FULL, mA[k], srcA1 ^ 16);
              const Real b1 = __shfl_sync(FULL, mB[k], srcB1), b2 = __shfl_sync(FULL, mB[k], srcB1 ^ 16);
              const Real v = fwd ? Real(Real(Real(th[k] + a1) + a2) + b1) + b2 : Real(Real(Real(th[k] + b1) + b2) + a1) + a2;
              const Real d = wq * v;
              const Real d_ir = e_ir ? d : Real(0), d_or = e_or ? d : Real(0);
              const Real d_ic = has_ic ? d : Real(0), d_oc = (RINGOUT || has_oc) ? d : Real(0);
              oi[k] = oi[k] + (hi ? d_ic : d_ir);
              oo[k] = oo[k] + (hi ? d_oc : d_or);
              th[k] = Real(Real(Real(v - d_ir) - d_or) - d_ic) - d_oc;
              tv[k] = v;
            }
            // hand-off first: the row below is waiting for it
            if constexpr (RINGOUT) sts_n_if(hi, Rout + (uint32_t)jj * LB, oo);
            else st_generic_n_if(out_gen_on, Gout + jj * L, oo);
            sts_n(Tth + ro, th);                            // (both halves hold theta)
            sts_n(Tin + ro, oi);
            sts_n(Tout + ro, oo);
#pragma unroll
            for (int k = 0; k < LPL; ++k) oth[k] = oo[k];
            if constexpr (PRM) keep_label(tv, jj);             // after the stores: nothing of the chain waits for the label
          }
          publish(s0 + K);
          if constexpr (PRM) write_tile_labels(s0);
        };
        const int j_last = nchunks - 1;
        if (ring_row) {
          uni_tile(std::true_type{}, std::false_type{}, std::true_type{}, 0);
#pragma unroll 1
          for (int j = 1; j < j_last; ++j) uni_tile(std::false_type{}, std::false_type{}, std::true_type{}, j);
          uni_tile(std::false_type{}, std::true_type{}, std::true_type{}, j_last);
        } else {
          uni_tile(std::true_type{}, std::false_type{}, std::false_type{}, 0);
#pragma unroll 1
          for (int j = 1; j < j_last; ++j) uni_tile(std::false_type{}, std::false_type{}, std::false_type{}, j);
          uni_tile(std::false_type{}, std::true_type{}, std::false_type{}, j_last);
        }
      } else {
      wait_chunk(0);

      // ---- single steps: kind 0: general (any column, rounding, corner weights); 1: weight 1/2, any column (existence of the row
      // edges by selects); 2: interior column ----
      auto do_step = [&](auto kind_tag, const int s) {
        constexpr int KIND = decltype(kind_tag)::value;
        const int sp = s + qoff;
        if ((sp & (K - 1)) == 0 && s > 0) wait_chunk(sp >> LOGK);
        const uint32_t qq = (cbase << LOGK) + (uint32_t)sp;
        const uint32_t slot = (qq >> LOGK) & (NSLOT - 1), kk = (qq & (K - 1)) ^ (uint32_t)kxor;
        const uint32_t roffs = slot * SLOT_REC_B + kk * REC_B;
        const uint32_t rslot = ((uint32_t)sp & (uint32_t)(RING - 1)) * LB;
        Real par[P];
        lds_par<Real, P>(par, f_par + slot * SLOT_PAR_B + kk * PREC_B);
        Real th[LPL], oi[LPL], oo[LPL], pv[LPL];
        lds_n<Real, LPL>(th, f_th + roffs);
        lds_n<Real, LPL>(oi, f_in + roffs);
        lds_n<Real, LPL>(oo, f_out + roffs);
        {   // the column neighbour's vector: every lane watches the slot (identical outcome in every lane)
          unsigned spins = 0;
          bool missing;
          do {
            lds_n_volatile<Real, LPL>(pv, ring_in + rslot);
            missing = is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1]);
            if (missing && timed_out(spins, 1u << 27)) {
#pragma unroll
              for (int k = 0; k < LPL; ++k) pv[k] = Real(0);
              missing = false;
            }
          } while (__any_sync(FULL, missing));
        }
        {
          Real sent[LPL];
#pragma unroll
          for (int k = 0; k < LPL; ++k) sent[k] = sentinel<Real>();
          sts_n_if(hi, ring_in + rslot, sent);             // the slot is free again
        }
#pragma unroll
        for (int k = 0; k < LPL; ++k) oth[k] = hi ? pv[k] : oth[k];
        const bool has_ir = s > 0, has_or = s + 1 < (int)W;

        if constexpr (KIND >= 1) {
          Real h[LPL], m[LPL], mo[LPL];
          split_min_plus<Real, PW, LPL, WIN>(par, oth, h);
#pragma unroll
          for (int k = 0; k < LPL; ++k) { m[k] = h[k] + oi[k]; oi[k] = oi[k] - m[k]; }
#pragma unroll
          for (int k = 0; k < LPL; ++k) mo[k] = __shfl_xor_sync(FULL, m[k], 16);
          // missing edges: the left column and the first row have all-zero in-side slots and parameters, so their receives are
          // exact no-ops; the last row / column has no out_col / out_row edge (select)
          const bool or_on = KIND == 2 || has_or;
          Real tv[LPL];
#pragma unroll
          for (int k = 0; k < LPL; ++k) {
            const Real v = (th[k] + (self_first ? m[k] : mo[k])) + (self_first ? mo[k] : m[k]);
            const Real d = w_fast * v;
            const Real d_oc = out_col_exists ? d : Real(0), d_or = or_on ? d : Real(0);
            oo[k] = oo[k] + (hi ? d_oc : d_or);
            th[k] = (v - d_oc) - d_or;
            tv[k] = v;
          }
          if constexpr (PRM) write_label(tv, s);
          if constexpr (LBF) lb_terms(th, oi, h, true);
        } else {
          const bool has_ic = in_col_exists, has_oc = out_col_exists;
          const bool has_i = hi ? has_ic : has_ir;
          const int n_in = (int)has_ir + (int)has_ic, n_out = (int)has_or + (int)has_oc;
          Real pz[P], oz[LPL], h[LPL], m[LPL], mo[LPL];
#pragma unroll
          for (int p = 0; p < P; ++p) pz[p] = has_i ? par[p] : Real(0);
#pragma unroll
          for (int k = 0; k < LPL; ++k) oz[k] = has_i ? oth[k] : Real(0);
          split_min_plus<Real, PW, LPL, WIN>(pz, oz, h);
          // m = h + own, own -= m   (RepamLeft / RepamRight)
#pragma unroll
          for (int k = 0; k < LPL; ++k) { m[k] = has_i ? Real(h[k] + oi[k]) : Real(0); oi[k] = has_i ? Real(oi[k] - m[k]) : oi[k]; }
#pragma unroll
          for (int k = 0; k < LPL; ++k) mo[k] = __shfl_xor_sync(FULL, m[k], 16);
#pragma unroll
          for (int k = 0; k < LPL; ++k) {
            const Real m_ir = hi ? mo[k] : m[k], m_ic = hi ? m[k] : mo[k];
            Real v = th[k];
            if (recv_row_first) { if (has_ir) v = v + m_ir; if (has_ic) v = v + m_ic; }
            else { if (has_ic) v = v + m_ic; if (has_ir) v = v + m_ir; }
            th[k] = v;
          }
          if (A.primal) {   // first minimiser of theta (include/factors_messages.hxx:2363,2382-2389); both halves hold theta
            Real lm = th[0];
#pragma unroll
            for (int k = 1; k < LPL; ++k) lm = rmin(lm, th[k]);
            const Real g = half_min<Real>(lm);
            const unsigned il = (unsigned)__ffs((int)(__ballot_sync(FULL, lm == g) & 0xffffu)) - 1u;
            int kf = 0;
#pragma unroll
            for (int k = LPL - 1; k >= 0; --k) if (th[k] == g) kf = k;
            const int k1 = __shfl_sync(FULL, kf, (int)il);
            if (lane == 0) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)s : W - 1 - (uint32_t)s)] = il * LPL + (unsigned)k1;
          }
          // SendMessages with the snapshot t = theta (weights: include/LP_MP.h:1397)
          const double wd = n_out > 0 ? 1.0 / (double)(n_in > n_out ? n_in : n_out) : 0.0;
          const Real wgt = (Real)wd;
#pragma unroll
          for (int k = 0; k < LPL; ++k) {
            const Real d = wgt * th[k];
            Real v = th[k];
            if (has_oc) { if (hi) oo[k] = oo[k] + d; v = v - d; }
            if (has_or) { if (!hi) oo[k] = oo[k] + d; v = v - d; }
            th[k] = v;
          }
          if constexpr (LBF) lb_terms(th, oi, h, has_i);
        }
        sts_n_if(!hi, f_th + roffs, th);
        sts_n_if(true, f_in + roffs, oi);
        sts_n_if(true, f_out + roffs, oo);
        // ---- hand the freshly written column vector on: next warp (ring), next band / stripe (global slots) ----
        if (warp_out_ring && s >= free_upto) {              // warp-uniform, rare: the next warp's ring is full
          unsigned spins = 0;
          do {
            free_upto = (int)(lds_volatile_u32(ctr_next) - step_base) + RING;
          } while (s >= free_upto && !timed_out(spins, 1u << 27));
        }
        sts_n_if(out_ring, ring_out + rslot, oo);
        st_cluster_n_if(out_dsm, ring_dsm + rslot, oo);
        stg_n_relaxed_if(out_glob, out_ptr + (int64_t)s * L, oo);
#pragma unroll
        for (int k = 0; k < LPL; ++k) oth[k] = oo[k];       // lower half: the row neighbour's vector of the next step
        publish(s + 1);
      };

      // ---- steady steps, one tile of 8 columns at a time: every address is a per-tile base + an immediate ----
      // Specialised by where the column vector goes (RINGOUT: the next warp's ring, the 7 upper rows of a band; otherwise ONE generic
      // store into the next CTA's ring / the L2 slots / the neighbouring stripe) and by LAST (the tile ends with the last column):
      // an in-order warp pays ~4.4 cycles per instruction of the step (profiles/r02_step_microbench.txt), so the step carries no
      // predicate it does not need -- theta, the sentinel and both record vectors are stored by both half-warps (same words, same
      // values), the two min-marginals reach both halves in the order of the additions by two indexed shuffles (no selects).
      const int src_first = (lane & 15) | (recv_row_first ? 0 : 16);
      Real* const out_gen = to_next_cta ? reinterpret_cast<Real*>(sp_generic_of_cluster(ring_dsm)) : out_ptr;
      const bool out_gen_on = out_dsm || out_glob;
      auto do_tile = [&](auto last_tag, auto ring_tag, const int j) {
        constexpr bool LAST = decltype(last_tag)::value;     // the tile ends with the last column of the row (no edge to a next column)
        constexpr bool RINGOUT = decltype(ring_tag)::value;  // the row below is the next warp of this CTA
        wait_chunk(j);
        const int s0 = (j << LOGK) - qoff;                  // step (= virtual column) at which the row enters tile j
        const uint32_t slot = (cbase + (uint32_t)j) & (NSLOT - 1);
        const uint32_t toff = slot * SLOT_REC_B + (fwd ? 0u : (uint32_t)(K - 1) * REC_B);
        const uint32_t Tth = f_th + toff, Tin = f_in + toff, Tout = f_out + toff;
        const uint32_t Tpar = f_par + slot * SLOT_PAR_B + (fwd ? 0u : (uint32_t)(K - 1) * PREC_B);
        const uint32_t cq = (uint32_t)(j << LOGK) & (uint32_t)(RING - 1);
        const uint32_t Rin = ring_in + cq * LB, Rout = ring_out + cq * LB;
        Real* const Gout = out_gen + (to_next_cta ? (int64_t)cq * L : (int64_t)s0 * L);
        if (warp_out_ring && s0 + K - 1 >= free_upto) {     // room in the next warp's ring for the whole tile (warp-uniform, rare)
          unsigned spins = 0;
          do {
            free_upto = (int)(lds_volatile_u32(ctr_next) - step_base) + RING;
          } while (s0 + K - 1 >= free_upto && !timed_out(spins, 1u << 27));
        }
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          constexpr int SG = fwd ? 1 : -1;
          const uint32_t ro = (uint32_t)(SG * jj * (int)REC_B), po = (uint32_t)(SG * jj * (int)PREC_B);
          const uint32_t rin = Rin + (uint32_t)jj * LB;
          Real pv[LPL], par[P], th[LPL], oi[LPL], oo[LPL];
          lds_n_volatile<Real, LPL>(pv, rin);               // the polled slot: every lane reads the words of its labels (no vote)
          lds_par<Real, P>(par, Tpar + po);
          lds_n<Real, LPL>(th, Tth + ro);
          lds_n<Real, LPL>(oi, Tin + ro);
          lds_n<Real, LPL>(oo, Tout + ro);
          if (is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) {
            unsigned spins = 0;
            do { lds_n_volatile<Real, LPL>(pv, rin); } while ((is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) && ((++spins & 0xffffu) != 0 || !wait_expired(A.control, t_launch)));
            if (is_sentinel(pv[0]) || is_sentinel(pv[LPL - 1])) {
#pragma unroll
              for (int k = 0; k < LPL; ++k) pv[k] = Real(0);
            }
          }
          {
            Real sent[LPL];
#pragma unroll
            for (int k = 0; k < LPL; ++k) sent[k] = sentinel<Real>();
            sts_n(rin, sent);                               // the slot is free again (both halves: same words)
          }
#pragma unroll
          for (int k = 0; k < LPL; ++k) oth[k] = hi ? pv[k] : oth[k];
          Real h[LPL], m[LPL], tv[LPL];
          split_min_plus<Real, PW, LPL, WIN>(par, oth, h);
#pragma unroll
          for (int k = 0; k < LPL; ++k) { m[k] = h[k] + oi[k]; oi[k] = oi[k] - m[k]; }
#pragma unroll
          for (int k = 0; k < LPL; ++k) {
            const Real m1 = __shfl_sync(FULL, m[k], src_first), m2 = __shfl_sync(FULL, m[k], src_first ^ 16);
            const Real v = (th[k] + m1) + m2;
            const Real d = w_fast * v;
            const Real d_oc = (RINGOUT || out_col_exists) ? d : Real(0), d_or = (LAST && jj == K - 1) ? Real(0) : d;
            if (RINGOUT && !(LAST && jj == K - 1)) oo[k] = oo[k] + d;   // (folded after unrolling)
            else oo[k] = oo[k] + (hi ? d_oc : d_or);
            th[k] = (v - d_oc) - d_or;
            tv[k] = v;
          }
          // hand-off first: the row below is waiting for it
          if constexpr (RINGOUT) sts_n_if(hi, Rout + (uint32_t)jj * LB, oo);
          else st_generic_n_if(out_gen_on, Gout + jj * L, oo);
          sts_n(Tth + ro, th);                              // (both halves hold theta)
          sts_n(Tin + ro, oi);
          sts_n(Tout + ro, oo);
#pragma unroll
          for (int k = 0; k < LPL; ++k) oth[k] = oo[k];
          if constexpr (PRM) keep_label(tv, jj);               // after the stores: nothing of the chain waits for the label
          if constexpr (LBF) lb_terms(th, oi, h, true);       // after the stores: nothing of the chain waits for these minima
        }
        publish(s0 + K);
        if constexpr (PRM) write_tile_labels(s0);
      };

      // Step kinds along the row.  With in-/out-edge counts (n_in, n_out) the weight is 1 / max(n_in, n_out) = 1/2 everywhere except at
      // column 0 of a row without a lower neighbour and at the last column of a row without an upper or a lower neighbour (general
      // step); everything else runs the fast steps, whole tiles of them through do_tile -- INCLUDING the tile of column 0 (its
      // missing row edge is an exact no-op, see above) and the tile of the last column (LAST), so that for W % 8 == 0 a row is
      // tiles only: the skew between consecutive rows is then the hand-off latency of a tile step (~150 ns) instead of that of a
      // single step (~235 ns; scripts/trace_rows.py), and a pass costs W * S + H * skew.
      const bool all_general = (!PRM && A.primal != 0) || (int)W < 4 || (H < 2 && !(A.in_halo || A.out_halo));
      const bool corner_row = (u == 0 && !A.in_halo) || (u + 1 == (int)H && !A.out_halo);
      const bool first_fast = !all_general && out_col_exists, last_fast = !all_general && !corner_row;
      const int s_lo = first_fast ? 0 : 1, s_hi = last_fast ? (int)W - 1 : (int)W - 2;   // fast steps: s_lo .. s_hi
      const int j0 = all_general ? 1 : (s_lo + qoff + K - 1) >> LOGK, j1 = all_general ? 0 : ((s_hi + 1 + qoff) >> LOGK) - 1;   // whole tiles of fast steps
      const bool ring_row = !grid_last_row && w + 1 < NW;   // the row below is the next warp of this CTA
      const bool mark = A.trace_steps && lane == 0 && u < 4096;   // debug: row timeline (scripts/trace_rows.py)
      bool marked = false;
      if (mark) A.trace_steps[4096 + u] = globaltimer_ns();
      int s = 0;
#pragma unroll 1
      while (s < nsteps) {
        const int sp = s + qoff;
        if ((sp & (K - 1)) == 0 && (sp >> LOGK) >= j0 && (sp >> LOGK) <= j1) {
          int j = sp >> LOGK;
          const int j_plain = ((j1 + 1) << LOGK) - qoff == nsteps ? j1 - 1 : j1;   // the tile of the last column is its own instance
          if (ring_row) {
#pragma unroll 1
            for (; j <= j_plain; ++j) do_tile(std::false_type{}, std::true_type{}, j);
            if (j <= j1) do_tile(std::true_type{}, std::true_type{}, j++);
          } else {
#pragma unroll 1
            for (; j <= j_plain; ++j) do_tile(std::false_type{}, std::false_type{}, j);
            if (j <= j1) do_tile(std::true_type{}, std::false_type{}, j++);
          }
          s = (j << LOGK) - qoff;
        } else {
          const int kind = all_general ? 0 : s == 0 ? (first_fast ? 1 : 0) : s + 1 == nsteps ? (last_fast ? 1 : 0) : 2;
          if (kind == 0) do_step(std::integral_constant<int, 0>{}, s);
          else if (kind == 1) do_step(std::integral_constant<int, 1>{}, s);
          else do_step(std::integral_constant<int, 2>{}, s);
          ++s;
        }
        if (mark && !marked) { A.trace_steps[u] = globaltimer_ns(); marked = true; }   // column 0 (or the tile of column 0) handed on
      }
      if constexpr (LBF) { if (lane == 0) A.lb_rows[r] = lb_acc; }
      if (mark) A.trace_steps[3 * 4096 + u] = globaltimer_ns();    // debug: row finished
      }
    }
    // ---- end of the band: every ring is drained and re-armed, every tile has left shared memory ----
#pragma unroll
    for (int w = 0; w < NW; ++w) if (warp_live(kb, w)) chunk_base[w] += (uint32_t)nchunks;
    if (UNI && halo_live(kb)) chunk_base_halo += (uint32_t)nchunks;
    step_base += (uint32_t)nsteps;
    __syncthreads();
  }
  if (wid == NW + 1) tma_wait_all();   // the global writes of the last tiles are complete before the CTA exits
  // no CTA of the cluster leaves while a neighbour may still write into its rings or counters
  if (csize > 1) sp_cluster_sync();
}

}  // namespace lpmp

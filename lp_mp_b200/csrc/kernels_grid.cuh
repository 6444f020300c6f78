// Anti-diagonal wavefront pass over a 4-connected grid MRF: one persistent kernel per pass (sm_100a).
//
// Replaces LP::ComputePass(it...) / ComputePassAndPrimal over forwardUpdateOrdering_ / backwardUpdateOrdering_
// (/root/reference/include/LP_MP.h:981-1005,1592-1602) and the per-node FactorContainer::UpdateFactor
// (include/factors_messages.hxx:2256-2261).  By SURVEY.md App. B every linear extension of the relation DAG
// u_i -> p_e -> u_j gives bit-identical state, so the sequential sweep is re-scheduled as a pipeline.
//
// Storage (DESIGN.md "HBM layout"): one RECORD per node, rec[H][W][5][L] = (theta, aR, aD, bL, bU): the node's unary
// and ITS side of its four edges (aR/aD: first-variable side of the edges to the right / down neighbour, bL/bU:
// second-variable side of the edges to the left / up neighbour), plus a read-only parameter record
// par[H][W][4][P] = parameters of the (R, D, L, U) edges.  A pass reads and writes every record exactly once.
//
//   * a WARP owns a band of T consecutive rows (G = L/4 lanes per node, 4 labels per lane, T = min(32/G, 8)) and
//     sweeps it column by column with the rows skewed by one column: at step s lane-group t updates node
//     (row t of the band, column s - t), i.e. the warp always works on an anti-diagonal fragment;
//   * records move HBM <-> shared memory in TILES of T(+1) rows x K columns with the TMA (cp.async.bulk.tensor 4-D
//     tile loads completing on an mbarrier, tile stores from shared memory), through a ring of NSLOT tiles per
//     warp: one elected lane issues one load and one store per K steps, the other lanes never form a global address;
//     nodes are updated in place in shared memory;
//   * the row neighbour's fresh vector stays in the registers of the same lanes from one step to the next, the
//     column neighbour's vector moves one lane-group up by a warp shuffle -- no memory round trip inside a band;
//   * bands are claimed in order from an atomic ticket; the last row of band k hands its outgoing column vectors to
//     band k+1 through a hand-off buffer in L2 whose slots carry their own validity (a slot holds a NaN sentinel
//     until the producer stores the vector; the consumer polls the slot itself -- two steps ahead -- and puts the
//     sentinel back after reading).  No flag, no fence: one store and one load per lane and column;
//   * interior steps of interior bands run a straight-line FAST path without any existence predicate (missing
//     edges of the first row are exact no-ops because their record slots and parameters are zero); the first /
//     last T steps of a band, the last band and the uniform modes run the GENERAL path.
//
// "Virtual" coordinates (u, q) run in processing order: forward pass u = r, q = c; backward pass u = H-1-r,
// q = W-1-c.  Virtual edges of a node: in_row (towards q-1), in_col (towards u-1), out_row (q+1), out_col (u+1).
// Message order of a node (include/factors_messages.hxx:3402-3419, 2030-2041): first-variable edges, then
// second-variable edges; LIFO inside a type => D, R, L, U (FIFO: R, D, U, L).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "device_common.cuh"

namespace lpmp {

enum { REC_THETA = 0, REC_AR = 1, REC_AD = 2, REC_BL = 3, REC_BU = 4, REC_VECS = 5 };
enum { PAR_R = 0, PAR_D = 1, PAR_L = 2, PAR_U = 3, PAR_EDGES = 4 };

template <typename Real>
struct GridArgs {
  uint32_t H, W;            // rows / columns of this LP (one stripe of a taller grid when striped)
  uint32_t num_bands;
  int forward, primal, fifo, repam;
  int in_halo, out_halo;    // striped: the band before virtual band 0 / after the last band lives on another device
  int tl_reach;             // truncated linear: max over edges of #{d >= 1 : slope*d < trunc} (host-computed)
  unsigned int* control;    // [0] band ticket (zeroed before the launch), [1] sticky abort flag (poll timeout)
  Real* handoff;            // [num_bands][W][L]: boundary b (1 <= b < num_bands) sits between virtual band b-1 and band b;
                            // all-sentinel between passes
  Real* in_boundary;        // [W][L] boundary 0 when in_halo: the neighbouring stripe's outgoing vectors (same protocol)
  Real* out_boundary;       // [W][L] where the last band delivers its outgoing vectors when out_halo (the neighbour's
                            // in_boundary, possibly peer memory), or null: not delivered
  uint32_t* labels;         // [H*W]
};

// ---- hand-off slots: L2-resident, self-validating -------------------------------------------------------------------
// Sentinel = all-ones bit pattern (a NaN no arithmetic of this library produces; cudaMemset(0xff) writes it).
__device__ __forceinline__ bool is_sentinel(float v) { return __float_as_uint(v) == 0xffffffffu; }
__device__ __forceinline__ bool is_sentinel(double v) { return __double_as_longlong(v) == -1ll; }
template <typename Real> __device__ __forceinline__ Real sentinel();
template <> __device__ __forceinline__ float sentinel<float>() { return __uint_as_float(0xffffffffu); }
template <> __device__ __forceinline__ double sentinel<double>() { return __longlong_as_double(-1ll); }

// relaxed system-scope vector accesses (served by L2 / the peer's memory; `volatile` keeps them inside spin loops)
template <int BYTES> __device__ __forceinline__ void ld_relaxed(void* dst, const void* src) {
  static_assert(BYTES == 16 || BYTES == 32, "unsupported hand-off vector width");
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(d[4 * o]), "=r"(d[4 * o + 1]), "=r"(d[4 * o + 2]), "=r"(d[4 * o + 3])
                 : "l"(reinterpret_cast<const char*>(src) + 16 * o));
}
template <int BYTES> __device__ __forceinline__ void st_relaxed(void* dst, const void* src) {
  const uint32_t* d = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<char*>(dst) + 16 * o), "r"(d[4 * o]),
                 "r"(d[4 * o + 1]), "r"(d[4 * o + 2]), "r"(d[4 * o + 3])
                 : "memory");
}

// ---- shared-memory vector access: 4 labels per lane ----------------------------------------------------------------
__device__ __forceinline__ void lds4(float (&v)[4], const float* p) {
  const float4 x = *reinterpret_cast<const float4*>(p);
  v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
}
__device__ __forceinline__ void lds4(double (&v)[4], const double* p) {
  const double2 x = *reinterpret_cast<const double2*>(p), y = *reinterpret_cast<const double2*>(p + 2);
  v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}
__device__ __forceinline__ void sts4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void sts4(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// ---- TMA / mbarrier primitives (inline PTX) -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];" ::"r"(dst),
               "l"(tm), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4}], [%5];" ::"r"(dst),
               "l"(tm), "r"(0), "r"(c1), "r"(c2), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1,%2,%3,%4}], [%5];" ::"l"(tm), "r"(0), "r"(0), "r"(c2), "r"(c3),
               "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int PW> struct ParamCount { static constexpr int value = PW == LPMP_PAIRWISE_POTTS ? 1 : 2; };

// compile-time shape of one kernel instance; the host launcher uses the same struct for tensor maps and smem size
template <typename Real, int PW, int L>
struct WavefrontCfg {
  static constexpr int G = L / 4;                        // lanes per node (4 labels per lane)
  static constexpr int T = (32 / G) < 8 ? (32 / G) : 8;  // rows per band (lanes beyond T*G idle for tiny L)
  static constexpr int P = ParamCount<PW>::value;
  static constexpr int REC = REC_VECS * L;               // reals per node record
  static constexpr int PREC = PAR_EDGES * P;             // reals per parameter record
  static constexpr int NSLOT = 4;                        // tiles in the ring
  static constexpr int PDC = 2;                          // tiles of prefetch distance
  static constexpr size_t slot_bytes(int k) { return (size_t)((T + 1) * k * REC + T * k * PREC) * sizeof(Real); }
  static constexpr int K = slot_bytes(8) * NSLOT <= 200 * 1024 ? 8 : 4;   // columns per tile
  static constexpr int LOGK = K == 8 ? 3 : 2;
  static constexpr int SLOT_REC = (T + 1) * K * REC;     // reals
  static constexpr int SLOT_PAR = T * K * PREC;
  static constexpr int SCRATCH = PW == LPMP_PAIRWISE_TRUNCLIN ? 2 * (32 / G) * 3 * L : 0;   // staging of `other` for the label window
  static constexpr size_t SMEM_BYTES = 128 + (size_t)(NSLOT * (SLOT_REC + SLOT_PAR) + SCRATCH) * sizeof(Real);
  static_assert(L % 4 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "L/4 must be a power of two <= 32");
  static_assert((NSLOT - PDC - 1) * K > T - 2, "a ring slot must be stored before it is reloaded");
  static_assert((SLOT_REC * sizeof(Real)) % 128 == 0 && (SLOT_PAR * sizeof(Real)) % 128 == 0, "TMA destination alignment");
};

// h(x) = min_y [c(x,y) + oth(y)] for the 4 labels x = 4*gl + k of this lane.  oth: registers.
// Potts: one group minimum (two smallest entries when some lambda < 0).  Truncated linear: every y with
// slope*|x-y| >= trunc costs trunc, so those terms collapse to trunc + min(oth); the near terms come from a window of
// 4*HOPS labels either side staged through shared memory (HOPS = 0: all L-1 distances, any parameters).
template <typename Real, int PW, int L, int HOPS>
__device__ __forceinline__ void lane_min_plus(const Real* par, const Real (&oth)[4], Real (&h)[4], Real* scratch, int reach, int gl) {
  constexpr int G = L / 4;
  constexpr unsigned FULL = 0xffffffffu;
  Real g = rmin(rmin(oth[0], oth[1]), rmin(oth[2], oth[3]));
  if constexpr (PW == LPMP_PAIRWISE_POTTS) {
    const Real lambda = par[0];
    if (__any_sync(FULL, lambda < Real(0))) {   // repulsive Potts somewhere in the warp: the two smallest entries
      const Real INF = RealTraits<Real>::inf();
      TwoMin<Real> t{INF, INF, 0u};
#pragma unroll
      for (int k = 0; k < 4; ++k) two_min_insert(t, oth[k], (uint32_t)(gl * 4 + k));
      group_two_min(t, G, FULL);
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = rmin(oth[k], Real(lambda + ((uint32_t)(gl * 4 + k) == t.i1 ? t.m2 : t.m1)));
    } else {
      // lambda >= 0: at the minimiser min(m1, fl(lambda + m2)) = m1 = min(m1, fl(lambda + m1)) because rounding is
      // monotone, so one group minimum gives bit-identical results to the two_min formula
      g = group_min(g, G, FULL);
      const Real far = lambda + g;
#pragma unroll
      for (int k = 0; k < 4; ++k) h[k] = rmin(oth[k], far);
    }
  } else {
    const Real slope = par[0], trunc = par[1];
    Real* sc = scratch + L + 4 * gl;             // this lane's labels inside the group's [INF pad | L labels | INF pad] staging
    sts4(sc, oth);
    g = group_min(g, G, FULL);                   // the shuffles also separate the staging stores from the loads below
    __syncwarp();
    const Real far = trunc + g;
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = rmin(oth[k], far);
    if constexpr (HOPS > 0) {
      Real win[4 * (2 * HOPS + 1)];              // labels 4*(gl-HOPS) .. 4*(gl+HOPS)+3
#pragma unroll
      for (int hh = 0; hh < 2 * HOPS + 1; ++hh) {
        Real v[4];
        lds4(v, sc + 4 * (hh - HOPS));
#pragma unroll
        for (int k = 0; k < 4; ++k) win[4 * hh + k] = v[k];
      }
#pragma unroll
      for (int d = 1; d <= 4 * HOPS; ++d) {
        if (d > reach) break;
        const Real c = rmin(Real(slope * Real(d)), trunc);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = rmin(h[k], rmin(Real(c + win[4 * HOPS + k - d]), Real(c + win[4 * HOPS + k + d])));
      }
    } else {
      for (int d = 1; d < L; ++d) {
        const Real c = rmin(Real(slope * Real(d)), trunc);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = rmin(h[k], rmin(Real(c + sc[k - d]), Real(c + sc[k + d])));
      }
    }
  }
}

template <typename Real, int PW, int L, int HOPS, bool ANISO>
__global__ void __launch_bounds__(32) wavefront_pass_kernel(const __grid_constant__ CUtensorMap tm_rec, const __grid_constant__ CUtensorMap tm_rec1,
                                                            const __grid_constant__ CUtensorMap tm_par, const GridArgs<Real> A) {
  using Cfg = WavefrontCfg<Real, PW, L>;
  constexpr int G = Cfg::G, T = Cfg::T, P = Cfg::P, K = Cfg::K, LOGK = Cfg::LOGK, REC = Cfg::REC, PREC = Cfg::PREC, NSLOT = Cfg::NSLOT, PDC = Cfg::PDC;
  constexpr int BOXROWS = ANISO ? T : T + 1;           // the uniform modes also read the first row of the next band
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  constexpr uint32_t TILE_BYTES = (uint32_t)((BOXROWS * K * REC + T * K * PREC) * sizeof(Real));
  constexpr bool PARTIAL_WARP = T * G < 32;            // tiny L: lanes beyond T groups shadow group T-1 and never store
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  Real* rec_ring = reinterpret_cast<Real*>(smem_raw + 128);
  Real* par_ring = rec_ring + (size_t)NSLOT * Cfg::SLOT_REC;
  Real* scratch_all = par_ring + (size_t)NSLOT * Cfg::SLOT_PAR;
  const int lane = threadIdx.x & 31;
  const bool lane_used = !PARTIAL_WARP || lane / G < T;
  const int t = lane_used ? lane / G : T - 1, gl = lane % G;
  const bool fwd = A.forward != 0;
  const uint32_t H = A.H, W = A.W;
  const int nchunks = ((int)W + K - 1) >> LOGK;
  const int nb = (int)A.num_bands;
  // TMA tile stores must not start at negative coordinates, so tiles are anchored at row 0 / column 0 in BOTH
  // directions: backward, the partially filled tile column / band is the first one processed.  qoff / roff = number
  // of padding columns / rows in front of virtual column / row 0; q' = q + qoff and u' = u + roff index the tiles.
  const int qoff = fwd ? 0 : nchunks * K - (int)W;
  const int roff = fwd ? 0 : nb * T - (int)H;
  const int sp_end = qoff + (int)W + T - 1;
  const CUtensorMap* const tm_load = ANISO ? &tm_rec : &tm_rec1;
  Real* const scratch0 = scratch_all + (size_t)(2 * (lane / G)) * 3 * L;   // two staging areas per lane group
  Real* const scratch1 = scratch0 + 3 * L;

  if (lane == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    for (int i = lane; i < Cfg::SCRATCH; i += 32) scratch_all[i] = RealTraits<Real>::inf();
  }
  __syncwarp();
  uint32_t chunk_base = 0;   // tiles loaded by earlier bands of this warp: slot / mbarrier-phase sequence continues

  // record vector / parameter roles in this direction
  const int v_in_row = fwd ? REC_BL : REC_AR, v_in_col = fwd ? REC_BU : REC_AD;
  const int v_out_row = fwd ? REC_AR : REC_BL, v_out_col = fwd ? REC_AD : REC_BU;
  const int p_in_row = fwd ? PAR_L : PAR_R, p_in_col = fwd ? PAR_U : PAR_D;
  const int p_out_row = fwd ? PAR_R : PAR_L, p_out_col = fwd ? PAR_D : PAR_U;
  const int kxor = fwd ? 0 : K - 1;
  const int trow = fwd ? t : BOXROWS - 1 - t;            // row of this lane group inside the record tile
  const int trow_par = fwd ? t : T - 1 - t;
  const int below = fwd ? K * REC : -K * REC;            // smem distance to the same column of virtual row u+1
  // receive order of the two processed-side edges: LIFO: (L, U) forward, (D, R) backward
  const bool recv_row_first = fwd ? !A.fifo : (A.fifo != 0);
  const int reach = A.tl_reach;

  for (;;) {
    unsigned kb = 0;
    if (lane == 0) kb = atomicAdd(A.control, 1u);
    kb = __shfl_sync(FULL, kb, 0);
    if (kb >= A.num_bands) break;
    const int u = (int)kb * T + t - roff;                  // virtual row of this lane group
    const bool row_active = u >= 0 && u < (int)H && lane_used;
    const uint32_t r = fwd ? (uint32_t)u : H - 1 - (uint32_t)u;
    const bool in_col_exists = u > 0 || A.in_halo, out_col_exists = u + 1 < (int)H || A.out_halo;
    const bool last_band = kb + 1 == A.num_bands;
    const bool poll_in = kb > 0 || A.in_halo;              // group t_in takes its column neighbour from the hand-off buffer
    const int t_in = kb == 0 ? roff : 0;                   // first active group of the band
    const bool hand_out = !last_band || (A.out_halo && A.out_boundary != nullptr);
    const Real* const hand_in = (kb == 0 ? A.in_boundary : A.handoff + (uint64_t)kb * W * L) + gl * 4;
    Real* const hand_dst = (last_band ? A.out_boundary : A.handoff + (uint64_t)(kb + 1) * W * L) + gl * 4;
    const int t_hand = last_band ? (int)H - 1 - (int)kb * T + roff : T - 1;   // group that owns the band's last row
    // tile origin rows (actual row of tile row 0); columns follow from the chunk index
    const int row0_own = fwd ? (int)kb * T : (nb - 1 - (int)kb) * T;
    const int row0_rec = fwd ? row0_own : row0_own - (BOXROWS - T);
    const int store_skip = fwd ? 0 : (BOXROWS - T) * K * REC;   // backward the extra (read-only) tile row comes first

    auto issue_load = [&](int j) {   // lane 0 only
      const int slot = (int)((chunk_base + (uint32_t)j) & (NSLOT - 1));
      const int col0 = fwd ? j * K : (nchunks - 1 - j) * K;
      const uint32_t bar = smem_u32(bars + slot);
      mbar_expect_tx(bar, TILE_BYTES);
      tma_load_4d(smem_u32(rec_ring + (size_t)slot * Cfg::SLOT_REC), tm_load, col0, row0_rec, bar);
      tma_load_3d(smem_u32(par_ring + (size_t)slot * Cfg::SLOT_PAR), &tm_par, col0, row0_own, bar);
    };
    auto wait_chunk = [&](int j) {   // all lanes
      const uint32_t n = chunk_base + (uint32_t)j;
      const uint32_t bar = smem_u32(bars + (n & (NSLOT - 1)));
      const uint32_t parity = (n / NSLOT) & 1u;
      while (!mbar_try_wait(bar, parity)) {}
    };
    auto store_chunk = [&](int j) {  // all lanes: the tile is complete in shared memory
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int slot = (int)((chunk_base + (uint32_t)j) & (NSLOT - 1));
        const int col0 = fwd ? j * K : (nchunks - 1 - j) * K;
        tma_store_4d(&tm_rec, col0, row0_own, smem_u32(rec_ring + (size_t)slot * Cfg::SLOT_REC + store_skip));
        tma_commit();
      }
    };

    if (lane == 0) {
      for (int j = 0; j <= PDC && j < nchunks; ++j) issue_load(j);
    }

    // hand-off polls are issued POLL_AHEAD steps before use so that their L2 round trip is off the dependency chain
    constexpr int POLL_AHEAD = 2;
    Real poll[POLL_AHEAD][4];
#pragma unroll
    for (int d = 0; d < POLL_AHEAD; ++d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) poll[d][k] = Real(0);
    }
    if (poll_in && t == t_in) {   // column q is consumed at step q + t_in, which uses poll[(q + t_in) & 1]
      if ((t_in & 1) == 0) { ld_relaxed<HAND_BYTES>(poll[0], hand_in); if (W > 1) ld_relaxed<HAND_BYTES>(poll[1], hand_in + L); }
      else { ld_relaxed<HAND_BYTES>(poll[1], hand_in); if (W > 1) ld_relaxed<HAND_BYTES>(poll[0], hand_in + L); }
    }

    Real carry_row[4], carry_col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { carry_row[k] = Real(0); carry_col[k] = Real(0); }
    // weight of the FAST path: (n_in, n_out) = (2, 2), first row (1, 2): 1 / max(n_in, n_out) = 1/2 (include/LP_MP.h:1397)
    const Real w_fast = (Real)(1.0 / 2.0);
    int next_store = 0;

    wait_chunk(0);
    if (!ANISO && nchunks > 1) wait_chunk(1);   // uniform modes: the out_row neighbour may sit in the next tile

    // sp = step index shifted by qoff; group t is at tile column q' = sp - t, virtual column q = q' - qoff
    auto do_step = [&](auto fast_tag, const int sp, Real (&polled)[4]) {
      constexpr bool FAST = decltype(fast_tag)::value;
      // ---- tile ring management (warp-uniform) ----
      if ((sp & (K - 1)) == 0 && sp > 0) {
        const int j = sp >> LOGK;
        const int jw = ANISO ? j : j + 1;
        if (jw < nchunks) wait_chunk(jw);
        if (lane == 0 && j + PDC < nchunks) {
          tma_wait_read<(T < 2 ? 1 : 0)>();     // the tile that last lived in this slot has left shared memory
          issue_load(j + PDC);
        }
        __syncwarp();
      }
      const int qp = sp - t;
      const int q = qp - qoff;
      const int slot = (int)((chunk_base + (uint32_t)(qp >> LOGK)) & (NSLOT - 1));
      const int kk = (qp & (K - 1)) ^ kxor;
      Real* const rec = rec_ring + (size_t)slot * Cfg::SLOT_REC + (size_t)(trow * K + kk) * REC + gl * 4;
      const Real* const par = par_ring + (size_t)slot * Cfg::SLOT_PAR + (size_t)(trow_par * K + kk) * PREC;

      // ---- the column neighbour's vector: shuffle inside the band, hand-off slot for group 0 ----
      Real col_in[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) col_in[k] = carry_col[k];
      if (poll_in && t == t_in && q >= 0 && q < (int)W) {
        const Real* hslot = hand_in + (uint64_t)q * L;
        unsigned spins = 0;
        while (is_sentinel(polled[0]) || is_sentinel(polled[1]) || is_sentinel(polled[2]) || is_sentinel(polled[3])) {
          ld_relaxed<HAND_BYTES>(polled, hslot);
          if ((++spins & 0x3ffu) == 0) {      // bounded wait: a lost producer must not hang the device
            if (spins > (1u << 22)) atomicExch(A.control + 1, 1u);
            if (*reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u) {
#pragma unroll
              for (int k = 0; k < 4; ++k) polled[k] = Real(0);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) col_in[k] = polled[k];
        Real sent[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) sent[k] = sentinel<Real>();
        st_relaxed<HAND_BYTES>(const_cast<Real*>(hslot), sent);        // the slot is free again for the next pass
        if (q + POLL_AHEAD < (int)W) ld_relaxed<HAND_BYTES>(polled, hslot + (uint64_t)POLL_AHEAD * L);
      }

      Real th[4], own_ir[4], own_ic[4], own_or[4], own_oc[4];
      lds4(th, rec + REC_THETA * L);
      lds4(own_ir, rec + v_in_row * L);
      lds4(own_ic, rec + v_in_col * L);
      lds4(own_or, rec + v_out_row * L);
      lds4(own_oc, rec + v_out_col * L);
      Real par_ir[P], par_ic[P];
#pragma unroll
      for (int p = 0; p < P; ++p) { par_ir[p] = par[p_in_row * P + p]; par_ic[p] = par[p_in_col * P + p]; }

      Real out_col[4];
      if constexpr (FAST) {
        // ---- ReceiveMessages from the two processed neighbours, SendMessages to the two unprocessed ones ----
        Real h_ir[4], h_ic[4];
        lane_min_plus<Real, PW, L, HOPS>(par_ir, carry_row, h_ir, scratch0, reach, gl);
        lane_min_plus<Real, PW, L, HOPS>(par_ic, col_in, h_ic, scratch1, reach, gl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real m_ir = h_ir[k] + own_ir[k], m_ic = h_ic[k] + own_ic[k];
          own_ir[k] = own_ir[k] - m_ir;
          own_ic[k] = own_ic[k] - m_ic;
          th[k] = recv_row_first ? (th[k] + m_ir) + m_ic : (th[k] + m_ic) + m_ir;
        }
        if (A.primal) {   // MaximizePotentialAndComputePrimal: first minimum
          Real bv = th[0];
          uint32_t bi = (uint32_t)(gl * 4);
#pragma unroll
          for (int k = 1; k < 4; ++k) if (th[k] < bv) { bv = th[k]; bi = (uint32_t)(gl * 4 + k); }
          group_argmin(bv, bi, G, FULL);
          if (gl == 0 && lane_used) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)q : W - 1 - (uint32_t)q)] = bi;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real d = w_fast * th[k];   // every send uses the same snapshot value
          own_oc[k] = own_oc[k] + d;
          own_or[k] = own_or[k] + d;
          th[k] = (th[k] - d) - d;
          carry_row[k] = own_or[k];
          out_col[k] = own_oc[k];
        }
        if (lane_used) {
          sts4(rec + REC_THETA * L, th);
          sts4(rec + v_in_row * L, own_ir);
          sts4(rec + v_in_col * L, own_ic);
          sts4(rec + v_out_row * L, own_or);
          sts4(rec + v_out_col * L, own_oc);
        }
      } else {
        const bool act = row_active && q >= 0 && q < (int)W;
        const bool has_ir = act && q > 0, has_or = act && q + 1 < (int)W;
        const bool has_ic = act && in_col_exists, has_oc = act && out_col_exists;
        const int n_in = (int)has_ir + (int)has_ic, n_out = (int)has_or + (int)has_oc;
        Real oth_or[4], oth_oc[4], par_or[P], par_oc[P];
        if constexpr (!ANISO) {
          // the other side of the edges towards unprocessed neighbours: their in-side vectors, still in the state of
          // the previous pass, in the neighbouring records of the tile ring
          const int q1 = qp + 1;
          const int slot1 = (int)((chunk_base + (uint32_t)(q1 >> LOGK)) & (NSLOT - 1));
          const Real* rec_q1 = rec_ring + (size_t)slot1 * Cfg::SLOT_REC + (size_t)(trow * K + ((q1 & (K - 1)) ^ kxor)) * REC + gl * 4;
          lds4(oth_or, rec_q1 + v_in_row * L);
          lds4(oth_oc, rec + below + v_in_col * L);
#pragma unroll
          for (int p = 0; p < P; ++p) { par_or[p] = par[p_out_row * P + p]; par_oc[p] = par[p_out_col * P + p]; }
        }
        // min-marginal of one edge (independent of theta): m = h + own, own -= m   (RepamLeft / RepamRight)
        auto marginal = [&](bool on, const Real (&pr)[P], const Real (&oth)[4], Real (&own)[4], Real (&m)[4], Real* sc) {
          Real pz[P], oz[4], h[4];
#pragma unroll
          for (int p = 0; p < P; ++p) pz[p] = on ? pr[p] : Real(0);
#pragma unroll
          for (int k = 0; k < 4; ++k) oz[k] = on ? oth[k] : Real(0);
          lane_min_plus<Real, PW, L, HOPS>(pz, oz, h, sc, reach, gl);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            m[k] = on ? Real(h[k] + own[k]) : Real(0);
            own[k] = on ? Real(own[k] - m[k]) : own[k];
          }
        };
        Real m_ir[4], m_ic[4];
        marginal(has_ir, par_ir, carry_row, own_ir, m_ir, scratch0);
        marginal(has_ic, par_ic, col_in, own_ic, m_ic, scratch1);
        if constexpr (ANISO) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            Real x = th[k];
            if (recv_row_first) { if (has_ir) x = x + m_ir[k]; if (has_ic) x = x + m_ic[k]; }
            else { if (has_ic) x = x + m_ic[k]; if (has_ir) x = x + m_ir[k]; }
            th[k] = x;
          }
        } else {
          Real m_or[4], m_oc[4];
          marginal(has_or, par_or, oth_or, own_or, m_or, scratch0);
          marginal(has_oc, par_oc, oth_oc, own_oc, m_oc, scratch1);
          // actual edges: forward (in_row, in_col, out_row, out_col) = (L, U, R, D); backward = (R, D, L, U)
          const Real* mR = fwd ? m_or : m_ir; const Real* mD = fwd ? m_oc : m_ic;
          const Real* mL = fwd ? m_ir : m_or; const Real* mU = fwd ? m_ic : m_oc;
          const bool hR = fwd ? has_or : has_ir, hD = fwd ? has_oc : has_ic, hL = fwd ? has_ir : has_or, hU = fwd ? has_ic : has_oc;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            Real x = th[k];
            if (!A.fifo) { if (hD) x = x + mD[k]; if (hR) x = x + mR[k]; if (hL) x = x + mL[k]; if (hU) x = x + mU[k]; }
            else { if (hR) x = x + mR[k]; if (hD) x = x + mD[k]; if (hU) x = x + mU[k]; if (hL) x = x + mL[k]; }
            th[k] = x;
          }
        }
        if (A.primal) {
          Real bv = th[0];
          uint32_t bi = (uint32_t)(gl * 4);
#pragma unroll
          for (int k = 1; k < 4; ++k) if (th[k] < bv) { bv = th[k]; bi = (uint32_t)(gl * 4 + k); }
          group_argmin(bv, bi, G, FULL);
          if (act && gl == 0) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)q : W - 1 - (uint32_t)q)] = bi;
        }
        // SendMessages with the snapshot t = theta (weights: include/LP_MP.h:1397,1437)
        const int deg = n_in + n_out;
        double wd;
        if (ANISO) wd = n_out > 0 ? 1.0 / (double)(n_in > n_out ? n_in : n_out) : 0.0;
        else if (A.repam == LPMP_REPAM_UNIFORM) wd = deg > 0 ? 1.0 / (double)deg : 0.0;
        else wd = 1.0 / (double)(deg + 1);
        const Real w = (Real)wd;
        const bool send_ir = !ANISO && has_ir, send_ic = !ANISO && has_ic;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real d = w * th[k];
          Real x = th[k];
          if (has_oc) { own_oc[k] = own_oc[k] + d; x = x - d; }
          if (has_or) { own_or[k] = own_or[k] + d; x = x - d; }
          if (send_ir) { own_ir[k] = own_ir[k] + d; x = x - d; }
          if (send_ic) { own_ic[k] = own_ic[k] + d; x = x - d; }
          th[k] = x;
          carry_row[k] = own_or[k];
          out_col[k] = own_oc[k];
        }
        if (act) {
          sts4(rec + REC_THETA * L, th);
          sts4(rec + v_in_row * L, own_ir);
          sts4(rec + v_in_col * L, own_ic);
          sts4(rec + v_out_row * L, own_or);
          sts4(rec + v_out_col * L, own_oc);
        }
      }

      // ---- hand the freshly written column vector to the node below: next lane group, or the next band ----
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Real up = __shfl_up_sync(FULL, out_col[k], G);
        carry_col[k] = lane < G ? Real(0) : up;
      }
      if (hand_out && t == t_hand && lane_used && q >= 0 && q < (int)W) st_relaxed<HAND_BYTES>(hand_dst + (uint64_t)q * L, out_col);

      // ---- a tile leaves once its last row has passed its last column ----
      if (((sp - T + 2) & (K - 1)) == 0 && sp >= T + K - 2) {
        store_chunk(next_store);
        ++next_store;
      }
    };

    // FAST path: full bands away from the last grid row, all groups strictly inside the row (1 <= q <= W-2)
    const bool fast_band = ANISO && !last_band && !(kb == 0 && roff > 0) && W >= 3;
    for (int sp = qoff; sp < sp_end; sp += POLL_AHEAD) {
      const int s = sp - qoff;
      if (fast_band && s >= T && s + 1 <= (int)W - 2) {
        do_step(std::true_type{}, sp, poll[0]);
        do_step(std::true_type{}, sp + 1, poll[1]);
      } else {
        do_step(std::false_type{}, sp, poll[0]);
        if (sp + 1 < sp_end) do_step(std::false_type{}, sp + 1, poll[1]);
      }
    }
    while (next_store < nchunks) { store_chunk(next_store); ++next_store; }
    if (lane == 0) tma_wait_read<0>();   // every tile of this band has left shared memory before the ring is reused
    __syncwarp();
    chunk_base += (uint32_t)nchunks;
  }
  if (lane == 0) tma_wait_all();   // the global writes of the last tiles are complete before the warp exits
}

}  // namespace lpmp

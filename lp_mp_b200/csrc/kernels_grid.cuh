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
//     until the producer stores the vector; the consumer polls the slot itself -- PD steps ahead, PD loads in flight -- and puts the
//     sentinel back after reading).  No flag, no fence: one store and one load per lane and column;
//   * a CTA holds NW warps (1..8, as many as the shared memory allows) that work on NW CONSECUTIVE bands (a "super-band",
//     claimed together from the ticket): inside the CTA the hand-off goes through a small ring of self-validating slots
//     in SHARED memory (same sentinel protocol, ~30 cycles instead of an L2 round trip, plus back-pressure: the producer
//     only writes a slot that holds the sentinel), so only every NW-th band boundary pays the L2 latency;
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
  int lean_flags;           // lean variant, experiments: 2 = rotate the warp roles per SM arrival, 4 = idle movers do NOT sleep
  int potts_nonneg;         // Potts: the host has seen no negative lambda (the kernels then skip the two-smallest-entries test)
  int tl_reach;             // truncated linear: max over edges of #{d >= 1 : slope*d < trunc} (host-computed)
  unsigned int* control;    // [0] band ticket (zeroed before the launch), [1] sticky abort flag (poll timeout),
                            // [2 + smid] CTAs that have arrived on SM smid (lean variant: rotation of the warp roles)
  Real* handoff;            // [num_bands][W][L]: boundary b (1 <= b < num_bands) sits between virtual band b-1 and band b;
                            // all-sentinel between passes
  Real* in_boundary;        // [W][L] boundary 0 when in_halo: the neighbouring stripe's outgoing vectors (same protocol)
  Real* out_boundary;       // [W][L] where the last band delivers its outgoing vectors when out_halo (the neighbour's
                            // in_boundary, possibly peer memory), or null: not delivered
  Real* later_boundary;     // [W][L] uniform modes on a stripe with out_halo: the in-side vectors of the NOT yet processed
                            // neighbouring stripe's first virtual row (state of the previous pass, delivered before the pass;
                            // same slot protocol: the last virtual row of the last band consumes them), or null
  uint32_t* labels;         // [H*W]
  double* lb_rows;          // [H] or null; split variant, backward pass: LP::LowerBound terms of every grid row, accumulated on the fly
  Real* rec;                // node records [H][W][5][L] (the split variant moves whole-row tile runs with 1-D bulk copies, no tensor map)
  Real* par;                // parameter records [H][W][4][P]
  unsigned long long* trace;   // debug: per band {claimed, first tile ready, first hand-off consumed, done} in ns (globaltimer), or null
  unsigned long long* trace_steps;   // debug (lean variant): [2][8192] end-of-step stamps of bands trace_band and trace_band + 1, or null
  uint32_t trace_band;
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
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int PW> struct ParamCount { static constexpr int value = PW == LPMP_PAIRWISE_POTTS ? 1 : 2; };

// compile-time shape of one kernel instance; the host launcher uses the same struct for tensor maps and smem size.
// KSEL = columns per tile: 8 (fewer TMA operations; latency-bound grids) or 4 (half the shared memory per band, twice
// the resident bands per SM; bandwidth-bound grids).
template <typename Real, int PW, int L, int KSEL>
struct WavefrontCfg {
  static constexpr int G = L / 4;                        // lanes per node (4 labels per lane)
  static constexpr int TG = 32 / G;                      // lane groups per warp
  static constexpr int T = TG < 8 ? TG : 8;              // rows per band (lanes beyond T*G idle for tiny L)
  static constexpr int P = ParamCount<PW>::value;
  static constexpr int REC = REC_VECS * L;               // reals per node record
  static constexpr int PREC = PAR_EDGES * P;             // reals per parameter record
  static constexpr int NSLOT = 4;                        // tiles in the ring
  static constexpr int PDC = 2;                          // tiles of prefetch distance
  static constexpr int K = KSEL;
  static constexpr int PD = KSEL == 4 ? 4 : 2;           // hand-off slots polled ahead (L2 round trips in flight per band):
                                                         // narrow tiles = bandwidth-bound grids, where the loaded L2 latency is several steps
  static constexpr int LOGK = K == 8 ? 3 : 2;
  static constexpr int SLOT_REC = (T + 1) * K * REC;     // reals
  static constexpr int SLOT_PAR = T * K * PREC;
  static constexpr int SCRATCH = PW == LPMP_PAIRWISE_TRUNCLIN ? 2 * TG * 3 * L : 0;   // staging of `other` for the label window
  static constexpr size_t SMEM_BYTES = 128 + (size_t)(NSLOT * (SLOT_REC + SLOT_PAR) + SCRATCH) * sizeof(Real);   // per warp
  static constexpr size_t WARP_SMEM = (SMEM_BYTES + 127) / 128 * 128;
  static constexpr int RING = 16;                        // columns of an intra-CTA hand-off ring (power of two, > PD)
  static constexpr int MAX_WARPS = 8;
  // dynamic shared memory of a CTA of nw warps: nw private regions, the super-band queue (128 B), nw-1 hand-off rings
  static constexpr size_t cta_smem(int nw) { return (size_t)nw * WARP_SMEM + 128 + (size_t)(nw - 1) * RING * L * sizeof(Real); }
  static constexpr bool FITS = SMEM_BYTES <= 220 * 1024 && (NSLOT - PDC - 1) * K > T - 2;
  static_assert(K == 8 || K == 4, "tile width");
  static_assert(L % 4 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "L/4 must be a power of two <= 32");
  static_assert((SLOT_REC * sizeof(Real)) % 128 == 0 && (SLOT_PAR * sizeof(Real)) % 128 == 0, "TMA destination alignment");
};

// predicated variants (no branch in the instruction stream; a skipped load leaves dst unchanged)
template <int BYTES> __device__ __forceinline__ void ld_relaxed_if(bool on, void* dst, const void* src) {
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4]; }"
                 : "+r"(d[4 * o]), "+r"(d[4 * o + 1]), "+r"(d[4 * o + 2]), "+r"(d[4 * o + 3])
                 : "l"(reinterpret_cast<const char*>(src) + 16 * o), "r"((uint32_t)on));
}
template <int BYTES> __device__ __forceinline__ void st_relaxed_if(bool on, void* dst, const void* src) {
  const uint32_t* d = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4}; }" ::"l"(reinterpret_cast<char*>(dst) + 16 * o),
                 "r"(d[4 * o]), "r"(d[4 * o + 1]), "r"(d[4 * o + 2]), "r"(d[4 * o + 3]), "r"((uint32_t)on)
                 : "memory");
}
// hand-off slots in shared memory (rings between the warps of a CTA): same protocol, 32-bit shared addresses
template <int BYTES> __device__ __forceinline__ void lds_hand(void* dst, uint32_t a) {
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d[4 * o]), "=r"(d[4 * o + 1]), "=r"(d[4 * o + 2]), "=r"(d[4 * o + 3]) : "r"(a + 16 * o));
}
template <int BYTES> __device__ __forceinline__ void lds_hand_if(bool on, void* dst, uint32_t a) {
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4]; }"
                 : "+r"(d[4 * o]), "+r"(d[4 * o + 1]), "+r"(d[4 * o + 2]), "+r"(d[4 * o + 3])
                 : "r"(a + 16 * o), "r"((uint32_t)on));
}
template <int BYTES> __device__ __forceinline__ void sts_hand_if(bool on, uint32_t a, const void* src) {
  const uint32_t* d = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4}; }" ::"r"(a + 16 * o), "r"(d[4 * o]),
                 "r"(d[4 * o + 1]), "r"(d[4 * o + 2]), "r"(d[4 * o + 3]), "r"((uint32_t)on)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_volatile_u32(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// pins a kernel parameter in a register (the compiler otherwise re-reads the constant bank inside the step loop)
__device__ __forceinline__ uint32_t pin(uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }

// ---- shared-memory vector access by 32-bit shared address: 4 labels per lane ----------------------------------------
// volatile: program order among these accesses is kept (tiles are also written by the TMA and re-read in place)
__device__ __forceinline__ void lds4(float (&v)[4], uint32_t a) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
}
__device__ __forceinline__ void lds4(double (&v)[4], uint32_t a) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(v[2]), "=d"(v[3]) : "r"(a));
}
__device__ __forceinline__ void sts4(uint32_t a, const float (&v)[4]) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void sts4(uint32_t a, const double (&v)[4]) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v[0]), "d"(v[1]) : "memory");
  asm volatile("st.shared.v2.f64 [%0+16], {%1,%2};" ::"r"(a), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ float lds1f(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds1d(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <typename Real> __device__ __forceinline__ Real lds1(uint32_t a);
template <> __device__ __forceinline__ float lds1<float>(uint32_t a) { return lds1f(a); }
template <> __device__ __forceinline__ double lds1<double>(uint32_t a) { return lds1d(a); }
template <typename Real, int P> __device__ __forceinline__ void lds_par(Real (&v)[P], uint32_t a) {
#pragma unroll
  for (int p = 0; p < P; ++p) v[p] = lds1<Real>(a + p * (uint32_t)sizeof(Real));
}

// two interleaved butterfly minima over the G lanes of a group (the shuffles of the two chains overlap)
template <typename Real, int G>
__device__ __forceinline__ void group_min2(Real& a, Real& b) {
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) {
    const Real ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
    a = rmin(a, ta);
    b = rmin(b, tb);
  }
}

// Min-plus of TWO edges at once (independent chains, interleaved so that their latencies overlap):
//   h(x) = min_y [c(x,y) + oth(y)] for the 4 labels x = 4*gl + k of this lane, oth in registers.
// Potts: one group minimum (two smallest entries when some lambda < 0).  Truncated linear: every y with
// slope*|x-y| >= trunc costs trunc, so those terms collapse to trunc + min(oth); the near terms come from a window of
// 4*HOPS labels either side, staged through shared memory (sa / sb: shared address of this lane's 4 labels inside the
// group's [INF pad | L labels | INF pad] staging areas).  HOPS = 0: all L-1 distances, any parameters.
template <typename Real, int PW, int L, int HOPS>
__device__ __forceinline__ void lane_min_plus2(const Real* pa, const Real (&oa)[4], Real (&ha)[4], const Real* pb, const Real (&ob)[4], Real (&hb)[4],
                                               uint32_t sa, uint32_t sb, int gl, bool potts_nonneg = false) {
  constexpr int G = L / 4;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  Real ga = rmin(rmin(oa[0], oa[1]), rmin(oa[2], oa[3]));
  Real gb = rmin(rmin(ob[0], ob[1]), rmin(ob[2], ob[3]));
  if constexpr (PW == LPMP_PAIRWISE_POTTS) {
    const Real la = pa[0], lb = pb[0];
    // (Potts instances use HOPS as a flag: 1 = the host has seen no negative lambda, the two-smallest-entries branch is compiled out)
    if (HOPS == 0 && !potts_nonneg && __any_sync(FULL, la < Real(0) || lb < Real(0))) {   // repulsive Potts somewhere in the warp: the two smallest entries
      const Real INF = RealTraits<Real>::inf();
      TwoMin<Real> ta{INF, INF, 0u}, tb{INF, INF, 0u};
#pragma unroll
      for (int k = 0; k < 4; ++k) { two_min_insert(ta, oa[k], (uint32_t)(gl * 4 + k)); two_min_insert(tb, ob[k], (uint32_t)(gl * 4 + k)); }
      group_two_min(ta, G, FULL);
      group_two_min(tb, G, FULL);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ha[k] = rmin(oa[k], Real(la + ((uint32_t)(gl * 4 + k) == ta.i1 ? ta.m2 : ta.m1)));
        hb[k] = rmin(ob[k], Real(lb + ((uint32_t)(gl * 4 + k) == tb.i1 ? tb.m2 : tb.m1)));
      }
    } else {
      // lambda >= 0: at the minimiser min(m1, fl(lambda + m2)) = m1 = min(m1, fl(lambda + m1)) because rounding is
      // monotone, so one group minimum gives bit-identical results to the two_min formula
      group_min2<Real, G>(ga, gb);
      const Real fa = la + ga, fb = lb + gb;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ha[k] = rmin(oa[k], fa); hb[k] = rmin(ob[k], fb); }
    }
  } else {
    const Real slope_a = pa[0], trunc_a = pa[1], slope_b = pb[0], trunc_b = pb[1];
    sts4(sa, oa);
    sts4(sb, ob);
    group_min2<Real, G>(ga, gb);
    __syncwarp();
    if constexpr (HOPS > 0) {
      Real wa[4 * (2 * HOPS + 1)], wb[4 * (2 * HOPS + 1)];   // labels 4*(gl-HOPS) .. 4*(gl+HOPS)+3
#pragma unroll
      for (int hh = 0; hh < 2 * HOPS + 1; ++hh) {
        if (hh == HOPS) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { wa[4 * hh + k] = oa[k]; wb[4 * hh + k] = ob[k]; }
        } else {
          Real va[4], vb[4];
          lds4(va, sa + (uint32_t)(4 * (hh - HOPS)) * S);
          lds4(vb, sb + (uint32_t)(4 * (hh - HOPS)) * S);
#pragma unroll
          for (int k = 0; k < 4; ++k) { wa[4 * hh + k] = va[k]; wb[4 * hh + k] = vb[k]; }
        }
      }
      const Real fa = trunc_a + ga, fb = trunc_b + gb;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ha[k] = rmin(oa[k], fa); hb[k] = rmin(ob[k], fb); }
#pragma unroll
      for (int d = 1; d <= 4 * HOPS; ++d) {
        const Real ca = rmin(Real(slope_a * Real(d)), trunc_a), cb = rmin(Real(slope_b * Real(d)), trunc_b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ha[k] = rmin(ha[k], rmin(Real(ca + wa[4 * HOPS + k - d]), Real(ca + wa[4 * HOPS + k + d])));
          hb[k] = rmin(hb[k], rmin(Real(cb + wb[4 * HOPS + k - d]), Real(cb + wb[4 * HOPS + k + d])));
        }
      }
    } else {
      const Real fa = trunc_a + ga, fb = trunc_b + gb;
#pragma unroll
      for (int k = 0; k < 4; ++k) { ha[k] = rmin(oa[k], fa); hb[k] = rmin(ob[k], fb); }
      for (int d = 1; d < L; ++d) {
        const Real ca = rmin(Real(slope_a * Real(d)), trunc_a), cb = rmin(Real(slope_b * Real(d)), trunc_b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ha[k] = rmin(ha[k], rmin(Real(ca + lds1<Real>(sa + (uint32_t)(k - d) * S)), Real(ca + lds1<Real>(sa + (uint32_t)(k + d) * S))));
          hb[k] = rmin(hb[k], rmin(Real(cb + lds1<Real>(sb + (uint32_t)(k - d) * S)), Real(cb + lds1<Real>(sb + (uint32_t)(k + d) * S))));
        }
      }
    }
    __syncwarp();   // the staging areas are rewritten by the next call
  }
}

template <typename Real, int PW, int L, int HOPS, bool ANISO, int KSEL, bool RINGS>
__global__ void __launch_bounds__(RINGS ? 32 * WavefrontCfg<Real, PW, L, KSEL>::MAX_WARPS : 32) wavefront_pass_kernel(const __grid_constant__ CUtensorMap tm_rec, const __grid_constant__ CUtensorMap tm_rec1,
                                                            const __grid_constant__ CUtensorMap tm_par, const GridArgs<Real> A) {
  using Cfg = WavefrontCfg<Real, PW, L, KSEL>;
  constexpr int G = Cfg::G, T = Cfg::T, P = Cfg::P, K = Cfg::K, PD = Cfg::PD, LOGK = Cfg::LOGK, REC = Cfg::REC, PREC = Cfg::PREC, NSLOT = Cfg::NSLOT, PDC = Cfg::PDC;
  constexpr int BOXROWS = ANISO ? T : T + 1;           // the uniform modes also read the first row of the next band
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  constexpr uint32_t TILE_BYTES = (uint32_t)((BOXROWS * K * REC + T * K * PREC) * sizeof(Real));
  constexpr uint32_t REC_B = REC * S, PREC_B = PREC * S, SLOT_REC_B = Cfg::SLOT_REC * S, SLOT_PAR_B = Cfg::SLOT_PAR * S, LB = L * S;
  constexpr bool PARTIAL_WARP = T * G < 32;            // tiny L: lanes beyond T groups shadow group T-1 and never store
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // RINGS: nw > 1 warps per CTA, warp w works on band sb * nw + w; else one warp = one band, no shared-memory hand-off
  const int wid = RINGS ? (int)(threadIdx.x >> 5) : 0, nw = RINGS ? (int)(blockDim.x >> 5) : 1;
  unsigned char* const smem_warp = smem_raw + (size_t)wid * Cfg::WARP_SMEM;
  unsigned char* const smem_cta = smem_raw + (size_t)nw * Cfg::WARP_SMEM;  // super-band queue, then the hand-off rings
  volatile unsigned long long* const sb_queue = reinterpret_cast<volatile unsigned long long*>(smem_cta);
  Real* const rings = reinterpret_cast<Real*>(smem_cta + 128);            // [nw-1][RING][L], ring w: warp w -> warp w+1
  const uint32_t smem_base = smem_u32(smem_warp);
  const uint32_t bars = smem_base;                                     // NSLOT mbarriers
  const uint32_t rec_ring = smem_base + 128;
  const uint32_t par_ring = rec_ring + NSLOT * SLOT_REC_B;
  const uint32_t scratch_all = par_ring + NSLOT * SLOT_PAR_B;
  const int lane = threadIdx.x & 31;
  const bool lane_used = !PARTIAL_WARP || lane / G < T;
  const int t = lane_used ? lane / G : T - 1, gl = lane % G;
  const bool fwd = A.forward != 0;
  const uint32_t H = A.H, W = pin(A.W);
  const int nchunks = ((int)W + K - 1) >> LOGK;
  const int nb = (int)A.num_bands;
  // TMA tile stores must not start at negative coordinates, so tiles are anchored at row 0 / column 0 in BOTH
  // directions: backward, the partially filled tile column / band is the first one processed.  qoff / roff = number
  // of padding columns / rows in front of virtual column / row 0; q' = q + qoff and u' = u + roff index the tiles.
  const int qoff = fwd ? 0 : nchunks * K - (int)W;
  const int roff = fwd ? 0 : nb * T - (int)H;
  const int sp_end = qoff + (int)W + T - 1;
  const CUtensorMap* const tm_load = ANISO ? &tm_rec : &tm_rec1;
  // two staging areas per lane group: [INF pad L | L labels | INF pad L]
  const uint32_t stage_a = scratch_all + (uint32_t)((2 * (lane / G)) * 3 * L + L + 4 * gl) * S;
  const uint32_t stage_b = stage_a + 3 * LB;

  if (lane == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
    Real* sc = reinterpret_cast<Real*>(smem_warp + 128 + NSLOT * (SLOT_REC_B + SLOT_PAR_B));
    for (int i = lane; i < Cfg::SCRATCH; i += 32) sc[i] = RealTraits<Real>::inf();
  }
  if constexpr (RINGS) {
    for (int i = (int)threadIdx.x; i < (nw - 1) * Cfg::RING * L; i += (int)blockDim.x) rings[i] = sentinel<Real>();
    if (threadIdx.x < 16) sb_queue[threadIdx.x] = 0ull;
    __syncthreads();
  }
  __syncwarp();
  const unsigned num_super = (A.num_bands + (unsigned)nw - 1u) / (unsigned)nw;
  uint32_t chunk_base = 0;
  uint32_t ring_in_n = 0, ring_out_n = 0;   // columns through the ring before / after this warp so far: slots are used cyclically   // tiles loaded by earlier bands of this warp: slot / mbarrier-phase sequence continues

  // record vector / parameter roles in this direction (byte offsets inside a record)
  const uint32_t o_in_row = (fwd ? REC_BL : REC_AR) * LB, o_in_col = (fwd ? REC_BU : REC_AD) * LB;
  const uint32_t o_out_row = (fwd ? REC_AR : REC_BL) * LB, o_out_col = (fwd ? REC_AD : REC_BU) * LB;
  const uint32_t p_in_row = (fwd ? PAR_L : PAR_R) * P * S, p_in_col = (fwd ? PAR_U : PAR_D) * P * S;
  const uint32_t p_out_row = (fwd ? PAR_R : PAR_L) * P * S, p_out_col = (fwd ? PAR_D : PAR_U) * P * S;
  const int kxor = fwd ? 0 : K - 1;
  const int trow = fwd ? t : BOXROWS - 1 - t;            // row of this lane group inside the record tile
  const int trow_par = fwd ? t : T - 1 - t;
  const uint32_t rec_lane = rec_ring + (uint32_t)(trow * K) * REC_B + (uint32_t)(gl * 4) * S;
  const uint32_t par_lane = par_ring + (uint32_t)(trow_par * K) * PREC_B;
  const int below = fwd ? (int)(K * REC_B) : -(int)(K * REC_B);   // smem distance to the same column of virtual row u+1
  // receive order of the two processed-side edges: LIFO: (L, U) forward, (D, R) backward
  const bool recv_row_first = fwd ? !A.fifo : (A.fifo != 0);
  // weight of the FAST path: (n_in, n_out) = (2, 2), first row (1, 2): 1 / max(n_in, n_out) = 1/2 (include/LP_MP.h:1397)
  const Real w_fast = (Real)(1.0 / 2.0);
  const bool H_total_ge2 = H >= 2 || A.in_halo || A.out_halo;   // a single-row grid has (n_in, n_out) = (1, 1): weight 1
  const bool potts_nonneg = A.potts_nonneg != 0;

  for (uint32_t iter = 0;; ++iter) {
    // warp 0 claims the next super-band and publishes it to the other warps of the CTA ({iteration + 1, index} in a
    // queue of 16 entries: the rings' back-pressure keeps the warps of a CTA within two iterations of each other)
    unsigned sb = 0;
    if (lane == 0) {
      if (wid == 0) {
        sb = atomicAdd(A.control, 1u);
        if (RINGS) sb_queue[iter & 15u] = ((unsigned long long)(iter + 1u) << 32) | sb;
      } else {
        unsigned long long v;
        do { v = sb_queue[iter & 15u]; } while ((uint32_t)(v >> 32) != iter + 1u);
        sb = (unsigned)v;
      }
    }
    sb = __shfl_sync(FULL, sb, 0);
    if (sb >= num_super) break;
    const unsigned kb = sb * (unsigned)nw + (unsigned)wid;
    if (kb >= A.num_bands) continue;   // the last super-band may be partial
    const int u = (int)kb * T + t - roff;                  // virtual row of this lane group
    const bool row_active = u >= 0 && u < (int)H && lane_used;
    const uint32_t r = fwd ? (uint32_t)u : H - 1 - (uint32_t)u;
    const bool in_col_exists = u > 0 || A.in_halo, out_col_exists = u + 1 < (int)H || A.out_halo;
    const bool last_band = kb + 1 == A.num_bands;
    const bool poll_in = kb > 0 || A.in_halo;              // group t_in takes its column neighbour from the hand-off buffer
    const int t_in = kb == 0 ? roff : 0;                   // first active group of the band
    const bool poller = poll_in && t == t_in && lane_used;
    const bool hand_out = !last_band || (A.out_halo && A.out_boundary != nullptr);
    // hand-off slots: shared-memory ring between warps of this CTA (column index wraps), else the L2 / peer buffers
    const bool ring_in = RINGS && wid > 0, ring_out = RINGS && !last_band && wid + 1 < nw;
    const uint32_t ring_in_s = smem_u32(rings + (size_t)(wid > 0 ? wid - 1 : 0) * Cfg::RING * L + gl * 4);
    const uint32_t ring_out_s = smem_u32(rings + (size_t)wid * Cfg::RING * L + gl * 4);
    auto in_slot = [&](int q) { return ring_in_s + ((ring_in_n + (uint32_t)q) & (uint32_t)(Cfg::RING - 1)) * LB; };
    auto out_slot = [&](int q) { return ring_out_s + ((ring_out_n + (uint32_t)q) & (uint32_t)(Cfg::RING - 1)) * LB; };
    const Real* const hand_in = (kb == 0 ? A.in_boundary : A.handoff + (uint64_t)kb * W * L) + gl * 4;
    Real* const hand_dst = (last_band ? A.out_boundary : A.handoff + (uint64_t)(kb + 1) * W * L) + gl * 4;
    const int t_hand = last_band ? (int)H - 1 - (int)kb * T + roff : T - 1;   // group that owns the band's last row
    const bool hander = hand_out && t == t_hand && lane_used;
    // tile origin rows (actual row of tile row 0); columns follow from the chunk index
    const int row0_own = fwd ? (int)kb * T : (nb - 1 - (int)kb) * T;
    const int row0_rec = fwd ? row0_own : row0_own - (BOXROWS - T);
    const uint32_t store_skip = fwd ? 0u : (uint32_t)(BOXROWS - T) * K * REC_B;   // backward the extra (read-only) tile row comes first
    const int lane_q0 = (int)(chunk_base << LOGK) - t;   // tile-ring column of this lane group at sp = 0

    auto issue_load = [&](int j) {   // lane 0 only
      const uint32_t slot = (chunk_base + (uint32_t)j) & (NSLOT - 1);
      const int col0 = fwd ? j * K : (nchunks - 1 - j) * K;
      const uint32_t bar = bars + 8 * slot;
      mbar_expect_tx(bar, TILE_BYTES);
      tma_load_4d(rec_ring + slot * SLOT_REC_B, tm_load, col0, row0_rec, bar);
      tma_load_3d(par_ring + slot * SLOT_PAR_B, &tm_par, col0, row0_own, bar);
    };
    auto wait_chunk = [&](int j) {   // all lanes
      const uint32_t n = chunk_base + (uint32_t)j;
      const uint32_t bar = bars + 8 * (n & (NSLOT - 1));
      const uint32_t parity = (n / NSLOT) & 1u;
      while (!mbar_try_wait(bar, parity)) {}
    };
    auto store_chunk = [&](int j) {  // all lanes: the tile is complete in shared memory
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const uint32_t slot = (chunk_base + (uint32_t)j) & (NSLOT - 1);
        const int col0 = fwd ? j * K : (nchunks - 1 - j) * K;
        tma_store_4d(&tm_rec, col0, row0_own, rec_ring + slot * SLOT_REC_B + store_skip);
        tma_commit();
      }
    };

    if (lane == 0) {
      for (int j = 0; j <= PDC && j < nchunks; ++j) issue_load(j);
    }
    auto stamp = [&](int which) {
      if (A.trace != nullptr && lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        A.trace[4ull * kb + which] = now;
      }
    };
    stamp(0);

    // the hand-off slot of column q+PD is polled right after column q has been consumed, PD steps before its use,
    // so that the L2 round trip (336 cycles idle, several times that under load) is off the dependency chain
    Real polled[PD][4];
#pragma unroll
    for (int d = 0; d < PD; ++d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) polled[d][k] = Real(0);
      ld_relaxed_if<HAND_BYTES>(poller && !ring_in && d < (int)W, polled[d], hand_in + d * L);
    }
    // a shared-memory ring answers in ~30 cycles: its slots are read ONE step ahead (into polled[0])
    if constexpr (RINGS) lds_hand_if<HAND_BYTES>(poller && ring_in, polled[0], in_slot(0));

    Real carry_row[4], carry_col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { carry_row[k] = Real(0); carry_col[k] = Real(0); }
    int next_store = 0;
    // uniform modes on a stripe: the row below the last virtual row lives on the neighbouring device; its in-side vectors
    // (previous pass) wait in `later_boundary`, read one step ahead
    const bool later_band = !ANISO && last_band && A.out_halo && A.later_boundary != nullptr;   // warp-uniform
    const bool later_poller = later_band && lane_used && t == (int)H - 1 - (int)kb * T + roff;
    const Real* const later_in = A.later_boundary + gl * 4;
    Real later_pf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) later_pf[k] = Real(0);
    if constexpr (!ANISO) ld_relaxed_if<HAND_BYTES>(later_poller, later_pf, later_in);

    wait_chunk(0);
    if (!ANISO && nchunks > 1) wait_chunk(1);   // uniform modes: the out_row neighbour may sit in the next tile
    stamp(1);

    // sp = step index shifted by qoff; group t is at tile column q' = sp - t, virtual column q = q' - qoff
    auto do_step = [&](auto fast_tag, const int sp) {
      constexpr bool FAST = decltype(fast_tag)::value;
      // ---- tile ring management (warp-uniform, one branch per step) ----
      const bool ev_store = ((sp - T + 1) & (K - 1)) == 0 && sp >= T + K - 1;   // the previous step completed a tile
      const bool ev_load = (sp & (K - 1)) == 0 && sp > 0;                        // group 0 enters a new tile
      // back-pressure of the outgoing ring, once per RING/2 columns: slots are written and freed in cyclic order, so when
      // the slot of the last column of the next half ring holds the sentinel, all slots up to it are free
      const int q_h = sp - t_hand - qoff;        // column of the band's last row
      const bool ev_ring = RINGS && ring_out && q_h >= 0 && q_h < (int)W && (q_h == 0 || ((ring_out_n + (uint32_t)q_h) & (uint32_t)(Cfg::RING / 2 - 1)) == 0);
      if (ev_store || ev_load || ev_ring) {
        if constexpr (RINGS) {
          if (ev_ring) {
            if (hander) {
              const int rem = Cfg::RING / 2 - 1 - (int)((ring_out_n + (uint32_t)q_h) & (uint32_t)(Cfg::RING / 2 - 1));
              const int qa = q_h + rem < (int)W ? q_h + rem : (int)W - 1;
              Real f[4];
              unsigned spins = 0;
              for (;;) {
                lds_hand<HAND_BYTES>(f, out_slot(qa));
                if (is_sentinel(f[0]) && is_sentinel(f[1]) && is_sentinel(f[2]) && is_sentinel(f[3])) break;
                if ((++spins & 0x3ffu) == 0) {
                  if (spins > (1u << 27)) atomicExch(A.control + 1, 1u);
                  if (*reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u) break;
                }
              }
            }
            __syncwarp();
          }
        }
        if (ev_store) { store_chunk(next_store); ++next_store; }
        if (ev_load) {
          if (sp == K) stamp(2);
          const int j = sp >> LOGK;
          const int jw = ANISO ? j : j + 1;
          if (jw < nchunks) wait_chunk(jw);
          if (lane == 0 && j + PDC < nchunks) {
            tma_wait_read<(T < 2 ? 1 : 0)>();     // the tile that last lived in this slot has left shared memory
            issue_load(j + PDC);
          }
          __syncwarp();
        }
      }
      const int q = sp - t - qoff;
      const int qq = sp + lane_q0;               // column in the tile ring (slot sequence continues across bands)
      const uint32_t slot = ((uint32_t)qq >> LOGK) & (NSLOT - 1);
      const uint32_t kk = ((uint32_t)qq & (K - 1)) ^ (uint32_t)kxor;
      const uint32_t rec = rec_lane + slot * SLOT_REC_B + kk * REC_B;
      const uint32_t par = par_lane + slot * SLOT_PAR_B + kk * PREC_B;

      Real par_ir[P], par_ic[P];
      lds_par<Real, P>(par_ir, par + p_in_row);
      lds_par<Real, P>(par_ic, par + p_in_col);
      Real th[4], own_ir[4], own_ic[4], own_or[4], own_oc[4];
      lds4(th, rec);
      lds4(own_ir, rec + o_in_row);
      lds4(own_ic, rec + o_in_col);
      lds4(own_or, rec + o_out_row);
      lds4(own_oc, rec + o_out_col);

      // ---- the column neighbour's vector: shuffle inside the band, hand-off slot for group t_in ----
      // (the common case -- the slot polled during the previous step is valid -- is branch-free for the warp)
      Real col_in[4];
      const bool polling = poller && q >= 0 && q < (int)W;
      const Real* const hslot = hand_in + (int64_t)q * L;
      // the wait is a warp-UNIFORM loop (every lane takes every iteration; the loads are predicated): lanes parked at a
      // reconvergence point during a long divergent spin wake up slowly (~1.4 us measured), and a band can never make
      // up time on its predecessor
      bool missing = polling && (is_sentinel(polled[0][0]) || is_sentinel(polled[0][1]) || is_sentinel(polled[0][2]) || is_sentinel(polled[0][3]));
      if (__any_sync(FULL, missing)) {
        unsigned spins = 0;
        do {
          if constexpr (RINGS) lds_hand_if<HAND_BYTES>(missing && ring_in, polled[0], in_slot(q));
          ld_relaxed_if<HAND_BYTES>(missing && !ring_in, polled[0], hslot);
          missing = polling && (is_sentinel(polled[0][0]) || is_sentinel(polled[0][1]) || is_sentinel(polled[0][2]) || is_sentinel(polled[0][3]));
          if ((++spins & 0x3ffu) == 0) {      // bounded wait: a lost producer must not hang the device
            if (spins > (ring_in ? (1u << 27) : (1u << 22))) atomicExch(A.control + 1, 1u);   // ~1 s either way
            if (*reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u) {
#pragma unroll
              for (int k = 0; k < 4; ++k) polled[0][k] = Real(0);
              missing = false;
            }
          }
        } while (__any_sync(FULL, missing));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) col_in[k] = polling ? polled[0][k] : carry_col[k];
      {
        Real sent[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) sent[k] = sentinel<Real>();
        st_relaxed_if<HAND_BYTES>(polling && !ring_in, const_cast<Real*>(hslot), sent);        // the slot is free again for the next pass
        if constexpr (RINGS) sts_hand_if<HAND_BYTES>(polling && ring_in, in_slot(q), sent);
        // the poll ring advances once the polling group has entered the grid (warp-uniform); slot q + PD is requested
        // now and consumed PD steps later, so PD L2 round trips are in flight per band
        if (sp - t_in - qoff >= 0) {
#pragma unroll
          for (int d = 0; d + 1 < PD; ++d) {
#pragma unroll
            for (int k = 0; k < 4; ++k) polled[d][k] = polled[d + 1][k];
          }
        }
        ld_relaxed_if<HAND_BYTES>(polling && !ring_in && q + PD < (int)W, polled[PD - 1], hslot + PD * L);
        if constexpr (RINGS) lds_hand_if<HAND_BYTES>(polling && ring_in && q + 1 < (int)W, polled[0], in_slot(q + 1));
      }

      Real out_col[4];
      if constexpr (FAST) {
        // ---- ReceiveMessages from the two processed neighbours, SendMessages to the two unprocessed ones ----
        Real h_ir[4], h_ic[4];
        lane_min_plus2<Real, PW, L, HOPS>(par_ir, carry_row, h_ir, par_ic, col_in, h_ic, stage_a, stage_b, gl, potts_nonneg);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real m_ir = h_ir[k] + own_ir[k], m_ic = h_ic[k] + own_ic[k];
          own_ir[k] = own_ir[k] - m_ir;
          own_ic[k] = own_ic[k] - m_ic;
          th[k] = recv_row_first ? (th[k] + m_ir) + m_ic : (th[k] + m_ic) + m_ir;
        }
        // missing edges: the left column (q = 0) and the first row have all-zero in-side slots and parameters, so their
        // receives are exact no-ops; the last grid row has no out_col edge (select); groups that have not reached
        // column 0 yet (q < 0) and rows beyond the grid compute on garbage and neither store nor carry
        const bool started = q >= 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real d = w_fast * th[k];   // every send uses the same snapshot value
          const Real d_oc = out_col_exists ? d : Real(0);
          own_oc[k] = own_oc[k] + d_oc;
          own_or[k] = own_or[k] + d;
          th[k] = (th[k] - d_oc) - d;
          carry_row[k] = started ? own_or[k] : Real(0);
          out_col[k] = own_oc[k];
        }
        if (row_active && started) {
          sts4(rec, th);
          sts4(rec + o_in_row, own_ir);
          sts4(rec + o_in_col, own_ic);
          sts4(rec + o_out_row, own_or);
          sts4(rec + o_out_col, own_oc);
        }
      } else {
        const bool act = row_active && q >= 0 && q < (int)W;
        const bool has_ir = act && q > 0, has_or = act && q + 1 < (int)W;
        const bool has_ic = act && in_col_exists, has_oc = act && out_col_exists;
        const int n_in = (int)has_ir + (int)has_ic, n_out = (int)has_or + (int)has_oc;
        // min-marginals of a pair of edges (independent of theta): m = h + own, own -= m   (RepamLeft / RepamRight)
        auto marginal2 = [&](bool on_a, const Real (&pr_a)[P], const Real (&oth_a)[4], Real (&own_a)[4], Real (&m_a)[4],
                             bool on_b, const Real (&pr_b)[P], const Real (&oth_b)[4], Real (&own_b)[4], Real (&m_b)[4]) {
          Real pza[P], pzb[P], oza[4], ozb[4], ha[4], hb[4];
#pragma unroll
          for (int p = 0; p < P; ++p) { pza[p] = on_a ? pr_a[p] : Real(0); pzb[p] = on_b ? pr_b[p] : Real(0); }
#pragma unroll
          for (int k = 0; k < 4; ++k) { oza[k] = on_a ? oth_a[k] : Real(0); ozb[k] = on_b ? oth_b[k] : Real(0); }
          lane_min_plus2<Real, PW, L, HOPS>(pza, oza, ha, pzb, ozb, hb, stage_a, stage_b, gl, potts_nonneg);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            m_a[k] = on_a ? Real(ha[k] + own_a[k]) : Real(0);
            own_a[k] = on_a ? Real(own_a[k] - m_a[k]) : own_a[k];
            m_b[k] = on_b ? Real(hb[k] + own_b[k]) : Real(0);
            own_b[k] = on_b ? Real(own_b[k] - m_b[k]) : own_b[k];
          }
        };
        Real m_ir[4], m_ic[4];
        marginal2(has_ir, par_ir, carry_row, own_ir, m_ir, has_ic, par_ic, col_in, own_ic, m_ic);
        if constexpr (ANISO) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            Real x = th[k];
            if (recv_row_first) { if (has_ir) x = x + m_ir[k]; if (has_ic) x = x + m_ic[k]; }
            else { if (has_ic) x = x + m_ic[k]; if (has_ir) x = x + m_ir[k]; }
            th[k] = x;
          }
        } else {
          // the other side of the edges towards unprocessed neighbours: their in-side vectors, still in the state of
          // the previous pass, in the neighbouring records of the tile ring
          Real oth_or[4], oth_oc[4], par_or[P], par_oc[P], m_or[4], m_oc[4];
          const int qq1 = qq + 1;
          const uint32_t rec_q1 = rec_lane + (((uint32_t)qq1 >> LOGK) & (NSLOT - 1)) * SLOT_REC_B + (((uint32_t)qq1 & (K - 1)) ^ (uint32_t)kxor) * REC_B;
          lds4(oth_or, rec_q1 + o_in_row);
          lds4(oth_oc, (uint32_t)((int)rec + below) + o_in_col);
          if (later_band) {
            const bool lpoll = later_poller && q >= 0 && q < (int)W;
            const Real* const lslot = later_in + (int64_t)q * L;
            bool lmiss = lpoll && (is_sentinel(later_pf[0]) || is_sentinel(later_pf[1]) || is_sentinel(later_pf[2]) || is_sentinel(later_pf[3]));
            if (__any_sync(FULL, lmiss)) {     // warp-uniform wait, bounded like the hand-off poll
              unsigned spins = 0;
              do {
                ld_relaxed_if<HAND_BYTES>(lmiss, later_pf, lslot);
                lmiss = lpoll && (is_sentinel(later_pf[0]) || is_sentinel(later_pf[1]) || is_sentinel(later_pf[2]) || is_sentinel(later_pf[3]));
                if ((++spins & 0x3ffu) == 0) {
                  if (spins > (1u << 22)) atomicExch(A.control + 1, 1u);
                  if (*reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) later_pf[k] = Real(0);
                    lmiss = false;
                  }
                }
              } while (__any_sync(FULL, lmiss));
            }
            Real sent[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { if (lpoll) oth_oc[k] = later_pf[k]; sent[k] = sentinel<Real>(); }
            st_relaxed_if<HAND_BYTES>(lpoll, const_cast<Real*>(lslot), sent);        // free for the delivery of the next pass
            ld_relaxed_if<HAND_BYTES>(lpoll && q + 1 < (int)W, later_pf, lslot + L);
          }
          lds_par<Real, P>(par_or, par + p_out_row);
          lds_par<Real, P>(par_oc, par + p_out_col);
          marginal2(has_or, par_or, oth_or, own_or, m_or, has_oc, par_oc, oth_oc, own_oc, m_oc);
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
          sts4(rec, th);
          sts4(rec + o_in_row, own_ir);
          sts4(rec + o_in_col, own_ic);
          sts4(rec + o_out_row, own_or);
          sts4(rec + o_out_col, own_oc);
        }
      }

      // ---- hand the freshly written column vector to the node below: next lane group, or the next band ----
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Real up = __shfl_up_sync(FULL, out_col[k], G);
        carry_col[k] = t <= t_in ? Real(0) : up;
      }
      const bool handing = hander && q >= 0 && q < (int)W;
      st_relaxed_if<HAND_BYTES>(handing && !ring_out, hand_dst + (int64_t)q * L, out_col);
      if constexpr (RINGS) sts_hand_if<HAND_BYTES>(handing && ring_out, out_slot(q), out_col);

    };

    // FAST path: anisotropic non-rounding passes, every group left of the last column (q <= W-2); the last band
    // (whose last row has (n_in, n_out) = (1, 1) at q = 0, weight 1 instead of 1/2) joins after its left ramp
    const bool fast_band = ANISO && !A.primal && W >= 3 && H_total_ge2;
    const int sp_fast0 = fast_band ? (last_band ? qoff + T : qoff) : sp_end, sp_fast1 = fast_band ? qoff + (int)W - 2 : sp_end;   // fast for sp in [fast0, fast1]
    for (int sp = qoff; sp < sp_end; ++sp) {
      if (sp >= sp_fast0 && sp <= sp_fast1) do_step(std::true_type{}, sp);
      else do_step(std::false_type{}, sp);
    }
    stamp(3);

    while (next_store < nchunks) { store_chunk(next_store); ++next_store; }
    if (lane == 0) tma_wait_read<0>();   // every tile of this band has left shared memory before the ring is reused
    __syncwarp();
    chunk_base += (uint32_t)nchunks;
    if (ring_in) ring_in_n += W;
    if (ring_out) ring_out_n += W;
  }
  if (lane == 0) tma_wait_all();   // the global writes of the last tiles are complete before the warp exits
}

}  // namespace lpmp

// Wavefront pass, "rows" variant for the anisotropic (SRMP) passes of latency-bound grids: ONE LABEL PER LANE.
//
// Replaces LP::ComputePass over forward/backwardUpdateOrdering_ (/root/reference/include/LP_MP.h:981-1005,1592-1602) and
// FactorContainer::UpdateFactor (include/factors_messages.hxx:2256-2261), like the other wavefront kernels, with the same
// storage (node records rec[H][W][5][L], parameter records par[H][W][4][P]), the same TMA tiles and the same self-validating
// hand-off slots.  What changes is the mapping of a node update onto the machine:
//
//   A pass in the reference order is a dependency chain of H + W - 1 node updates, so its duration is
//   W * tau_x + H * tau_y, where tau_x / tau_y is the latency of one update as seen by the row / column neighbour -- NOT
//   the amount of arithmetic.  The band kernels (kernels_grid.cuh, kernels_grid_lean.cuh) give a node L/4 lanes with 4 labels
//   each and a warp 8 rows: ~300 instructions of one warp per step, tau ~ 880 cycles (profiles/r01d_steps.txt).  Here a node
//   gets L lanes with ONE label each and a warp RPW = 32/L rows (2 for L = 16), so that
//     * the label window of the truncated-linear min-plus is 2*WIN independent shuffles (no shared-memory staging, no
//       __syncwarp), the group minimum is ONE instruction (redux.sync.min.f32 = CREDUX.MIN.F32 on sm_100a, 19 cycles measured
//       against 4 dependent shuffle+min rounds = 112; scripts/repro/lat.cu), and the whole step is ~100 instructions;
//     * the rows of a warp hand their column vectors on by one shuffle (24 cycles); consecutive warps of a CTA hand on
//       through rings of self-validating slots in shared memory (~160 cycles one way, measured), consecutive CTAs through the
//       L2 hand-off slots (an in-mover warp polls them ahead of time, as in the lean kernel).
//
// CTA = NW compute warps (band of T = NW * RPW consecutive rows) + one in-mover warp + one tile warp.  The tile warp issues
// all TMA tile loads / stores of the CTA (one ring of NSLOT tiles of RPW rows x K columns per compute warp); compute warps
// never form a global address except for the single store that hands the band's last row on to the next band.
//
// The arithmetic of a node update is the same sequence of IEEE operations as in the other kernels (bit-identical state).
#pragma once
#include "kernels_grid.cuh"

namespace lpmp {

template <typename Real, int PW, int L>
struct RowsCfg {
  static constexpr int RPW = 32 / L;                       // rows per compute warp (one label per lane)
  static constexpr int NW = 4;                             // compute warps per CTA
  static constexpr int T = NW * RPW;                       // rows per band
  static constexpr int K = 8, LOGK = 3;                    // columns per tile
  static constexpr int P = ParamCount<PW>::value;
  static constexpr int REC = REC_VECS * L, PREC = PAR_EDGES * P;
  static constexpr int SLOT_REC = RPW * K * REC, SLOT_PAR = RPW * K * PREC;   // reals per tile
  static constexpr int RING = 32;                          // columns per hand-off ring (power of two, >= 4 tiles of columns)
  static constexpr size_t SLOT_BYTES = (size_t)(SLOT_REC + SLOT_PAR) * sizeof(Real);
  static constexpr size_t GSTORE_OFF = 384;                // [0,256) mbarriers, [256,384) control words, [384,640) tile warp's bookkeeping
  static constexpr size_t RINGS_OFF = 640;
  static constexpr size_t TILES_OFF = (RINGS_OFF + (size_t)NW * RING * L * sizeof(Real) + 127) / 128 * 128;
  static constexpr int NSLOT = TILES_OFF + 8 * NW * SLOT_BYTES <= 200 * 1024 ? 8 : 4;   // tiles per ring (power of two)
  static constexpr size_t WARP_TILE_BYTES = (size_t)NSLOT * SLOT_BYTES;
  static constexpr size_t SMEM = TILES_OFF + NW * WARP_TILE_BYTES;
  static constexpr int THREADS = 32 * (NW + 2);
  static constexpr bool FITS = L <= 32 && 32 % L == 0 && L >= 8 && SMEM <= 220 * 1024 && NW * NSLOT * 8 <= 256;
  static_assert((SLOT_REC * sizeof(Real)) % 128 == 0 && (SLOT_PAR * sizeof(Real)) % 128 == 0, "TMA destination alignment");
};

__device__ __forceinline__ float credux_min(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// minima of two values over the L-lane segments of the warp (the two chains are interleaved)
template <typename Real, int L>
__device__ __forceinline__ void seg_min2(Real& a, Real& b) {
  if constexpr (std::is_same<Real, float>::value && L == 32) {
    a = credux_min(a);
    b = credux_min(b);
  } else if constexpr (std::is_same<Real, float>::value && L == 16) {
    const bool lo = (threadIdx.x & 16u) == 0u;
    const float INF = RealTraits<float>::inf();
    const float a0 = credux_min(lo ? a : INF), a1 = credux_min(lo ? INF : a);
    const float b0 = credux_min(lo ? b : INF), b1 = credux_min(lo ? INF : b);
    a = lo ? a0 : a1;
    b = lo ? b0 : b1;
  } else {
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) {
      const Real ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
      a = rmin(a, ta);
      b = rmin(b, tb);
    }
  }
}
template <typename Real, int L>
__device__ __forceinline__ Real seg_min1(Real a) {
  Real b = a;
  seg_min2<Real, L>(a, b);
  return a;
}

// Min-plus of two edges, one label per lane: h(x) = min_y [c(x,y) + o(y)], x = lane % L.
// Truncated linear (slope, trunc >= 0): terms with |x - y| > WIN cost trunc (host guarantees slope * (WIN + 1) >= trunc), so they
// collapse to trunc + min(o); the near terms are 2 * WIN segment shuffles -- a shuffle that leaves the segment returns the
// lane's own value, and c_d + o(x) >= o(x) is harmless.  Potts: WIN = 1: all lambda >= 0 (one minimum); WIN = 0: two smallest.
template <typename Real, int PW, int L, int WIN>
__device__ __forceinline__ void rows_min_plus2(const Real* pa, Real oa, Real& ha, const Real* pb, Real ob, Real& hb) {
  constexpr unsigned FULL = 0xffffffffu;
  Real ga = oa, gb = ob;
  if constexpr (PW == LPMP_PAIRWISE_POTTS) {
    const Real la = pa[0], lb = pb[0];
    if (WIN == 0 && __any_sync(FULL, la < Real(0) || lb < Real(0))) {
      const Real INF = RealTraits<Real>::inf();
      const unsigned lane = threadIdx.x & 31u;
      const unsigned segmask = (L == 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u)) << (lane & ~(unsigned)(L - 1));
      seg_min2<Real, L>(ga, gb);
      const unsigned ia = (unsigned)__ffs((int)(__ballot_sync(FULL, oa == ga) & segmask)) - 1u;
      const unsigned ib = (unsigned)__ffs((int)(__ballot_sync(FULL, ob == gb) & segmask)) - 1u;
      Real sa = lane == ia ? INF : oa, sb = lane == ib ? INF : ob;
      seg_min2<Real, L>(sa, sb);
      ha = rmin(oa, Real(la + (lane == ia ? sa : ga)));
      hb = rmin(ob, Real(lb + (lane == ib ? sb : gb)));
    } else {
      seg_min2<Real, L>(ga, gb);
      ha = rmin(oa, Real(la + ga));
      hb = rmin(ob, Real(lb + gb));
    }
  } else {
    const Real slope_a = pa[0], trunc_a = pa[1], slope_b = pb[0], trunc_b = pb[1];
    Real ua[WIN], da[WIN], ub[WIN], db[WIN];
#pragma unroll
    for (int d = 1; d <= WIN; ++d) {
      ua[d - 1] = __shfl_up_sync(FULL, oa, d, L);
      da[d - 1] = __shfl_down_sync(FULL, oa, d, L);
      ub[d - 1] = __shfl_up_sync(FULL, ob, d, L);
      db[d - 1] = __shfl_down_sync(FULL, ob, d, L);
    }
    seg_min2<Real, L>(ga, gb);
    ha = rmin(oa, Real(trunc_a + ga));
    hb = rmin(ob, Real(trunc_b + gb));
#pragma unroll
    for (int d = 1; d <= WIN; ++d) {
      const Real ca = rmin(Real(slope_a * Real(d)), trunc_a), cb = rmin(Real(slope_b * Real(d)), trunc_b);
      ha = rmin(ha, rmin(Real(ca + ua[d - 1]), Real(ca + da[d - 1])));
      hb = rmin(hb, rmin(Real(cb + ub[d - 1]), Real(cb + db[d - 1])));
    }
  }
}

// scalar accesses of the compute warps (volatile: program order among them is kept; tiles are also written by the TMA)
template <typename Real> __device__ __forceinline__ Real lds1v(uint32_t a);
template <> __device__ __forceinline__ float lds1v<float>(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
template <> __device__ __forceinline__ double lds1v<double>(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
template <typename Real> __device__ __forceinline__ Real lds1_volatile(uint32_t a);
template <> __device__ __forceinline__ float lds1_volatile<float>(uint32_t a) { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
template <> __device__ __forceinline__ double lds1_volatile<double>(uint32_t a) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts1_if(bool on, uint32_t a, float v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f32 [%0], %1; }" ::"r"(a), "f"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void sts1_if(bool on, uint32_t a, double v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void lds1_volatile_if(bool on, float& v, uint32_t a) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.volatile.shared.f32 %0, [%1]; }" : "+f"(v) : "r"(a), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void lds1_volatile_if(bool on, double& v, uint32_t a) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.volatile.shared.f64 %0, [%1]; }" : "+d"(v) : "r"(a), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void sts1_volatile_if(bool on, uint32_t a, float v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.volatile.shared.f32 [%0], %1; }" ::"r"(a), "f"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void sts1_volatile_if(bool on, uint32_t a, double v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.volatile.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void stg1_relaxed_if(bool on, float* p, float v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.relaxed.sys.global.f32 [%0], %1; }" ::"l"(p), "f"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void stg1_relaxed_if(bool on, double* p, double v) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.relaxed.sys.global.f64 [%0], %1; }" ::"l"(p), "d"(v), "r"((uint32_t)on) : "memory");
}

// bounded waits: a lost partner must not hang the device.  Called every 2^10 .. 2^16 spins of a wait loop: true once the sticky
// abort flag is set (checked by the host after the pass) or the kernel has been running for kWaitLimitNs
constexpr unsigned long long kWaitLimitNs = 4000000000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
static __device__ __noinline__ bool wait_expired(unsigned int* control, unsigned long long t_launch) {
  if (*reinterpret_cast<volatile unsigned int*>(control + 1) != 0u) return true;
  if (globaltimer_ns() - t_launch > kWaitLimitNs) { atomicExch(control + 1, 1u); return true; }
  return false;
}

template <typename Real, int PW, int L, int WIN, bool FWD>
__global__ void __launch_bounds__(RowsCfg<Real, PW, L>::THREADS, 1) wavefront_rows_kernel(const __grid_constant__ CUtensorMap tm_rec, const __grid_constant__ CUtensorMap tm_par,
                                                                                         const GridArgs<Real> A) {
  using Cfg = RowsCfg<Real, PW, L>;
  constexpr int RPW = Cfg::RPW, NW = Cfg::NW, T = Cfg::T, K = Cfg::K, LOGK = Cfg::LOGK, P = Cfg::P, REC = Cfg::REC, PREC = Cfg::PREC, NSLOT = Cfg::NSLOT, RING = Cfg::RING;
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  constexpr uint32_t REC_B = REC * S, PREC_B = PREC * S, SLOT_REC_B = Cfg::SLOT_REC * S, SLOT_PAR_B = Cfg::SLOT_PAR * S, LB = L * S;
  constexpr uint32_t TILE_BYTES = SLOT_REC_B + SLOT_PAR_B;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  constexpr bool fwd = FWD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bars = smem_base;                         // [NW][NSLOT] mbarriers
  const uint32_t ctrl = smem_base + 256;                   // +0: band ticket, +16 + 4 w: steps completed by compute warp w (monotone over the kernel)
  const uint32_t rings = smem_base + (uint32_t)Cfg::RINGS_OFF;   // [NW][RING][L]: ring w feeds the first row of compute warp w
  const uint32_t tiles = smem_base + (uint32_t)Cfg::TILES_OFF;   // per compute warp: [NSLOT] record tiles, then [NSLOT] parameter tiles
  const int lane = threadIdx.x & 31, wid = (int)(threadIdx.x >> 5);
  const uint32_t H = A.H, W = pin(A.W);
  const int nchunks = ((int)W + K - 1) >> LOGK;
  const int nb = (int)A.num_bands;
  const int qoff = fwd ? 0 : nchunks * K - (int)W;         // padding columns / rows in front of virtual column / row 0
  const int roff = fwd ? 0 : nb * T - (int)H;              // (tiles are anchored at row 0 / column 0 in both directions)
  const int nsteps = (int)W + RPW - 1;
  constexpr int kxor = fwd ? 0 : K - 1;
  // ring slot of a column: its tile-space index q' = q + qoff modulo RING (so that tiles and ring slots wrap together)

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * NSLOT; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    Real* rg = reinterpret_cast<Real*>(smem_raw + Cfg::RINGS_OFF);
    for (int i = (int)threadIdx.x; i < NW * RING * L; i += Cfg::THREADS) rg[i] = sentinel<Real>();
    if (threadIdx.x < 32) reinterpret_cast<volatile uint32_t*>(smem_raw + 256)[threadIdx.x] = 0u;
  }
  __syncthreads();

  // bounded waits: a lost partner must not hang the device (sticky abort flag, checked by the host after the pass)
  const unsigned long long t_launch = globaltimer_ns();
  auto timed_out = [&](unsigned& spins, unsigned) {
    if ((++spins & 0x3ffu) != 0) return false;
    return wait_expired(A.control, t_launch);
  };
  // rows [kb T - roff + w RPW, + RPW) of virtual band kb: does compute warp w own a grid row?
  auto warp_live = [&](unsigned kb, int w) {
    const int u0 = (int)kb * T + w * RPW - roff;
    return u0 + RPW > 0 && u0 < (int)H;
  };

  uint32_t step_base = 0;                 // steps of earlier bands: the published step counters are monotone
  uint32_t chunk_base[NW];                // per compute warp: tiles of earlier bands (slot / mbarrier-phase sequence continues)
#pragma unroll
  for (int w = 0; w < NW; ++w) chunk_base[w] = 0;

  for (;;) {
    if (threadIdx.x == 0) sts_volatile_u32(ctrl, atomicAdd(A.control, 1u));
    __syncthreads();
    const unsigned kb = lds_volatile_u32(ctrl);
    if (kb >= A.num_bands) break;
    const int t_in = kb == 0 ? roff : 0;                   // first grid row of the band (virtual row inside the band)
    const bool band_in = kb > 0 || A.in_halo;              // row t_in has a column neighbour in the previous band / stripe

    if (wid == NW + 1) {
      // ================= tile warp: bulk loads as slots become free, bulk stores as the compute warps complete tiles =================
      // LANE w serves compute warp w (its own ring of NSLOT tiles, its own bulk groups): the bookkeeping of all rings is one SIMT
      // instruction stream -- a single thread serving every ring in turn was the bottleneck of the whole CTA (profiles/r02_split_tilewarp.txt)
      constexpr int DROP = NSLOT >= 8 ? 2 : 1;   // tiles of look-ahead given up so that a load never waits for the store just issued
      const int w = lane < NW ? lane : NW - 1;
      bool open = lane < NW && warp_live(kb, w);
      uint32_t cb = 0;
#pragma unroll
      for (int k = 0; k < NW; ++k) cb = k == w ? chunk_base[k] : cb;
      const int row0 = (int)((fwd ? kb : (unsigned)nb - 1u - kb) * T) + (fwd ? w : NW - 1 - w) * RPW;
      const uint32_t rec_ring = tiles + (uint32_t)w * (uint32_t)Cfg::WARP_TILE_BYTES, par_ring = rec_ring + NSLOT * SLOT_REC_B;
      const uint32_t ctr_w = ctrl + 16 + 4 * (uint32_t)w;
      const bool mine = open;
      int j_load = 0, j_store = 0;
      unsigned spins = 0;
      while (__any_sync(FULL, open)) {
        bool moved = false;
        // ---- stores: a tile is final when the last row of the warp has left it ----
        const int c = (int)(lds_volatile_u32(ctr_w) - step_base);
        int need = (j_store + 1) * K + RPW - 1 - qoff;
        if (need > nsteps) need = nsteps;
        if (open && c >= need) {
          const uint32_t slot = (cb + (uint32_t)j_store) & (NSLOT - 1);
          const int col0 = fwd ? j_store * K : (nchunks - 1 - j_store) * K;
          tma_store_4d(&tm_rec, col0, row0, rec_ring + slot * SLOT_REC_B);
          tma_commit();
          ++j_store;
          moved = true;
        }
        // ---- loads: into a slot whose previous tile has left shared memory (never the slot of the store just issued) ----
        if (open && j_load < nchunks && j_load < j_store + NSLOT - DROP) {
          const uint32_t slot = (cb + (uint32_t)j_load) & (NSLOT - 1);
          if (j_load >= NSLOT) tma_wait_read<DROP>();   // at least DROP younger stores of this lane were committed after the one that read the slot
          const int col0 = fwd ? j_load * K : (nchunks - 1 - j_load) * K;
          const uint32_t bar = bars + 8 * ((uint32_t)w * NSLOT + slot);
          mbar_expect_tx(bar, TILE_BYTES);
          tma_load_4d(rec_ring + slot * SLOT_REC_B, &tm_rec, col0, row0, bar);
          tma_load_3d(par_ring + slot * SLOT_PAR_B, &tm_par, col0, row0, bar);
          ++j_load;
          moved = true;
        }
        open = open && j_store < nchunks;
        if (__any_sync(FULL, moved)) spins = 0;
        else { __nanosleep(100); if (timed_out(spins, 1u << 26)) break; }
      }
      if (mine) tma_wait_read<0>();   // every tile of this band has left shared memory before the rings are reused
      __syncwarp();
    } else if (wid == NW) {
      // ====== in-mover: predecessor's hand-off slots (L2) -> ring of the band's first grid row (zeros if there is no predecessor) ======
      constexpr int G4 = L / 4, TG = 32 / G4;               // lanes per column (4 labels each), columns per load
      constexpr int NBATCH = TG >= 16 ? 1 : 16 / TG;        // >= 16 columns polled per L2 round trip
      const int g = lane / G4, gl = lane % G4;
      const int w_in = t_in / RPW;                          // the compute warp that owns row t_in
      const uint32_t ring = rings + (uint32_t)w_in * RING * LB;
      const int hr_in = t_in % RPW;                         // that row processes column q at step q + hr_in
      const uint32_t ctr_in = ctrl + 16 + 4 * (uint32_t)w_in;
      int q_in = 0;
      unsigned spins = 0;
      if (!band_in) {
        Real zero[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) zero[k] = Real(0);
        while (q_in < (int)W) {
          const int c = (int)(lds_volatile_u32(ctr_in) - step_base);
          int limit = (c - hr_in > 0 ? c - hr_in : 0) + RING;
          if (limit > (int)W) limit = (int)W;
          if (q_in < limit) {
            for (int q = q_in + g; q < limit; q += TG) sts4(ring + (((uint32_t)(q + qoff) & (uint32_t)(RING - 1)) * L + (uint32_t)(gl * 4)) * S, zero);
            q_in = limit;
            spins = 0;
          } else { __nanosleep(200); if (timed_out(spins, 1u << 26)) break; }
        }
      } else {
        const Real* const hand_in = (kb == 0 ? A.in_boundary : A.handoff + (uint64_t)kb * W * L) + gl * 4;
        while (q_in < (int)W) {
          const int c = (int)(lds_volatile_u32(ctr_in) - step_base);
          int limit = (c - hr_in > 0 ? c - hr_in : 0) + RING;   // the consumer has consumed (and re-armed) c - hr_in columns
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
      if (lane == 0) sts_volatile_u32(ctrl + 16 + 4 * (uint32_t)wid, step_base + (uint32_t)nsteps);
    } else {
      // =================================================== compute warps ===================================================
      const int w = wid;
      const int hr = lane / L, x = lane % L;                // row of this lane inside the warp, label
      const int t = w * RPW + hr;
      const int u = (int)kb * T + t - roff;                 // virtual grid row
      const bool row_active = u >= 0 && u < (int)H;
      const uint32_t r = fwd ? (uint32_t)u : H - 1 - (uint32_t)u;
      const bool in_col_exists = row_active && (u > 0 || A.in_halo), out_col_exists = row_active && (u + 1 < (int)H || A.out_halo);
      const bool poller = row_active && (t == t_in || hr == 0);           // takes its column neighbour's vector from ring w
      const bool from_shfl = row_active && !poller;                        // ... from the row above in the same warp
      // where this row's outgoing column vector goes (besides the shuffle to the next row of the warp)
      const bool grid_last_row = row_active && u + 1 == (int)H;
      const bool out_ring = row_active && !grid_last_row && hr == RPW - 1 && w + 1 < NW;
      const bool out_glob = row_active && (grid_last_row ? (A.out_halo && A.out_boundary != nullptr) : (hr == RPW - 1 && w + 1 == NW));
      Real* const out_ptr = (grid_last_row ? A.out_boundary : A.handoff + (uint64_t)(kb + 1) * W * L) + x;
      const uint32_t ring_in = rings + (uint32_t)w * RING * LB + (uint32_t)x * S;
      const uint32_t ring_out = rings + (uint32_t)(w + 1 < NW ? w + 1 : w) * RING * LB + (uint32_t)x * S;
      const bool warp_out_ring = __any_sync(FULL, out_ring);
      const uint32_t rec_ring = tiles + (uint32_t)w * (uint32_t)Cfg::WARP_TILE_BYTES, par_ring = rec_ring + NSLOT * SLOT_REC_B;
      constexpr uint32_t o_in_row = (fwd ? REC_BL : REC_AR) * LB, o_in_col = (fwd ? REC_BU : REC_AD) * LB;
      constexpr uint32_t o_out_row = (fwd ? REC_AR : REC_BL) * LB, o_out_col = (fwd ? REC_AD : REC_BU) * LB;
      constexpr uint32_t p_in_row = (fwd ? PAR_L : PAR_R) * P * S, p_in_col = (fwd ? PAR_U : PAR_D) * P * S;
      const int trow = fwd ? hr : RPW - 1 - hr;             // row of this lane inside the tiles
      const uint32_t rec_lane = rec_ring + (uint32_t)(trow * K) * REC_B + (uint32_t)x * S;
      const uint32_t par_lane = par_ring + (uint32_t)(trow * K) * PREC_B;
      const bool recv_row_first = pin((uint32_t)(fwd ? !A.fifo : (A.fifo != 0))) != 0u;
      const Real w_fast = (Real)(1.0 / 2.0);
      const uint32_t cbase = chunk_base[w];
      const int lane_q0 = (int)(cbase << LOGK) - hr + qoff; // tile-ring column of this lane at step 0
      const uint32_t ctr = ctrl + 16 + 4 * (uint32_t)w;
      const uint32_t ctr_next = ctrl + 16 + 4 * (uint32_t)(w + 1 < NW ? w + 1 : w);

      auto wait_chunk = [&](int j) {
        const uint32_t n = cbase + (uint32_t)j;
        const uint32_t bar = bars + 8 * ((uint32_t)w * NSLOT + (n & (NSLOT - 1)));
        const uint32_t parity = (n / NSLOT) & 1u;
        unsigned spins = 0;
        while (!mbar_try_wait(bar, parity)) { if (timed_out(spins, 1u << 24)) break; }
      };

      Real carry_row = Real(0), carry_col = Real(0);
      int free_upto = RING;                                 // columns the next warp's ring is known to have room for
      wait_chunk(0);

      // ---- single steps ----
      // kind 0: general step (any column, rounding); 1: edge step (any column of a row that is neither the first nor the last of
      // the grid: weight 1/2, existence by selects); 2: interior column for every row of the warp
      auto do_step = [&](auto kind_tag, const int s) {
        constexpr int KIND = decltype(kind_tag)::value;
        const int sp = s + qoff;
        if ((sp & (K - 1)) == 0 && s > 0) {                // row 0 of the warp enters a new tile (warp-uniform)
          const int j = sp >> LOGK;
          if (j < nchunks) wait_chunk(j);
        }
        const int q = s - hr;
        const int qq = s + lane_q0;                         // column in the tile ring (slot sequence continues across bands)
        const uint32_t slot = ((uint32_t)qq >> LOGK) & (NSLOT - 1);
        const uint32_t kk = ((uint32_t)qq & (K - 1)) ^ (uint32_t)kxor;
        const uint32_t rec = rec_lane + slot * SLOT_REC_B + kk * REC_B;
        const uint32_t par = par_lane + slot * SLOT_PAR_B + kk * PREC_B;
        const bool valid = KIND == 2 || (q >= 0 && q < (int)W);
        const uint32_t rslot = ((uint32_t)(sp - hr) & (uint32_t)(RING - 1)) * LB;

        Real par_ir[P], par_ic[P];
        lds_par<Real, P>(par_ir, par + p_in_row);
        lds_par<Real, P>(par_ic, par + p_in_col);
        Real th = lds1v<Real>(rec), own_ir = lds1v<Real>(rec + o_in_row), own_ic = lds1v<Real>(rec + o_in_col);
        Real own_or = lds1v<Real>(rec + o_out_row), own_oc = lds1v<Real>(rec + o_out_col);

        // ---- the column neighbour's vector: ring slot or the shuffle of the previous step ----
        const bool polling = poller && valid;
        Real polled = Real(0);
        bool missing = polling;
        {                                                   // warp-uniform wait loop
          unsigned spins = 0;
          do {
            lds1_volatile_if(missing, polled, ring_in + rslot);
            missing = polling && is_sentinel(polled);
            if (timed_out(spins, 1u << 27)) { polled = Real(0); missing = false; }
          } while (__any_sync(FULL, missing));
        }
        const Real col_in = polling ? polled : carry_col;
        sts1_volatile_if(polling, ring_in + rslot, sentinel<Real>());   // the slot is free again

        Real out_col;
        if constexpr (KIND >= 1) {
          Real h_ir, h_ic;
          rows_min_plus2<Real, PW, L, WIN>(par_ir, carry_row, h_ir, par_ic, col_in, h_ic);
          const Real m_ir = h_ir + own_ir, m_ic = h_ic + own_ic;
          own_ir = own_ir - m_ir;
          own_ic = own_ic - m_ic;
          th = recv_row_first ? (th + m_ir) + m_ic : (th + m_ic) + m_ir;
          // missing edges: the left column and the first row have all-zero in-side slots and parameters, so their receives are
          // exact no-ops; the last row / column has no out_col / out_row edge (select)
          const Real d = w_fast * th;
          const Real d_oc = out_col_exists ? d : Real(0);
          const Real d_or = KIND == 2 || q + 1 < (int)W ? d : Real(0);
          own_oc = own_oc + d_oc;
          own_or = own_or + d_or;
          th = (th - d_oc) - d_or;
          carry_row = valid ? own_or : Real(0);
          out_col = own_oc;
          const bool on = row_active && valid;
          sts1_if(on, rec, th);
          sts1_if(on, rec + o_in_row, own_ir);
          sts1_if(on, rec + o_in_col, own_ic);
          sts1_if(on, rec + o_out_row, own_or);
          sts1_if(on, rec + o_out_col, own_oc);
        } else {
          const bool act = row_active && valid;
          const bool has_ir = act && q > 0, has_or = act && q + 1 < (int)W;
          const bool has_ic = act && in_col_exists, has_oc = act && out_col_exists;
          const int n_in = (int)has_ir + (int)has_ic, n_out = (int)has_or + (int)has_oc;
          Real pza[P], pzb[P];
#pragma unroll
          for (int p = 0; p < P; ++p) { pza[p] = has_ir ? par_ir[p] : Real(0); pzb[p] = has_ic ? par_ic[p] : Real(0); }
          Real h_ir, h_ic;
          rows_min_plus2<Real, PW, L, WIN>(pza, has_ir ? carry_row : Real(0), h_ir, pzb, has_ic ? col_in : Real(0), h_ic);
          // m = h + own, own -= m   (RepamLeft / RepamRight)
          const Real m_ir = has_ir ? Real(h_ir + own_ir) : Real(0), m_ic = has_ic ? Real(h_ic + own_ic) : Real(0);
          own_ir = has_ir ? Real(own_ir - m_ir) : own_ir;
          own_ic = has_ic ? Real(own_ic - m_ic) : own_ic;
          {
            Real v = th;
            if (recv_row_first) { if (has_ir) v = v + m_ir; if (has_ic) v = v + m_ic; }
            else { if (has_ic) v = v + m_ic; if (has_ir) v = v + m_ir; }
            th = v;
          }
          if (A.primal) {   // first minimiser of theta (include/factors_messages.hxx:2363,2382-2389)
            const Real g = seg_min1<Real, L>(th);
            const unsigned segmask = (L == 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u)) << ((unsigned)lane & ~(unsigned)(L - 1));
            const unsigned first = (unsigned)__ffs((int)(__ballot_sync(FULL, th == g) & segmask)) - 1u;
            if (act && x == 0) A.labels[(uint64_t)r * W + (fwd ? (uint32_t)q : W - 1 - (uint32_t)q)] = first & (unsigned)(L - 1);
          }
          // SendMessages with the snapshot t = theta (weights: include/LP_MP.h:1397)
          const double wd = n_out > 0 ? 1.0 / (double)(n_in > n_out ? n_in : n_out) : 0.0;
          const Real wgt = (Real)wd;
          const Real d = wgt * th;
          Real v = th;
          if (has_oc) { own_oc = own_oc + d; v = v - d; }
          if (has_or) { own_or = own_or + d; v = v - d; }
          th = v;
          carry_row = act ? own_or : Real(0);
          out_col = own_oc;
          sts1_if(act, rec, th);
          sts1_if(act, rec + o_in_row, own_ir);
          sts1_if(act, rec + o_in_col, own_ic);
          sts1_if(act, rec + o_out_row, own_or);
          sts1_if(act, rec + o_out_col, own_oc);
        }

        // ---- hand the freshly written column vector on: next row of the warp (shuffle), next warp (ring), next band (L2) ----
        if (warp_out_ring && s - (RPW - 1) >= free_upto) {  // warp-uniform, rare: the next warp's ring is full
          unsigned spins = 0;
          do {
            free_upto = (int)(lds_volatile_u32(ctr_next) - step_base) + RING;
          } while (s - (RPW - 1) >= free_upto && !timed_out(spins, 1u << 27));
        }
        sts1_volatile_if(out_ring && valid, ring_out + rslot, out_col);
        stg1_relaxed_if(out_glob && valid, out_ptr + (int64_t)q * L, out_col);
        if constexpr (RPW > 1) {
          const Real up = __shfl_up_sync(FULL, out_col, L);
          carry_col = from_shfl ? up : Real(0);
        }
        // ---- publish the step: records (and re-armed ring slots) of this step are in shared memory ----
        // (the tile warp's TMA store reads a completed tile through the async proxy)
        if (((sp - RPW + 2) & (K - 1)) == 0 || s + 1 == nsteps) fence_proxy_async();
        __syncwarp();
        asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0; @p st.volatile.shared.u32 [%1], %2; }" ::"r"((uint32_t)lane), "r"(ctr), "r"(step_base + (uint32_t)(s + 1)) : "memory");
      };

      // ---- steady steps, one tile (8 columns of row 0 of the warp) at a time: every address is a per-tile base + an immediate ----
      // (each instruction of a step is paid W + H times per pass: no index arithmetic, no vote, one publication per tile)
      auto do_tile = [&](const int j) {
        wait_chunk(j);
        const int s0 = (j << LOGK) - qoff;                  // step at which row 0 of the warp enters tile j
        const uint32_t n = cbase + (uint32_t)j;
        const uint32_t slot = n & (NSLOT - 1), slotp = (n - 1u) & (NSLOT - 1);
        // row hr works on column jj - hr of the tile at step jj: for jj < hr that is a column of the PREVIOUS tile
        const uint32_t Tcur = rec_lane + slot * SLOT_REC_B + (uint32_t)(fwd ? -hr : K - 1 + hr) * REC_B;
        const uint32_t Tprev = rec_lane + slotp * SLOT_REC_B + (uint32_t)(fwd ? K - hr : hr - 1) * REC_B;
        const uint32_t Pcur = par_lane + slot * SLOT_PAR_B + (uint32_t)(fwd ? -hr : K - 1 + hr) * PREC_B;
        const uint32_t Pprev = par_lane + slotp * SLOT_PAR_B + (uint32_t)(fwd ? K - hr : hr - 1) * PREC_B;
        const uint32_t c0 = (uint32_t)(j << LOGK) & (uint32_t)(RING - 1);           // ring slot of row 0's column at jj = 0
        const uint32_t Rin = ring_in + c0 * LB;                                      // every lane watches the slot of row 0's column
        const uint32_t Rout = ring_out + (c0 - (uint32_t)(RPW - 1)) * LB;            // the last row of the warp is RPW - 1 columns behind
        const uint32_t Routp = ring_out + (((uint32_t)(j << LOGK) - (uint32_t)(RPW - 1)) & (uint32_t)(RING - 1)) * LB;
        Real* const Gout = out_ptr + (int64_t)(s0 - hr) * L;
        if (warp_out_ring && s0 + K - 1 - (RPW - 1) >= free_upto) {   // room in the next warp's ring for the whole tile (warp-uniform, rare)
          unsigned spins = 0;
          do {
            free_upto = (int)(lds_volatile_u32(ctr_next) - step_base) + RING;
          } while (s0 + K - 1 - (RPW - 1) >= free_upto && !timed_out(spins, 1u << 27));
        }
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          constexpr int SG = fwd ? 1 : -1;
          const uint32_t rec = (jj < RPW - 1 ? (jj < hr ? Tprev : Tcur) : Tcur) + (uint32_t)(SG * jj * (int)REC_B);
          const uint32_t par = (jj < RPW - 1 ? (jj < hr ? Pprev : Pcur) : Pcur) + (uint32_t)(SG * jj * (int)PREC_B);
          const uint32_t rin = Rin + (uint32_t)jj * LB;
          // the polled slot: every lane of the warp reads the words of row 0's column (identical outcome in every lane, no vote)
          Real pv = lds1_volatile<Real>(rin);
          Real par_ir[P], par_ic[P];
          lds_par<Real, P>(par_ir, par + p_in_row);
          lds_par<Real, P>(par_ic, par + p_in_col);
          Real th = lds1v<Real>(rec), own_ir = lds1v<Real>(rec + o_in_row), own_ic = lds1v<Real>(rec + o_in_col);
          Real own_or = lds1v<Real>(rec + o_out_row), own_oc = lds1v<Real>(rec + o_out_col);
          if (is_sentinel(pv)) {
            unsigned spins = 0;
            do { pv = lds1_volatile<Real>(rin); } while (is_sentinel(pv) && ((++spins & 0xffffu) != 0 || !wait_expired(A.control, t_launch)));
            if (is_sentinel(pv)) pv = Real(0);
          }
          sts1_volatile_if(poller, rin, sentinel<Real>());   // the slot is free again
          const Real col_in = poller ? pv : carry_col;

          Real h_ir, h_ic;
          rows_min_plus2<Real, PW, L, WIN>(par_ir, carry_row, h_ir, par_ic, col_in, h_ic);
          const Real m_ir = h_ir + own_ir, m_ic = h_ic + own_ic;
          own_ir = own_ir - m_ir;
          own_ic = own_ic - m_ic;
          th = recv_row_first ? (th + m_ir) + m_ic : (th + m_ic) + m_ir;
          const Real d = w_fast * th;
          const Real d_oc = out_col_exists ? d : Real(0);
          own_oc = own_oc + d_oc;
          own_or = own_or + d;
          th = (th - d_oc) - d;
          carry_row = own_or;
          sts1_if(row_active, rec, th);
          sts1_if(row_active, rec + o_in_row, own_ir);
          sts1_if(row_active, rec + o_in_col, own_ic);
          sts1_if(row_active, rec + o_out_row, own_or);
          sts1_if(row_active, rec + o_out_col, own_oc);
          sts1_volatile_if(out_ring, (jj < RPW - 1 ? Routp : Rout) + (uint32_t)jj * LB, own_oc);
          stg1_relaxed_if(out_glob, Gout + jj * L, own_oc);
          if constexpr (RPW > 1) {
            const Real up = __shfl_up_sync(FULL, own_oc, L);
            carry_col = from_shfl ? up : Real(0);
          }
          if (jj == RPW - 1) {   // the last row of the warp has left the previous tile: publish (once per tile)
            fence_proxy_async();
            __syncwarp();
            asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0; @p st.volatile.shared.u32 [%1], %2; }" ::"r"((uint32_t)lane), "r"(ctr), "r"(step_base + (uint32_t)(s0 + jj + 1)) : "memory");
          }
        }
      };

      // general steps for the whole band: rounding, tiny grids; for the ramps of the warps that own the first / last grid row
      // (corner weights); everything else runs the edge / steady steps
      const bool all_general = A.primal != 0 || (int)W < 2 * RPW + 2 || (H < 2 && !(A.in_halo || A.out_halo));
      const int u_first = (int)kb * T + w * RPW - roff;
      const bool corner_warp = (u_first <= 0 && !A.in_halo) || (u_first + RPW >= (int)H && !A.out_halo);
      const bool poll_row0 = __all_sync(FULL, !poller || hr == 0);   // the polling row of the warp is its row 0 (else: single steps only)
      int s = 0;
      if (all_general) {
        for (; s < nsteps; ++s) do_step(std::integral_constant<int, 0>{}, s);
      } else {
        const int s_steady0 = RPW, s_steady1 = (int)W - 2;  // steady: 1 <= s - hr <= W - 2 for every row of the warp
        if (corner_warp) { for (; s < s_steady0; ++s) do_step(std::integral_constant<int, 0>{}, s); }
        else { for (; s < s_steady0; ++s) do_step(std::integral_constant<int, 1>{}, s); }
        // whole tiles inside the steady range
        const int j0 = (s_steady0 + qoff + K - 1) >> LOGK, j1 = ((s_steady1 + 1 + qoff) >> LOGK) - 1;   // tiles j0 .. j1
        if (poll_row0 && j1 >= j0) {
          for (; s < (j0 << LOGK) - qoff; ++s) do_step(std::integral_constant<int, 2>{}, s);
#pragma unroll 1
          for (int j = j0; j <= j1; ++j) do_tile(j);
          s = ((j1 + 1) << LOGK) - qoff;
        }
#pragma unroll 1
        for (; s <= s_steady1; ++s) do_step(std::integral_constant<int, 2>{}, s);
        if (corner_warp) { for (; s < nsteps; ++s) do_step(std::integral_constant<int, 0>{}, s); }
        else { for (; s < nsteps; ++s) do_step(std::integral_constant<int, 1>{}, s); }
      }
    }
    // ---- end of the band: every ring is drained and re-armed, every tile has left shared memory ----
#pragma unroll
    for (int w = 0; w < NW; ++w) if (warp_live(kb, w)) chunk_base[w] += (uint32_t)nchunks;
    step_base += (uint32_t)nsteps;
    __syncthreads();
  }
  if (wid == NW + 1) tma_wait_all();   // the global writes of the last tiles are complete before the CTA exits
}

}  // namespace lpmp

// Wavefront pass, "lean" variant for the anisotropic (SRMP) passes: the band's COMPUTE warp never touches global memory.
//
// Same schedule, storage, tiles and hand-off protocol as wavefront_pass_kernel (kernels_grid.cuh; replaces
// LP::ComputePass over forward/backwardUpdateOrdering_, /root/reference/include/LP_MP.h:981-1005,1592-1602, and
// FactorContainer::UpdateFactor, include/factors_messages.hxx:2256-2261) -- but a CTA is FOUR warps:
//
//   * warp 0 (compute) runs the node updates of the band out of shared memory only: per step it reads the records of
//     its anti-diagonal fragment from the tile ring, takes the column neighbour's vector of the band's first row from a
//     HALO ring in shared memory (self-validating slots, read one step ahead), writes the records back and publishes a
//     step counter.  No global address, no TMA instruction, no L2 poll on the dependency chain;
//   * warp 1 (in-mover) polls the predecessor band's hand-off slots in L2 -- >= 16 columns per round trip, as deep as the
//     halo ring is free, so loaded L2 latency is covered by parallelism instead of a fixed prefetch distance -- and moves
//     them into the halo ring;
//   * warp 2 (out-mover) claims the band ticket and copies the vectors that leave through the band's last row from the
//     tile ring to the successor as soon as the step counter says they are final: into the successor's halo ring through
//     distributed shared memory when it runs in the next CTA of the thread-block cluster, else into its L2 hand-off slots;
//   * warp 3 (tile warp) issues the TMA tile loads as slots become free and the tile stores as the compute warp completes
//     tiles (its waits for a store to leave shared memory block, so nothing latency-critical shares the warp).
//
// The arithmetic of a node update is the same code in the same order as in wavefront_pass_kernel (bit-identical state).
#pragma once
#include "kernels_grid.cuh"

namespace lpmp {

template <typename Real, int PW, int L, int KSEL>
struct LeanCfg : WavefrontCfg<Real, PW, L, KSEL> {
  using Base = WavefrontCfg<Real, PW, L, KSEL>;
  static constexpr int HALO = Base::T >= 8 ? 32 : 16;   // columns of the halo ring (power of two; short bands follow closely)
  // tile slots hold exactly the T rows of the band (the one-warp kernel reserves a row more for the uniform modes): with
  // L = 32 this is what lets FIVE bands instead of four be resident per SM
  static constexpr int SLOT_REC = Base::T * Base::K * Base::REC;
  static constexpr size_t SMEM_BYTES = 128 + (size_t)(Base::NSLOT * (SLOT_REC + Base::SLOT_PAR) + Base::SCRATCH) * sizeof(Real);
  static constexpr size_t WARP_SMEM = (SMEM_BYTES + 127) / 128 * 128;
  // [per-band region: mbarriers, tile ring, staging | control words (128 B) | halo ring]
  static constexpr size_t LEAN_SMEM = WARP_SMEM + 128 + (size_t)HALO * L * sizeof(Real);
  static constexpr bool LEAN_FITS = Base::FITS && LEAN_SMEM <= 220 * 1024;
  static_assert((SLOT_REC * sizeof(Real)) % 128 == 0, "TMA destination alignment");
};


// ---- thread-block cluster / distributed shared memory primitives ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
template <int BYTES> __device__ __forceinline__ void st_cluster(uint32_t a, const void* src) {
  const uint32_t* d = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
  for (int o = 0; o < BYTES / 16; ++o)
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a + 16 * o), "r"(d[4 * o]), "r"(d[4 * o + 1]), "r"(d[4 * o + 2]), "r"(d[4 * o + 3]) : "memory");
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }

// debug (wavefront_trace = 2): end-of-step stamp of one of the two traced bands; deliberately not inlined
static __device__ __noinline__ void trace_step_stamp(unsigned long long* trace_steps, uint32_t which, int step) {
  if ((threadIdx.x & 31) == 0 && step >= 0 && step < 8192) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    trace_steps[which * 8192u + (uint32_t)step] = now;
  }
}

template <typename Real, int PW, int L, int HOPS, int KSEL>
__global__ void __launch_bounds__(128) wavefront_lean_kernel(const __grid_constant__ CUtensorMap tm_rec, const __grid_constant__ CUtensorMap tm_par,
                                                            const GridArgs<Real> A) {
  using Cfg = LeanCfg<Real, PW, L, KSEL>;
  constexpr bool ANISO = true;
  constexpr int G = Cfg::G, T = Cfg::T, TG = Cfg::TG, P = Cfg::P, K = Cfg::K, LOGK = Cfg::LOGK, REC = Cfg::REC, PREC = Cfg::PREC, NSLOT = Cfg::NSLOT;
  constexpr int HALO = Cfg::HALO;
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  constexpr uint32_t S = (uint32_t)sizeof(Real);
  constexpr uint32_t TILE_BYTES = (uint32_t)((T * K * REC + T * K * PREC) * sizeof(Real));
  constexpr uint32_t REC_B = REC * S, PREC_B = PREC * S, SLOT_REC_B = Cfg::SLOT_REC * S, SLOT_PAR_B = Cfg::SLOT_PAR * S, LB = L * S;
  constexpr bool PARTIAL_WARP = T * G < 32;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bars = smem_base;                                     // NSLOT mbarriers
  const uint32_t rec_ring = smem_base + 128;
  const uint32_t par_ring = rec_ring + NSLOT * SLOT_REC_B;
  const uint32_t scratch_all = par_ring + NSLOT * SLOT_PAR_B;
  const uint32_t ctrl = smem_base + (uint32_t)Cfg::WARP_SMEM;          // +0..+31: super-band queue {index, iteration + 1} x 4, +64: step counter, +68: hand-out counter
  const uint32_t halo = ctrl + 128;                                    // [HALO][L]
  const int lane = threadIdx.x & 31;
  const bool fwd = A.forward != 0;
  const uint32_t H = A.H, W = pin(A.W);
  const int nchunks = ((int)W + K - 1) >> LOGK;
  const int nb = (int)A.num_bands;
  const int qoff = fwd ? 0 : nchunks * K - (int)W;      // padding columns / rows in front of virtual column / row 0
  const int roff = fwd ? 0 : nb * T - (int)H;           // (tiles are anchored at row 0 / column 0 in both directions)
  const int nsteps = (int)W + T - 1;
  const int kxor = fwd ? 0 : K - 1;
  const uint32_t o_in_row = (fwd ? REC_BL : REC_AR) * LB, o_in_col = (fwd ? REC_BU : REC_AD) * LB;
  const uint32_t o_out_row = (fwd ? REC_AR : REC_BL) * LB, o_out_col = (fwd ? REC_AD : REC_BU) * LB;
  const uint32_t p_in_row = (fwd ? PAR_L : PAR_R) * P * S, p_in_col = (fwd ? PAR_U : PAR_D) * P * S;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    Real* hr = reinterpret_cast<Real*>(smem_raw + Cfg::WARP_SMEM + 128);
    for (int i = (int)threadIdx.x; i < HALO * L; i += 128) hr[i] = sentinel<Real>();
    if (threadIdx.x < 32) reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::WARP_SMEM)[threadIdx.x] = 0u;
    if constexpr (PW == LPMP_PAIRWISE_TRUNCLIN) {
      Real* sc = reinterpret_cast<Real*>(smem_raw + 128 + NSLOT * (SLOT_REC_B + SLOT_PAR_B));
      for (int i = (int)threadIdx.x; i < Cfg::SCRATCH; i += 128) sc[i] = RealTraits<Real>::inf();
    }
  }
  // Role of each warp: CTA warp 0 computes, 1 = in-mover, 2 = out-mover, 3 = tile warp.  (Experiment kept behind lean_flags
  // bit 1: rotating the roles by the CTA's arrival order on its SM, so that the compute warps of co-resident CTAs sit on
  // different schedulers, made the passes 10-20 % SLOWER on B200 -- profiles/r01c_lean.txt.)
  if ((A.lean_flags & 2) && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    reinterpret_cast<volatile uint32_t*>(smem_raw + Cfg::WARP_SMEM)[24] = atomicAdd(A.control + 2 + (smid & 1023u), 1u);
  }
  // CTAs of a thread-block cluster work on CONSECUTIVE bands (rank r of the cluster on band sb * csize + r): the vectors
  // that leave band r go straight into the halo ring of rank r + 1 through distributed shared memory; only the boundary
  // between two clusters goes through the L2 hand-off slots
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const unsigned num_super = (A.num_bands + csize - 1u) / csize;
  if (csize > 1) cluster_sync_all();   // every ring of the cluster is armed before a neighbour may write into it
  else __syncthreads();
  const int wid = (int)(((threadIdx.x >> 5) + ((A.lean_flags & 2) ? lds_volatile_u32(ctrl + 96) : 0u)) & 3u);   // 0 compute, 1 in-mover, 2 out-mover, 3 tile warp

  // bounded waits: a lost partner must not hang the device (sticky abort flag, checked by the host after the pass)
  auto timed_out = [&](unsigned& spins, unsigned limit) {
    if ((++spins & 0x3ffu) != 0) return false;
    if (spins > limit) atomicExch(A.control + 1, 1u);
    return *reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u;
  };

  uint32_t chunk_base = 0;   // tiles of earlier bands of this CTA: slot / mbarrier-phase sequence continues
  uint32_t step_base = 0;    // steps of earlier bands: the published step counter is monotone over the kernel
  uint32_t halo_n = 0;       // columns that went through the halo ring in earlier bands: slots are used cyclically
  uint32_t out_base = 0;     // W per earlier band: the out-mover's published hand-out counter is monotone over the kernel

  if (wid == 2) {
    // ========================================== out-mover: ticket, hand-out ==========================================
    const int g = lane / G, gl = lane % G;              // column slot of this lane group inside a batch, label quad
    for (uint32_t iter = 0;; ++iter) {
      unsigned sb = 0;
      if (lane == 0) {
        if (crank == 0) {   // rank 0 claims the super-band and publishes it to every CTA of the cluster
          sb = atomicAdd(A.control, 1u);
          for (uint32_t rk = 0; rk < csize; ++rk)
            asm volatile("st.shared::cluster.v2.u32 [%0], {%1,%2};" ::"r"(mapa_u32(ctrl + 8 * (iter & 3u), rk)), "r"(sb), "r"(iter + 1u) : "memory");
        } else {
          unsigned it1, spins = 0;
          do {
            asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sb), "=r"(it1) : "r"(ctrl + 8 * (iter & 3u)) : "memory");
          } while (it1 != iter + 1u && !timed_out(spins, 1u << 27));
          if (it1 != iter + 1u) sb = 0xffffffffu;
        }
      }
      sb = __shfl_sync(FULL, sb, 0);
      if (sb >= num_super) break;
      const unsigned kb = sb * csize + crank;
      if (kb >= A.num_bands) continue;   // the last super-band may be partial
      const bool last_band = kb + 1 == A.num_bands;
      const bool hand_out = !last_band || (A.out_halo && A.out_boundary != nullptr);
      const bool dsm_out = !last_band && crank + 1 < csize;   // the successor band runs in the next CTA of this cluster
      const uint32_t r_halo = mapa_u32(halo + (uint32_t)(gl * 4) * S, dsm_out ? crank + 1 : crank);
      const uint32_t r_ctr = mapa_u32(ctrl + 64, dsm_out ? crank + 1 : crank);
      const uint32_t col_base = iter * W;   // the successor's running halo column index of this band's column 0 (it polls in every band)
      uint32_t free_upto = col_base;        // running column index up to which the successor's halo slots are known to be free
      const int t_hand = last_band ? (int)H - 1 - (int)kb * T + roff : T - 1;
      Real* const hand_dst = (last_band ? A.out_boundary : A.handoff + (uint64_t)(kb + 1) * W * L) + gl * 4;
      const int row0_own = fwd ? (int)kb * T : (nb - 1 - (int)kb) * T;
      const int trow_hand = fwd ? t_hand : T - 1 - t_hand;
      const uint32_t rec_hand = rec_ring + (uint32_t)(trow_hand * K) * REC_B + o_out_col + (uint32_t)(gl * 4) * S;
      const int n_out = hand_out ? (int)W : 0;
      int q_out = 0;
      unsigned spins = 0;
      if (A.trace != nullptr && lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        A.trace[4ull * kb + 0] = now;
      }
      for (;;) {
        const int c = (int)(lds_volatile_u32(ctrl + 64) - step_base);   // steps of this band the compute warp has completed
        bool moved = false;
        // ---- 1. vectors leaving through the band's last row: tile ring -> successor's hand-off slots ----
        if (q_out < n_out) {
          int avail = c - t_hand;                       // columns the last row has completed
          if (avail > (int)W) avail = (int)W;
          if (avail < 0) avail = 0;
          if (dsm_out) {
            if (q_out < avail && col_base + (uint32_t)q_out >= free_upto) {
              // steps the successor has completed (all bands) -> columns it has consumed and re-armed
              const uint32_t p = ld_cluster_u32(r_ctr), ib = p / (uint32_t)nsteps, cl = p - ib * (uint32_t)nsteps;
              free_upto = ib * W + (cl < W ? cl : W) + (uint32_t)HALO;
            }
            const int lim = (int)(free_upto - col_base);
            if (avail > lim) avail = lim > q_out ? lim : q_out;
          }
          while (q_out < avail) {
            const int q = q_out + g;
            if (q < avail && dsm_out) {
              const int qp = q + qoff;
              const uint32_t slot = (chunk_base + (uint32_t)(qp >> LOGK)) & (NSLOT - 1);
              const uint32_t kk = ((uint32_t)qp & (K - 1)) ^ (uint32_t)kxor;
              Real v[4];
              lds_hand<HAND_BYTES>(v, rec_hand + slot * SLOT_REC_B + kk * REC_B);
              st_cluster<HAND_BYTES>(r_halo + ((col_base + (uint32_t)q) & (uint32_t)(HALO - 1)) * LB, v);
            } else if (q < avail) {
              const int qp = q + qoff;
              const uint32_t slot = (chunk_base + (uint32_t)(qp >> LOGK)) & (NSLOT - 1);
              const uint32_t kk = ((uint32_t)qp & (K - 1)) ^ (uint32_t)kxor;
              Real v[4];
              lds_hand<HAND_BYTES>(v, rec_hand + slot * SLOT_REC_B + kk * REC_B);
              st_relaxed<HAND_BYTES>(hand_dst + (int64_t)q * L, v);
            }
            q_out += TG;
            moved = true;
          }
          if (q_out > avail) q_out = avail;
        }
        if (moved && lane == 0) sts_volatile_u32(ctrl + 68, out_base + (uint32_t)q_out);   // the tile warp stores a tile after its vectors have left
        if (c >= nsteps && q_out >= n_out) break;
        if (moved) spins = 0;
        else if (timed_out(spins, 1u << 26)) break;
      }
      out_base += W;
      if (lane == 0) sts_volatile_u32(ctrl + 68, out_base);
      chunk_base += (uint32_t)nchunks;
      step_base += (uint32_t)nsteps;
    }
  } else if (wid == 3) {
    // ================= tile warp: TMA loads as slots become free, TMA stores as the compute warp completes tiles =================
    // (its waits for a store to leave shared memory block; nothing on the hand-off path waits behind them)
    for (uint32_t iter = 0;; ++iter) {
      unsigned sb = 0;
      if (lane == 0) {
        unsigned it1;
        unsigned spins = 0;
        do {
          asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sb), "=r"(it1) : "r"(ctrl + 8 * (iter & 3u)) : "memory");
          if (it1 != iter + 1u && !(A.lean_flags & 4)) __nanosleep(400);   // idle until the next band: no issue slots taken from the compute warps
        } while (it1 != iter + 1u && !timed_out(spins, 1u << 27));
        if (it1 != iter + 1u) sb = 0xffffffffu;
      }
      sb = __shfl_sync(FULL, sb, 0);
      if (sb >= num_super) break;
      const unsigned kb = sb * csize + crank;
      if (kb >= A.num_bands) continue;
      const bool last_band = kb + 1 == A.num_bands;
      const int n_out = (!last_band || (A.out_halo && A.out_boundary != nullptr)) ? (int)W : 0;
      const int row0_own = fwd ? (int)kb * T : (nb - 1 - (int)kb) * T;
      int j_load = 0, j_store = 0;
      unsigned spins = 0;
      for (;;) {
        const int c = (int)(lds_volatile_u32(ctrl + 64) - step_base);   // steps of this band the compute warp has completed
        const int q_out = (int)(lds_volatile_u32(ctrl + 68) - out_base); // columns of this band the out-mover has handed on
        bool moved = false;
        // ---- 2. tile stores: a tile is final when the band's last group has left it (and its vectors have been handed on) ----
        if (j_store < nchunks) {
          int need = (j_store + 1) * K + T - 1 - qoff;
          if (need > nsteps) need = nsteps;
          int handed = (j_store + 1) * K - qoff;
          if (handed > (int)W) handed = (int)W;
          if (c >= need && (n_out == 0 || q_out >= handed)) {
            if (lane == 0) {
              const uint32_t slot = (chunk_base + (uint32_t)j_store) & (NSLOT - 1);
              const int col0 = fwd ? j_store * K : (nchunks - 1 - j_store) * K;
              tma_store_4d(&tm_rec, col0, row0_own, rec_ring + slot * SLOT_REC_B);
              tma_commit();
            }
            ++j_store;
            moved = true;
          }
        }
        // ---- 3. tile loads: as soon as the slot's previous tile has left shared memory ----
        if (j_load < nchunks && j_load < j_store + NSLOT) {
          if (lane == 0) {
            const int pending_ok = j_store - 1 - (j_load - NSLOT);   // younger stores that may still be reading shared memory
            if (j_load >= NSLOT) {
              if (pending_ok <= 0) tma_wait_read<0>();
              else if (pending_ok == 1) tma_wait_read<1>();
              else if (pending_ok == 2) tma_wait_read<2>();
              else tma_wait_read<3>();
            }
            const uint32_t slot = (chunk_base + (uint32_t)j_load) & (NSLOT - 1);
            const int col0 = fwd ? j_load * K : (nchunks - 1 - j_load) * K;
            const uint32_t bar = bars + 8 * slot;
            mbar_expect_tx(bar, TILE_BYTES);
            tma_load_4d(rec_ring + slot * SLOT_REC_B, &tm_rec, col0, row0_own, bar);
            tma_load_3d(par_ring + slot * SLOT_PAR_B, &tm_par, col0, row0_own, bar);
          }
          ++j_load;
          moved = true;
        }
        if (j_store >= nchunks) break;
        if (moved) spins = 0;
        else { if (!(A.lean_flags & 4)) __nanosleep(200); if (timed_out(spins, 1u << 26)) break; }   // nothing here is latency-critical: leave the issue slots to the compute warps
      }
      if (lane == 0) tma_wait_read<0>();   // every tile of this band has left shared memory before the ring is reused
      __syncwarp();
      chunk_base += (uint32_t)nchunks;
      step_base += (uint32_t)nsteps;
      out_base += W;
    }
    if (lane == 0) tma_wait_all();   // the global writes of the last tiles are complete before the warp exits
  } else if (wid == 1) {
    // ============================== in-mover: predecessor's hand-off slots (L2) -> halo ring ==============================
    constexpr int NBATCH = TG >= 16 ? 1 : 16 / TG;      // >= 16 columns polled per L2 round trip
    const int g = lane / G, gl = lane % G;
    for (uint32_t iter = 0;; ++iter) {
      unsigned sb = 0;
      if (lane == 0) {
        unsigned it1;
        unsigned spins = 0;
        do {
          asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sb), "=r"(it1) : "r"(ctrl + 8 * (iter & 3u)) : "memory");
          if (it1 != iter + 1u && !(A.lean_flags & 4)) __nanosleep(400);   // idle until the next band: no issue slots taken from the compute warps
        } while (it1 != iter + 1u && !timed_out(spins, 1u << 27));
        if (it1 != iter + 1u) sb = 0xffffffffu;
      }
      sb = __shfl_sync(FULL, sb, 0);
      if (sb >= num_super) break;
      const unsigned kb = sb * csize + crank;
      if (kb >= A.num_bands) continue;
      const bool poll_in = kb > 0 || A.in_halo;
      const bool from_l2 = poll_in && crank == 0;   // other ranks: the predecessor's out-mover fills the halo ring directly
      const int t_in = kb == 0 ? roff : 0;
      const Real* const hand_in = (kb == 0 ? A.in_boundary : A.handoff + (uint64_t)kb * W * L) + gl * 4;
      int q_in = 0;
      unsigned spins = 0;
      while (from_l2 && q_in < (int)W) {
        const int c = (int)(lds_volatile_u32(ctrl + 64) - step_base);
        int limit = (c - t_in > 0 ? c - t_in : 0) + HALO;   // the compute warp has consumed (and re-armed) c - t_in columns
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
        bool open = true;   // the halo fills in column order: stop at the first column that has not arrived
#pragma unroll
        for (int b = 0; b < NBATCH; ++b) {
          const int q = q_in + b * TG + g;
          const bool ok = q < limit && !(is_sentinel(v[b][0]) || is_sentinel(v[b][1]) || is_sentinel(v[b][2]) || is_sentinel(v[b][3]));
          const unsigned bad = ~__ballot_sync(FULL, ok);
          const int n_ok = !open ? 0 : bad ? (__ffs((int)bad) - 1) / G : TG;
          if (g < n_ok) {
            sts4(halo + (((halo_n + (uint32_t)q) & (uint32_t)(HALO - 1)) * L + (uint32_t)(gl * 4)) * S, v[b]);
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
      step_base += (uint32_t)nsteps;
      if (poll_in) halo_n += W;
    }
  } else {
  // =================================================== compute warp ===================================================
  const bool lane_used = !PARTIAL_WARP || lane / G < T;
  const int t = lane_used ? lane / G : T - 1, gl = lane % G;
  // two staging areas per lane group: [INF pad L | L labels | INF pad L]
  const uint32_t stage_a = scratch_all + (uint32_t)((2 * (lane / G)) * 3 * L + L + 4 * gl) * S;
  const uint32_t stage_b = stage_a + 3 * LB;
  const int trow = fwd ? t : T - 1 - t;                  // row of this lane group inside the tiles
  const uint32_t rec_lane = rec_ring + (uint32_t)(trow * K) * REC_B + (uint32_t)(gl * 4) * S;
  const uint32_t par_lane = par_ring + (uint32_t)(trow * K) * PREC_B;
  const uint32_t halo_lane = halo + (uint32_t)(gl * 4) * S;
  const bool recv_row_first = fwd ? !A.fifo : (A.fifo != 0);
  const Real w_fast = (Real)(1.0 / 2.0);
  const bool H_total_ge2 = H >= 2 || A.in_halo || A.out_halo;
  const bool potts_nonneg = A.potts_nonneg != 0;
  const int sp_end = qoff + (int)W + T - 1;

  for (uint32_t iter = 0;; ++iter) {
    unsigned sb = 0;
    if (lane == 0) {
      unsigned it1;
      unsigned spins = 0;
      do {
        asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sb), "=r"(it1) : "r"(ctrl + 8 * (iter & 3u)) : "memory");
      } while (it1 != iter + 1u && !timed_out(spins, 1u << 27));
      if (it1 != iter + 1u) sb = 0xffffffffu;
    }
    sb = __shfl_sync(FULL, sb, 0);
    if (sb >= num_super) break;
    const unsigned kb = sb * csize + crank;
    if (kb >= A.num_bands) continue;
    const int u = (int)kb * T + t - roff;                  // virtual row of this lane group
    const bool row_active = u >= 0 && u < (int)H && lane_used;
    const uint32_t r = fwd ? (uint32_t)u : H - 1 - (uint32_t)u;
    const bool in_col_exists = u > 0 || A.in_halo, out_col_exists = u + 1 < (int)H || A.out_halo;
    const bool last_band = kb + 1 == A.num_bands;
    const bool poll_in = kb > 0 || A.in_halo;              // group t_in takes its column neighbour from the halo ring
    const int t_in = kb == 0 ? roff : 0;                   // first active group of the band
    const bool poller = poll_in && t == t_in && lane_used;
    const int lane_q0 = (int)(chunk_base << LOGK) - t;     // tile-ring column of this lane group at sp = 0
    auto halo_slot = [&](int q) { return halo_lane + ((halo_n + (uint32_t)q) & (uint32_t)(HALO - 1)) * LB; };
    auto wait_chunk = [&](int j) {
      const uint32_t n = chunk_base + (uint32_t)j;
      const uint32_t bar = bars + 8 * (n & (NSLOT - 1));
      const uint32_t parity = (n / NSLOT) & 1u;
      unsigned spins = 0;
      while (!mbar_try_wait(bar, parity)) { if (timed_out(spins, 1u << 24)) break; }
    };
    auto stamp = [&](int which) {
      if (A.trace != nullptr && lane == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        A.trace[4ull * kb + which] = now;
      }
    };

    const bool trace_band = A.trace_steps != nullptr && kb - A.trace_band < 2u;
    Real polled[4];   // the halo slot of the NEXT column of group t_in, read one step ahead
#pragma unroll
    for (int k = 0; k < 4; ++k) polled[k] = Real(0);
    lds_hand_if<HAND_BYTES>(poller, polled, halo_slot(0));
    Real carry_row[4], carry_col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { carry_row[k] = Real(0); carry_col[k] = Real(0); }

    wait_chunk(0);
    stamp(1);

    // kind 0: general step; 1: fast step (interior columns, anisotropic, no rounding); 2: STEADY fast step -- in addition every
    // group has entered the grid, every row of the band exists and has a row below, and group 0 takes its column neighbour
    // from the halo ring: no existence selects at all (each instruction of this step is paid (H + W) times per pass)
    auto do_step = [&](auto kind_tag, const int sp) {
      constexpr int KIND = decltype(kind_tag)::value;
      constexpr bool FAST = KIND >= 1, STEADY = KIND == 2;
      if ((sp & (K - 1)) == 0 && sp > 0) {               // group 0 enters a new tile (warp-uniform)
        if (sp == K) stamp(2);
        const int j = sp >> LOGK;
        if (j < nchunks) wait_chunk(j);
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

      // ---- the column neighbour's vector: shuffle inside the band, halo ring for group t_in ----
      Real col_in[4];
      // (fast steps never reach the last column, so only the left ramp needs the range test there)
      const bool polling = STEADY ? poller : FAST ? poller && q >= 0 : poller && q >= 0 && q < (int)W;
      // (the wait is a warp-UNIFORM loop: lanes parked at a reconvergence point for a long wait wake up slowly -- measured
      // 1.4 us on the first step of every band, which then stays in the band-to-band lag for the whole pass)
      // (steady steps: only the polling lanes ever hold a sentinel -- the others keep their initial zeros -- so no lane test)
      bool missing = (STEADY || polling) && (is_sentinel(polled[0]) || is_sentinel(polled[1]) || is_sentinel(polled[2]) || is_sentinel(polled[3]));
      if (__any_sync(FULL, missing)) {
        unsigned spins = 0;
        do {
          lds_hand_if<HAND_BYTES>(missing, polled, halo_slot(q));
          missing = polling && (is_sentinel(polled[0]) || is_sentinel(polled[1]) || is_sentinel(polled[2]) || is_sentinel(polled[3]));
          if (timed_out(spins, 1u << 27)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) polled[k] = Real(0);
            missing = false;
          }
        } while (__any_sync(FULL, missing));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) col_in[k] = polling ? polled[k] : carry_col[k];
      {
        Real sent[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) sent[k] = sentinel<Real>();
        sts_hand_if<HAND_BYTES>(polling, halo_slot(q), sent);                         // the slot is free again
        lds_hand_if<HAND_BYTES>(FAST ? polling : polling && q + 1 < (int)W, polled, halo_slot(q + 1));
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
        const bool started = STEADY || q >= 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real d = w_fast * th[k];   // every send uses the same snapshot value
          const Real d_oc = STEADY || out_col_exists ? d : Real(0);
          own_oc[k] = own_oc[k] + d_oc;
          own_or[k] = own_or[k] + d;
          th[k] = (th[k] - d_oc) - d;
          carry_row[k] = STEADY || started ? own_or[k] : Real(0);
          out_col[k] = own_oc[k];
        }
        {   // predicated stores: no branch in the step
          const bool on = STEADY || (row_active && started);
          sts_hand_if<HAND_BYTES>(on, rec, th);
          sts_hand_if<HAND_BYTES>(on, rec + o_in_row, own_ir);
          sts_hand_if<HAND_BYTES>(on, rec + o_in_col, own_ic);
          sts_hand_if<HAND_BYTES>(on, rec + o_out_row, own_or);
          sts_hand_if<HAND_BYTES>(on, rec + o_out_col, own_oc);
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
        {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            Real x = th[k];
            if (recv_row_first) { if (has_ir) x = x + m_ir[k]; if (has_ic) x = x + m_ic[k]; }
            else { if (has_ic) x = x + m_ic[k]; if (has_ir) x = x + m_ir[k]; }
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

      // ---- hand the freshly written column vector to the node below (next lane group; the mover takes the last row's) ----
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const Real up = __shfl_up_sync(FULL, out_col[k], G);
        carry_col[k] = !STEADY && t <= t_in ? Real(0) : up;   // (steady: group 0 = t_in reads the halo ring, its carry is unused)
      }
      // ---- publish the step: records (and re-armed halo slots) of this step are in shared memory ----
      // (predicated, no branch: this step completes a tile, the out-mover's TMA store reads it through the async proxy)
      asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p fence.proxy.async.shared::cta; }" ::"r"((uint32_t)(((sp - T + 2) & (K - 1)) == 0 || (!STEADY && sp + 1 == sp_end))) : "memory");
      __syncwarp();
      asm volatile("{ .reg .pred p; setp.eq.u32 p, %0, 0; @p st.volatile.shared.u32 [%1], %2; }" ::"r"((uint32_t)lane), "r"(ctrl + 64), "r"(step_base + (uint32_t)(sp - qoff + 1)) : "memory");
      if (trace_band) trace_step_stamp(A.trace_steps, kb - A.trace_band, sp - qoff);   // debug; a call, so that nothing of it is if-converted into the step
    };

    const bool fast_band = !A.primal && W >= 3 && H_total_ge2;
    // steady steps: interior bands (a predecessor and a successor band, all T rows present) once the left ramp is over
    const bool steady_band = fast_band && !PARTIAL_WARP && kb > 0 && !last_band && (int)kb * T - roff >= 0 && (int)kb * T - roff + T <= (int)H;
    const int sp_fast0 = fast_band ? (last_band ? qoff + T : qoff) : sp_end, sp_fast1 = fast_band ? qoff + (int)W - 2 : sp_end;
    // three loops, so that the fast steps form one compact loop body (a lone warp pays every far jump with a fetch stall)
    // A band cannot run faster than its predecessor, so whatever its FIRST step loses (the loop body has never been
    // fetched on this SM: measured 1.4 us) stays in the band-to-band lag for the whole pass.  One dry step at "column -1"
    // (all groups left of the grid: nothing is stored, carried or published) warms the instruction cache while the
    // band is waiting for its predecessor anyway -- it runs as iteration -1 of the fast loop itself (same code addresses).
    int sp = qoff;
    for (; sp < sp_fast0 && sp < sp_end; ++sp) do_step(std::integral_constant<int, 0>{}, sp);
    if (fast_band && sp == qoff) sp = qoff - 1;
#pragma unroll 1
    for (; sp <= sp_fast1 && sp < sp_end && (!steady_band || sp < qoff + T - 1); ++sp) do_step(std::integral_constant<int, 1>{}, sp);
#pragma unroll 1
    for (; steady_band && sp <= sp_fast1 && sp < sp_end; ++sp) do_step(std::integral_constant<int, 2>{}, sp);
    for (; sp < sp_end; ++sp) do_step(std::integral_constant<int, 0>{}, sp);
    stamp(3);
    chunk_base += (uint32_t)nchunks;
    step_base += (uint32_t)(W + T - 1);
    if (poll_in) halo_n += W;
  }
  }
  // no CTA of the cluster leaves while a neighbour may still read its step counter or write into its halo ring
  if (csize > 1) cluster_sync_all();
}

}  // namespace lpmp

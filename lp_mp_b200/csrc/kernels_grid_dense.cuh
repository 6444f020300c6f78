// Row-dataflow pass over a 4-connected grid MRF with DENSE L x L pairwise tables (BASELINE config 5): one persistent
// kernel per pass (sm_100a).
//
// Same contract as kernels_grid.cuh: replaces LP::ComputePass(it...) / ComputePassAndPrimal
// (/root/reference/include/LP_MP.h:981-1005,1592-1602) + FactorContainer::UpdateFactor
// (include/factors_messages.hxx:2256-2261) for the anisotropic (SRMP) reparametrisation, bit-identical to the
// sequential order.  The min-marginals are the dense ones of PairwiseSimplexFactor
// (matrix::min1(v)/min2(v), include/vector.hxx:850-859,889-913): h(x) = min_y [c(x,y) + other(y)].
//
// Why a different shape than the Potts / truncated-linear kernel: a node update streams two tables (2*L*L reals,
// 32 KiB at L = 64 in f32) against 5*L reals of state, so a node costs ~0.3 us of min-plus work and the sweep is
// HBM-bound by the tables, not by the dependency chain.  Hence
//   * a CTA (4 warps) owns one grid ROW at a time (rows are claimed in order from an atomic ticket) and walks it node
//     by node; rows run concurrently, each a few nodes behind the row above -- the anti-diagonal wavefront emerges
//     from the data dependencies instead of being scheduled;
//   * tables and node records stream HBM -> shared memory through a ring of NSTAGE node stages filled by 1-D bulk
//     copies (cp.async.bulk, mbarrier completion), issued NSTAGE-1 nodes ahead by one elected thread; every table is
//     read exactly once per pass, every record read and written once (bulk store from shared memory);
//   * per node, warps 0-1 min-marginalise the row edge and warps 2-3 the column edge (64 threads per table; lanes
//     along the contiguous table dimension so that shared-memory reads are conflict-free: column mode when the node
//     is the second variable (forward pass), rotated row mode when it is the first variable (backward pass)), then
//     L/4 lanes of warp 0 apply ReceiveMessages / SendMessages on the node record in place;
//   * the vector for the node below travels through a hand-off buffer in L2 with self-validating slots (NaN sentinel,
//     same protocol as kernels_grid.cuh); the consumer row requests slots two nodes ahead;
//   * the BACKWARD pass also produces LP::LowerBound at no extra traffic: node i is the first variable of its R and D
//     edges, whose state is final once i has been updated, and h(x) = min_y [c(x,y) + b(y)] is already in registers, so
//     min theta_i + sum_{e in {R,D}} min_x [a_e(x) + h_e(x)] is accumulated per row in double (fixed order: deterministic).
#pragma once
#include "kernels_grid.cuh"

namespace lpmp {

template <typename Real>
struct DenseGridArgs {
  uint32_t H, W;
  int forward, primal, fifo;
  int in_halo, out_halo;     // striped: the row before virtual row 0 / after the last row lives on another device
  unsigned int* control;     // [0] row ticket (zeroed before the launch), [1] sticky abort flag (poll timeout)
  Real* rec;                 // [H][W][5][L] node records (theta, aR, aD, bL, bU)
  const Real* tabH;          // [H][W][L*L]: edge (r,c)-(r,c+1) at r*W+c, x_(r,c) major
  const Real* tabV;          // [H+1][W][L*L]: edge (r-1,c)-(r,c) at r*W+c (row 0 / row H: stripe-boundary edges)
  Real* handoff;             // [H][W][L]: slot (u,q) = vector written by virtual node (u,q) for (u+1,q); all-sentinel between passes
  Real* in_boundary;         // [W][L] slots feeding virtual row 0 when in_halo
  Real* out_boundary;        // [W][L] where the last virtual row delivers when out_halo, or null
  uint32_t* labels;          // [H*W]
  double* lb_rows;           // backward pass only, or null: [H] row sums of LP::LowerBound's terms (include/LP_MP.h:1507-1518), see below
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts1(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

template <typename Real, int L>
struct DenseCfg {
  static constexpr int NT = 128;                          // threads per CTA
  static constexpr int TPE = 64;                          // threads per edge table
  static constexpr int TAB = L * L;                       // reals per table
  static constexpr int REC = REC_VECS * L;
  static constexpr int STAGE = 2 * TAB + REC;             // [table of the row edge | table of the column edge | node record]
  static constexpr int NSTAGE = (size_t)3 * STAGE * sizeof(Real) > 200 * 1024 ? 2 : ((size_t)4 * STAGE * sizeof(Real) > 110 * 1024 ? 3 : 4);
  static constexpr int YQ = L / 4;                        // lanes across a table row (4 labels each)
  static constexpr int NXS = TPE / YQ;                    // column mode: slices of the minimised (row) index
  static constexpr int XPS = L / NXS;                     // ... rows per slice
  static constexpr int RS = TPE / L > 0 ? TPE / L : 1;    // row mode: threads per table row
  static constexpr int NJ = (L / RS) / 4;                 // ... 4-label chunks per thread
  static constexpr int NPART = NXS > RS ? NXS : RS;
  static constexpr int MISC = L /*carry*/ + 2 * L /*colin*/ + 2 * NPART * L /*hpart*/;
  static constexpr size_t SMEM_BYTES = 128 + (size_t)(NSTAGE * STAGE + MISC) * sizeof(Real);
  static constexpr bool FITS = L >= 16 && L <= 64 && (L & (L - 1)) == 0 && SMEM_BYTES <= 220 * 1024;
};

template <typename Real, int L>
__global__ void __launch_bounds__(128) dense_row_pass_kernel(const DenseGridArgs<Real> A) {
  using Cfg = DenseCfg<Real, L>;
  constexpr int TAB = Cfg::TAB, REC = Cfg::REC, STAGE = Cfg::STAGE, NSTAGE = Cfg::NSTAGE, YQ = Cfg::YQ, NXS = Cfg::NXS, XPS = Cfg::XPS, RS = Cfg::RS,
                NJ = Cfg::NJ, NPART = Cfg::NPART;
  constexpr uint32_t S = (uint32_t)sizeof(Real), LB = L * S, TAB_B = TAB * S, REC_B = REC * S, STAGE_B = STAGE * S;
  constexpr int HAND_BYTES = 4 * (int)sizeof(Real);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bars = smem_base;                                   // NSTAGE mbarriers, then the row ticket at +64
  const uint32_t stages = smem_base + 128;
  const uint32_t carry = stages + NSTAGE * STAGE_B;                  // out_row vector of the previous node of this row
  const uint32_t colin = carry + LB;                                 // [2][L] vector from the row above (double buffered)
  const uint32_t hpart = colin + 2 * LB;                             // [2 edges][NPART][L] partial min-plus results
  volatile unsigned int* const ticket_s = reinterpret_cast<volatile unsigned int*>(smem_raw + 64);
  const int tid = threadIdx.x, lane = tid & 31;
  const bool fwd = A.forward != 0;
  const uint32_t H = A.H, W = A.W;
  const Real INF = RealTraits<Real>::inf();

  const bool is_loader = tid == 96;                                  // issues every bulk copy of this CTA
  const bool is_poller = tid >= 64 && tid < 64 + YQ;                 // fetch the hand-off vector, 4 labels each
  const bool is_updater = tid < YQ;                                  // node update, 4 labels each
  const int e_id = tid >> 6, et = tid & 63;                          // phase A: edge (0 = row edge, 1 = column edge), thread within edge

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // record vector roles in this direction (byte offsets inside a record)
  const uint32_t o_in_row = (fwd ? REC_BL : REC_AR) * LB, o_in_col = (fwd ? REC_BU : REC_AD) * LB;
  const uint32_t o_out_row = (fwd ? REC_AR : REC_BL) * LB, o_out_col = (fwd ? REC_AD : REC_BU) * LB;
  const bool recv_row_first = fwd ? !A.fifo : (A.fifo != 0);         // LIFO: (L, U) forward, (D, R) backward
  uint32_t node_base = 0;                                            // nodes staged by earlier rows of this CTA (stage / phase sequence)

  for (;;) {
    if (tid == 0) *ticket_s = atomicAdd(A.control, 1u);
    __syncthreads();
    const uint32_t u = *ticket_s;
    __syncthreads();
    if (u >= H) break;
    const uint32_t r = fwd ? u : H - 1 - u;
    const bool has_ic = u > 0 || A.in_halo, has_oc = u + 1 < H || A.out_halo;
    const bool deliver = u + 1 < H || (A.out_halo && A.out_boundary != nullptr);
    const Real* const hand_in = (u == 0 ? A.in_boundary : A.handoff + (uint64_t)(u - 1) * W * L) + 4 * (tid & (YQ - 1));   // pollers: tid - 64
    Real* const hand_dst = (u + 1 < H ? A.handoff + (uint64_t)u * W * L : A.out_boundary) + 4 * (tid & (YQ - 1));          // updaters: tid
    Real* const rec_row = A.rec + (uint64_t)r * W * REC;
    // tables of virtual node q (actual column c): forward (r,c-1)-(r,c) and (r-1,c)-(r,c); backward (r,c)-(r,c+1) and (r,c)-(r+1,c)
    const Real* const tab_row_base = A.tabH + (uint64_t)r * W * TAB;
    const Real* const tab_col_base = A.tabV + (uint64_t)(fwd ? r : r + 1) * W * TAB;

    auto issue = [&](uint32_t q) {   // loader only
      const uint32_t c = fwd ? q : W - 1 - q;
      const uint32_t n = node_base + q, st = n % NSTAGE;
      const uint32_t bar = bars + 8 * st, dst = stages + st * STAGE_B;
      const bool ir = q > 0;
      mbar_expect_tx(bar, REC_B + (ir ? TAB_B : 0u) + (has_ic ? TAB_B : 0u));
      bulk_load(dst + 2 * TAB_B, rec_row + (uint64_t)c * REC, REC_B, bar);
      if (ir) bulk_load(dst, tab_row_base + (uint64_t)(fwd ? c - 1 : c) * TAB, TAB_B, bar);
      if (has_ic) bulk_load(dst + TAB_B, tab_col_base + (uint64_t)c * TAB, TAB_B, bar);
    };
    auto poll_valid = [&](Real (&v)[4]) { return !(is_sentinel(v[0]) || is_sentinel(v[1]) || is_sentinel(v[2]) || is_sentinel(v[3])); };
    auto poll_wait = [&](Real (&v)[4], uint32_t q) {   // pollers: spin until slot q holds a vector; returns it in v and re-arms the slot
      unsigned spins = 0;
      while (!poll_valid(v)) {
        ld_relaxed<HAND_BYTES>(v, hand_in + (uint64_t)q * L);
        if ((++spins & 0x3ffu) == 0) {      // bounded wait: a lost producer must not hang the device
          if (spins > (1u << 22)) atomicExch(A.control + 1, 1u);
          if (*reinterpret_cast<volatile unsigned int*>(A.control + 1) != 0u) { v[0] = v[1] = v[2] = v[3] = Real(0); }
        }
      }
      Real sent[4] = {sentinel<Real>(), sentinel<Real>(), sentinel<Real>(), sentinel<Real>()};
      st_relaxed<HAND_BYTES>(const_cast<Real*>(hand_in) + (uint64_t)q * L, sent);
    };

    if (is_loader) {
      tma_wait_read<0>();   // the last record stores of the previous row have left shared memory
      for (uint32_t q = 0; q + 1 < (uint32_t)NSTAGE && q < W; ++q) issue(q);
    }
    Real pend0[4], pend1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { pend0[k] = Real(0); pend1[k] = Real(0); }
    if (is_poller && has_ic) {
      Real v[4] = {sentinel<Real>(), sentinel<Real>(), sentinel<Real>(), sentinel<Real>()};
      if (W > 1) ld_relaxed<HAND_BYTES>(pend0, hand_in + (uint64_t)1 * L);
      if (W > 2) ld_relaxed<HAND_BYTES>(pend1, hand_in + (uint64_t)2 * L);
      poll_wait(v, 0);
      sts4(colin + (uint32_t)(4 * (tid - 64)) * S, v);
    }
    __syncthreads();

    double row_lb = 0.0;
    for (uint32_t q = 0; q < W; ++q) {
      const uint32_t c = fwd ? q : W - 1 - q;
      const uint32_t n = node_base + q, st = n % NSTAGE;
      const uint32_t stage = stages + st * STAGE_B;
      const bool has_ir = q > 0, has_or = q + 1 < W;
      {
        const uint32_t bar = bars + 8 * st, parity = (n / NSTAGE) & 1u;
        while (!mbar_try_wait(bar, parity)) {}
      }
      // ---- phase A: min-plus of the two incoming edges, 64 threads per table -------------------------------------------
      if (e_id == 0 ? has_ir : has_ic) {
        const uint32_t tab = stage + (uint32_t)e_id * TAB_B;
        const uint32_t oth = e_id == 0 ? carry : colin + (q & 1u) * LB;
        const uint32_t hp = hpart + (uint32_t)e_id * NPART * LB;
        if (fwd) {
          // node = second variable: h(y) = min_x [c(x,y) + oth(x)]; lanes along y (4 labels each), slices of x
          const int yq = et % YQ, xs = et / YQ;
          Real acc[4] = {INF, INF, INF, INF};
#pragma unroll 8
          for (int i = 0; i < XPS; ++i) {
            const int x = xs * XPS + i;
            Real cv[4];
            lds4(cv, tab + (uint32_t)(x * L + 4 * yq) * S);
            const Real o = lds1<Real>(oth + (uint32_t)x * S);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = rmin(acc[k], Real(cv[k] + o));
          }
          sts4(hp + (uint32_t)(xs * L + 4 * yq) * S, acc);
        } else {
          // node = first variable: h(x) = min_y [c(x,y) + oth(y)]; one table row per thread group, chunks rotated by the
          // thread index so that the 128-bit shared-memory reads of a quarter warp hit distinct banks
          const int x = et % L, hs = et / L;
          if (hs < RS) {
            Real acc = INF;
#pragma unroll 8
            for (int j = 0; j < NJ; ++j) {
              const int y0 = hs * (L / RS) + 4 * ((j + et) % NJ);
              Real cv[4], ov[4];
              lds4(cv, tab + (uint32_t)(x * L + y0) * S);
              lds4(ov, oth + (uint32_t)y0 * S);
              acc = rmin(acc, rmin(rmin(Real(cv[0] + ov[0]), Real(cv[1] + ov[1])), rmin(Real(cv[2] + ov[2]), Real(cv[3] + ov[3]))));
            }
            sts1(hp + (uint32_t)(hs * L + x) * S, acc);
          }
        }
      }
      __syncthreads();

      // ---- phase B: node update (warp 0), hand-off fetch for node q+1 (warp 2), staging of node q+NSTAGE-1 (loader) ---
      if (is_updater) {
        const uint32_t recs = stage + 2 * TAB_B + (uint32_t)(4 * tid) * S;
        Real th[4], own_ir[4], own_ic[4], own_or[4], own_oc[4], m_ir[4], m_ic[4];
        lds4(th, recs);
        lds4(own_ir, recs + o_in_row);
        lds4(own_ic, recs + o_in_col);
        lds4(own_or, recs + o_out_row);
        lds4(own_oc, recs + o_out_col);
        const int np = fwd ? NXS : RS;
#pragma unroll
        for (int k = 0; k < 4; ++k) { m_ir[k] = Real(0); m_ic[k] = Real(0); }
        Real h_ir[4] = {INF, INF, INF, INF}, h_ic[4] = {INF, INF, INF, INF};
        if (has_ir) {   // ReceiveMessages: m = min-marginal, theta += m (RepamLeft), own -= m (RepamRight)
          for (int p = 0; p < np; ++p) {
            Real v[4];
            lds4(v, hpart + (uint32_t)(p * L + 4 * tid) * S);
#pragma unroll
            for (int k = 0; k < 4; ++k) h_ir[k] = rmin(h_ir[k], v[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) { m_ir[k] = h_ir[k] + own_ir[k]; own_ir[k] = own_ir[k] - m_ir[k]; }
        }
        if (has_ic) {
          for (int p = 0; p < np; ++p) {
            Real v[4];
            lds4(v, hpart + NPART * LB + (uint32_t)(p * L + 4 * tid) * S);
#pragma unroll
            for (int k = 0; k < 4; ++k) h_ic[k] = rmin(h_ic[k], v[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) { m_ic[k] = h_ic[k] + own_ic[k]; own_ic[k] = own_ic[k] - m_ic[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          Real x = th[k];
          if (recv_row_first) { if (has_ir) x = x + m_ir[k]; if (has_ic) x = x + m_ic[k]; }
          else { if (has_ic) x = x + m_ic[k]; if (has_ir) x = x + m_ir[k]; }
          th[k] = x;
        }
        if (A.primal) {   // MaximizePotentialAndComputePrimal: first minimum
          Real bv = th[0];
          uint32_t bi = (uint32_t)(tid * 4);
#pragma unroll
          for (int k = 1; k < 4; ++k) if (th[k] < bv) { bv = th[k]; bi = (uint32_t)(tid * 4 + k); }
          group_argmin(bv, bi, YQ, (YQ == 32 ? 0xffffffffu : ((1u << YQ) - 1u)));
          if (tid == 0) A.labels[(uint64_t)r * W + c] = bi;
        }
        // SendMessages with the snapshot t = theta; weight 1/max(n_in, n_out) (include/LP_MP.h:1397)
        const int n_in = (int)has_ir + (int)has_ic, n_out = (int)has_or + (int)has_oc;
        const Real w = (Real)(n_out > 0 ? 1.0 / (double)(n_in > n_out ? n_in : n_out) : 0.0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Real d = w * th[k];
          Real x = th[k];
          if (has_oc) { own_oc[k] = own_oc[k] + d; x = x - d; }
          if (has_or) { own_or[k] = own_or[k] + d; x = x - d; }
          th[k] = x;
        }
        if (A.lb_rows != nullptr) {   // backward: in_row / in_col are the R / D edges of this node, now final
          Real mT = rmin(rmin(th[0], th[1]), rmin(th[2], th[3]));
          Real mR = rmin(rmin(Real(h_ir[0] + own_ir[0]), Real(h_ir[1] + own_ir[1])), rmin(Real(h_ir[2] + own_ir[2]), Real(h_ir[3] + own_ir[3])));
          Real mD = rmin(rmin(Real(h_ic[0] + own_ic[0]), Real(h_ic[1] + own_ic[1])), rmin(Real(h_ic[2] + own_ic[2]), Real(h_ic[3] + own_ic[3])));
          constexpr unsigned UM = YQ == 32 ? 0xffffffffu : ((1u << YQ) - 1u);
#pragma unroll
          for (int o = YQ >> 1; o > 0; o >>= 1) {
            mT = rmin(mT, __shfl_xor_sync(UM, mT, o));
            mR = rmin(mR, __shfl_xor_sync(UM, mR, o));
            mD = rmin(mD, __shfl_xor_sync(UM, mD, o));
          }
          row_lb += (double)mT + (has_ir ? (double)mR : 0.0) + (has_ic ? (double)mD : 0.0);
        }
        sts4(recs, th);
        sts4(recs + o_in_row, own_ir);
        sts4(recs + o_in_col, own_ic);
        sts4(recs + o_out_row, own_or);
        sts4(recs + o_out_col, own_oc);
        sts4(carry + (uint32_t)(4 * tid) * S, own_or);
        if (has_oc && deliver) st_relaxed<HAND_BYTES>(hand_dst + (uint64_t)q * L, own_oc);
        fence_proxy_async();   // the record is stored from shared memory by the bulk copy engine
      } else if (is_poller) {
        if (has_ic && q + 1 < W) {
          poll_wait(pend0, q + 1);
          sts4(colin + ((q + 1) & 1u) * LB + (uint32_t)(4 * (tid - 64)) * S, pend0);
#pragma unroll
          for (int k = 0; k < 4; ++k) pend0[k] = pend1[k];
          if (q + 3 < W) ld_relaxed<HAND_BYTES>(pend1, hand_in + (uint64_t)(q + 3) * L);
        }
      } else if (is_loader) {
        if (q + NSTAGE - 1 < W) {
          tma_wait_read<0>();   // the record store of node q-1 has left the stage that node q+NSTAGE-1 reuses
          issue(q + NSTAGE - 1);
        }
      }
      __syncthreads();
      if (is_loader) {
        bulk_store(rec_row + (uint64_t)c * REC, stage + 2 * TAB_B, REC_B);
        tma_commit();
      }
    }
    if (A.lb_rows != nullptr && tid == 0) A.lb_rows[r] = row_lb;
    node_base += W;
  }
  if (is_loader) tma_wait_all();   // the global writes of the last records are complete before the CTA exits
}

}  // namespace lpmp

// Anti-diagonal wavefront pass over a 4-connected grid MRF: one persistent kernel per pass.
//
// Replaces LP::ComputePass(it...) / ComputePassAndPrimal over forwardUpdateOrdering_ / backwardUpdateOrdering_
// (/root/reference/include/LP_MP.h:981-1005,1592-1602) and the per-node FactorContainer::UpdateFactor
// (include/factors_messages.hxx:2256-2261).  By SURVEY.md App. B every linear extension of the relation DAG
// u_i -> p_e -> u_j gives bit-identical state, so the sequential sweep is re-scheduled as a pipeline:
//
//   * a WARP owns a band of T = 32/G consecutive rows (G = L/VPL lanes per node, each lane holds VPL labels) and
//     sweeps it column by column with the rows skewed by one column: at step s lane-group t updates node
//     (row t of the band, column s - t), i.e. the warp always works on an anti-diagonal fragment;
//   * the row neighbour's message vector stays in the registers of the same lanes from one step to the next, the
//     column neighbour's vector moves one lane-group up by a warp shuffle -- no memory round trip inside a band;
//   * bands are claimed in order from an atomic ticket; band k waits on band k-1's progress counter (number of
//     columns its last row has finished) and reads that row's outgoing vectors from L2 (ld.global.cg);
//   * everything a step needs from HBM is known in advance and is prefetched STAGES steps ahead with cp.async
//     into a per-lane shared-memory ring, so HBM latency is off the dependency chain.
//
// "Virtual" coordinates (u, q) run in processing order: forward pass u = r, q = c; backward pass u = H-1-r,
// q = W-1-c.  Edges of node (r,c):  R right  slot h  = r*W+c      (node is first variable,  own a_h, other b_h)
//                                   D down   slot v  = r*W+c      (first variable,          own a_v, other b_v)
//                                   Lf left  slot h' = r*W+c-1    (second variable,         own b_h, other a_h)
//                                   U up     slot v' = (r-1)*W+c  (second variable,         own b_v, other a_v)
// Message order of a node (include/factors_messages.hxx:3402-3419, 2030-2041): first-variable edges, then
// second-variable edges; LIFO inside a type => D, R, Lf, U (FIFO: R, D, U, Lf).
#pragma once
#include <cuda_pipeline_primitives.h>

#include "device_common.cuh"

namespace lpmp {

template <typename Real>
struct GridArgs {
  Real* theta;       // [H*W][L]
  Real* ah; Real* bh;  // horizontal edges, slot r*W+c: a on (r,c), b on (r,c+1)
  Real* av; Real* bv;  // vertical edges,   slot r*W+c: a on (r,c), b on (r+1,c)
  const Real* ph; const Real* pv;  // parameters [H*W][P]
  uint32_t* labels;
  uint32_t H, W;
  uint32_t row0, H_total;   // stripe placement inside the full grid (row0 = 0, H_total = H when not striped)
  const Real* halo_top;     // [W][L] other side of the edges above row 0 of this stripe (a_v of the stripe above) or null
  const Real* halo_bot;     // [W][L] other side of the edges below the last row (b_v of the stripe below) or null
  Real* own_top;            // [W][L] this stripe's side (b_v) of the edges above row 0, or null
  Real* own_bot;            // [W][L] this stripe's side (a_v) of the edges below the last row, or null
  const Real* p_top; const Real* p_bot;  // [W][P] parameters of the stripe-boundary edges
  unsigned int* ticket;     // band ticket counter (zeroed before the launch)
  unsigned int* progress;   // [num_bands] columns finished by the last row of each band (zeroed before the launch)
  uint32_t num_bands;
  int forward, primal, fifo, repam;
};

template <typename Real, int N> struct Vec { Real v[N]; };

template <typename Real, int VPL>
__device__ __forceinline__ void cp_async_vec(Real* smem_dst, const Real* gsrc) {
  constexpr int BYTES = VPL * (int)sizeof(Real);
  if constexpr (BYTES % 16 == 0) {
    for (int k = 0; k < BYTES / 16; ++k) __pipeline_memcpy_async(reinterpret_cast<char*>(smem_dst) + 16 * k, reinterpret_cast<const char*>(gsrc) + 16 * k, 16);
  } else if constexpr (BYTES == 8) {
    __pipeline_memcpy_async(smem_dst, gsrc, 8);
  } else {
    static_assert(BYTES == 4, "unsupported vector width");
    __pipeline_memcpy_async(smem_dst, gsrc, 4);
  }
}

template <typename Real, int VPL>
__device__ __forceinline__ void ldcg_vec(Real* dst, const Real* gsrc) {
  for (int k = 0; k < VPL; ++k) dst[k] = __ldcg(gsrc + k);
}

// ring slots per stage and lane (in units of Real): vectors then scalars
enum { V_THETA = 0, V_OWN_R, V_OWN_D, V_OWN_L, V_OWN_U, V_OTH_OUTROW, V_OTH_OUTCOL, NUM_VEC };
enum { S_PAR_R = 0, S_PAR_D, S_PAR_L, S_PAR_U, NUM_PAR };

template <int PW> struct ParamCount { static constexpr int value = PW == LPMP_PAIRWISE_POTTS ? 1 : 2; };

template <typename Real, int PW, int L, int VPL>
struct WavefrontCfg {
  static constexpr int G = L / VPL;          // lanes per node
  static constexpr int T = 32 / G;           // rows per warp
  static constexpr int P = ParamCount<PW>::value;
  static constexpr int STAGE_REALS = 32 * (NUM_VEC * VPL + NUM_PAR * P);  // per warp per stage
  static constexpr int SCRATCH_REALS = 2 * 32 * VPL;                       // staging of `other` for windowed min-plus
  static_assert(L % VPL == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "L/VPL must be a power of two <= 32");
};

// h(x) = min_y [c(x,y) + oth(y)] for the VPL labels of this lane.  oth: registers.  Potts / truncated linear only.
template <typename Real, int PW, int L, int VPL>
__device__ __forceinline__ void lane_min_plus(const Real* par, const Real (&oth)[VPL], Real (&h)[VPL], Real* scratch,
                                              int gl, unsigned gmask) {
  constexpr int G = L / VPL;
  const Real INF = RealTraits<Real>::inf();
  if constexpr (PW == LPMP_PAIRWISE_POTTS) {
    TwoMin<Real> t{INF, INF, 0u};
#pragma unroll
    for (int k = 0; k < VPL; ++k) two_min_insert(t, oth[k], (uint32_t)(gl * VPL + k));
    group_two_min(t, G, gmask);
    const Real lambda = par[0];
#pragma unroll
    for (int k = 0; k < VPL; ++k) h[k] = rmin(oth[k], Real(lambda + ((uint32_t)(gl * VPL + k) == t.i1 ? t.m2 : t.m1)));
  } else {
    const Real slope = par[0], trunc = par[1];
    // stage this group's `other` vector (L entries) so that every lane can read x +- d
    Real* so = scratch + (size_t)((threadIdx.x & 31) / G) * L;
#pragma unroll
    for (int k = 0; k < VPL; ++k) so[gl * VPL + k] = oth[k];
    Real g = oth[0];
#pragma unroll
    for (int k = 1; k < VPL; ++k) g = rmin(g, oth[k]);
    g = group_min(g, G, gmask);   // shuffles also order the staging stores inside the group
    __syncwarp(gmask);
    if (slope >= Real(0) && trunc >= Real(0)) {
      const Real far = trunc + g;
#pragma unroll
      for (int k = 0; k < VPL; ++k) h[k] = rmin(far, oth[k]);
      for (int d = 1; d < L; ++d) {
        const Real cd = slope * Real(d);
        if (!(cd < trunc)) break;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int x = gl * VPL + k;
          if (x >= d) h[k] = rmin(h[k], Real(cd + so[x - d]));
          if (x + d < L) h[k] = rmin(h[k], Real(cd + so[x + d]));
        }
      }
    } else {  // degenerate parameters: plain O(L^2) over the materialised costs
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int x = gl * VPL + k;
        Real m = INF;
        for (int y = 0; y < L; ++y) {
          const int d = x > y ? x - y : y - x;
          m = rmin(m, Real(rmin(Real(slope * Real(d)), trunc) + so[y]));
        }
        h[k] = m;
      }
    }
    __syncwarp(gmask);
  }
}

template <typename Real, int PW, int L, int VPL, int STAGES>
__global__ void __launch_bounds__(128) wavefront_pass_kernel(GridArgs<Real> A) {
  using Cfg = WavefrontCfg<Real, PW, L, VPL>;
  constexpr int G = Cfg::G, T = Cfg::T, P = Cfg::P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Real* ring = reinterpret_cast<Real*>(smem_raw) + (size_t)warp * (STAGES * Cfg::STAGE_REALS + Cfg::SCRATCH_REALS);
  Real* scratch = ring + STAGES * Cfg::STAGE_REALS;
  const int t = lane / G, gl = lane % G;
  const unsigned gmask = group_mask(G);
  const bool fwd = A.forward != 0;
  const bool aniso = A.repam == LPMP_REPAM_ANISOTROPIC;
  const uint32_t H = A.H, W = A.W;
  const int steps = (int)W + T - 1;
  const Real INF = RealTraits<Real>::inf();
  (void)INF;

  // roles of the planes in this direction (see header comment)
  Real* const in_row_own_p = fwd ? A.bh : A.ah;    // own vector on the row edge towards the processed neighbour
  Real* const out_row_own_p = fwd ? A.ah : A.bh;
  Real* const in_col_own_p = fwd ? A.bv : A.av;
  Real* const out_col_own_p = fwd ? A.av : A.bv;
  const Real* const out_row_oth_p = fwd ? A.bh : A.ah;  // only read in the uniform modes
  const Real* const out_col_oth_p = fwd ? A.bv : A.av;
  const Real* const in_col_oth_p = fwd ? A.av : A.bv;   // written by the previous band, read through L2

  auto vec_slot = [&](int stage, int which) -> Real* { return ring + (size_t)stage * Cfg::STAGE_REALS + ((size_t)which * 32 + lane) * VPL; };
  auto par_slot = [&](int stage, int which) -> Real* { return ring + (size_t)stage * Cfg::STAGE_REALS + (size_t)NUM_VEC * 32 * VPL + ((size_t)which * 32 + lane) * P; };

  for (;;) {
    unsigned kb = 0;
    if (lane == 0) kb = atomicAdd(A.ticket, 1u);
    kb = __shfl_sync(0xffffffffu, kb, 0);
    if (kb >= A.num_bands) break;
    const uint32_t u = kb * T + t;                    // virtual row of this lane group
    const bool row_active = u < H;
    const uint32_t r = fwd ? u : H - 1 - u;
    const uint32_t gr = A.row0 + r;                   // row in the full grid (stripes)
    const int t_last = (int)min((uint32_t)T, H - kb * T) - 1;  // last active group of this band
    unsigned ready = kb == 0 ? W : 0u;                // columns of the previous band known to be complete
    const bool has_up = gr > 0, has_down = gr + 1 < A.H_total;
    // does the column edge towards the processed / unprocessed side exist, and is it inside this stripe?
    const bool in_col_exists = fwd ? has_up : has_down;
    const bool out_col_exists = fwd ? has_down : has_up;
    const bool in_col_halo = in_col_exists && (fwd ? r == 0 : r == H - 1);     // other side lives on another stripe
    const bool out_col_halo = out_col_exists && (fwd ? r == H - 1 : r == 0);

    // issue the loads of step sp into ring stage sp % STAGES
    auto prefetch = [&](int sp) {
      const int q = sp - t;
      if (row_active && q >= 0 && q < (int)W && sp < steps) {
        const uint32_t c = fwd ? (uint32_t)q : W - 1 - (uint32_t)q;
        const uint64_t node = (uint64_t)r * W + c;
        const int st = sp % STAGES;
        cp_async_vec<Real, VPL>(vec_slot(st, V_THETA), A.theta + node * L + gl * VPL);
        const bool hasR = c + 1 < W, hasL = c > 0;
        if (hasR) cp_async_vec<Real, VPL>(vec_slot(st, V_OWN_R), A.ah + node * L + gl * VPL);
        if (hasL) cp_async_vec<Real, VPL>(vec_slot(st, V_OWN_L), A.bh + (node - 1) * L + gl * VPL);
        if (has_down) {
          const Real* src = (r + 1 < H) ? A.av + node * L : A.own_bot + (uint64_t)c * L;
          cp_async_vec<Real, VPL>(vec_slot(st, V_OWN_D), src + gl * VPL);
        }
        if (has_up) {
          const Real* src = (r > 0) ? A.bv + (node - W) * L : A.own_top + (uint64_t)c * L;
          cp_async_vec<Real, VPL>(vec_slot(st, V_OWN_U), src + gl * VPL);
        }
        // parameters of the edges this node receives from
        const bool recvR = hasR && (!aniso || !fwd), recvL = hasL && (!aniso || fwd);
        const bool recvD = has_down && (!aniso || !fwd), recvU = has_up && (!aniso || fwd);
        for (int p = 0; p < P; ++p) {
          if (recvR) __pipeline_memcpy_async(par_slot(st, S_PAR_R) + p, A.ph + node * P + p, sizeof(Real));
          if (recvL) __pipeline_memcpy_async(par_slot(st, S_PAR_L) + p, A.ph + (node - 1) * P + p, sizeof(Real));
          if (recvD) __pipeline_memcpy_async(par_slot(st, S_PAR_D) + p, ((r + 1 < H) ? A.pv + node * P : A.p_bot + (uint64_t)c * P) + p, sizeof(Real));
          if (recvU) __pipeline_memcpy_async(par_slot(st, S_PAR_U) + p, ((r > 0) ? A.pv + (node - W) * P : A.p_top + (uint64_t)c * P) + p, sizeof(Real));
        }
        if (!aniso) {  // other side of the edges towards unprocessed neighbours (state of the previous pass)
          const bool out_row_exists = fwd ? hasR : hasL;
          if (out_row_exists) cp_async_vec<Real, VPL>(vec_slot(st, V_OTH_OUTROW), out_row_oth_p + (fwd ? node : node - 1) * L + gl * VPL);
          if (out_col_exists) {
            const Real* src = out_col_halo ? (fwd ? A.halo_bot : A.halo_top) + (uint64_t)c * L
                                           : out_col_oth_p + (fwd ? node : node - W) * L;
            cp_async_vec<Real, VPL>(vec_slot(st, V_OTH_OUTCOL), src + gl * VPL);
          }
        }
      }
      __pipeline_commit();
    };

    for (int sp = 0; sp < STAGES - 1; ++sp) prefetch(sp);

    Real carry_row[VPL], carry_col[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) { carry_row[k] = Real(0); carry_col[k] = Real(0); }

    for (int s = 0; s < steps; ++s) {
      prefetch(s + STAGES - 1);
      __pipeline_wait_prior(STAGES - 1);
      const int q = s - t;
      const bool act = row_active && q >= 0 && q < (int)W;
      const uint32_t c = act ? (fwd ? (uint32_t)q : W - 1 - (uint32_t)q) : 0u;
      const uint64_t node = (uint64_t)r * W + c;
      const int st = s % STAGES;

      // ---- the column neighbour's vector: shuffle inside the band, L2 (after the progress flag) for group 0 ----
      Real col_in[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) col_in[k] = carry_col[k];
      if (kb > 0 && s < (int)W) {   // warp-uniform: group 0 is at column q = s
        if (ready <= (unsigned)s) {
          unsigned v = 0;
          do {
            if (lane == 0) v = *reinterpret_cast<volatile unsigned int*>(A.progress + (kb - 1));
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v <= (unsigned)s) __nanosleep(32);
          } while (v <= (unsigned)s);
          __threadfence();   // acquire: the producer's vector stores precede its progress store
          ready = v;
        }
        if (t == 0 && act) ldcg_vec<Real, VPL>(col_in, in_col_oth_p + (fwd ? node - W : node) * L + gl * VPL);
      }
      if (t == 0 && act && in_col_halo) ldcg_vec<Real, VPL>(col_in, (fwd ? A.halo_top : A.halo_bot) + (uint64_t)c * L + gl * VPL);

      Real th[VPL], ownR[VPL], ownD[VPL], ownL[VPL], ownU[VPL];
      const bool hasR = act && c + 1 < W, hasL = act && c > 0, hasD = act && has_down, hasU = act && has_up;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        th[k] = act ? vec_slot(st, V_THETA)[k] : Real(0);
        ownR[k] = hasR ? vec_slot(st, V_OWN_R)[k] : Real(0);
        ownD[k] = hasD ? vec_slot(st, V_OWN_D)[k] : Real(0);
        ownL[k] = hasL ? vec_slot(st, V_OWN_L)[k] : Real(0);
        ownU[k] = hasU ? vec_slot(st, V_OWN_U)[k] : Real(0);
      }
      // `other` vectors of the four edges
      Real othR[VPL], othD[VPL], othL[VPL], othU[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const Real out_row_oth = aniso ? Real(0) : vec_slot(st, V_OTH_OUTROW)[k];
        const Real out_col_oth = aniso ? Real(0) : vec_slot(st, V_OTH_OUTCOL)[k];
        othR[k] = fwd ? out_row_oth : carry_row[k];
        othL[k] = fwd ? carry_row[k] : out_row_oth;
        othD[k] = fwd ? out_col_oth : col_in[k];
        othU[k] = fwd ? col_in[k] : out_col_oth;
      }
      const bool recvR = hasR && (!aniso || !fwd), recvL = hasL && (!aniso || fwd);
      const bool recvD = hasD && (!aniso || !fwd), recvU = hasU && (!aniso || fwd);

      // ---- ReceiveMessages in the node's message order (LIFO: D, R, Lf, U; FIFO: R, D, U, Lf) ----
      auto receive = [&](bool on, Real (&own)[VPL], const Real (&oth)[VPL], int par_which) {
        // all lanes of the warp run the min-plus (group shuffles inside); inactive groups discard the result
        Real par[P];
        for (int p = 0; p < P; ++p) par[p] = on ? par_slot(st, par_which)[p] : Real(0);
        Real h[VPL];
        lane_min_plus<Real, PW, L, VPL>(par, oth, h, scratch, gl, gmask);
        if (on) {
#pragma unroll
          for (int k = 0; k < VPL; ++k) {
            const Real m = h[k] + own[k];
            th[k] = th[k] + m;
            own[k] = own[k] - m;
          }
        }
      };
      // skip whole receives only on warp-uniform conditions (the shuffles inside need the full group)
      const bool any_first = !aniso || !fwd, any_second = !aniso || fwd;
      if (!A.fifo) {
        if (any_first) { receive(recvD, ownD, othD, S_PAR_D); receive(recvR, ownR, othR, S_PAR_R); }
        if (any_second) { receive(recvL, ownL, othL, S_PAR_L); receive(recvU, ownU, othU, S_PAR_U); }
      } else {
        if (any_first) { receive(recvR, ownR, othR, S_PAR_R); receive(recvD, ownD, othD, S_PAR_D); }
        if (any_second) { receive(recvU, ownU, othU, S_PAR_U); receive(recvL, ownL, othL, S_PAR_L); }
      }

      // ---- MaximizePotentialAndComputePrimal: first minimum ----
      if (A.primal) {
        Real bv = th[0];
        uint32_t bi = (uint32_t)(gl * VPL);
#pragma unroll
        for (int k = 1; k < VPL; ++k) if (th[k] < bv) { bv = th[k]; bi = (uint32_t)(gl * VPL + k); }
        group_argmin(bv, bi, G, gmask);
        if (act && gl == 0) A.labels[node] = bi;
      }

      // ---- SendMessages with the snapshot t = theta ----
      const int n_first = (int)hasR + (int)hasD, n_second = (int)hasL + (int)hasU;
      const int n_in = fwd ? n_second : n_first, n_out = fwd ? n_first : n_second;
      const int deg = n_in + n_out;
      double wd = 0.0;
      if (aniso) wd = n_out > 0 ? 1.0 / (double)(n_in > n_out ? n_in : n_out) : 0.0;
      else if (A.repam == LPMP_REPAM_UNIFORM) wd = deg > 0 ? 1.0 / (double)deg : 0.0;
      else wd = 1.0 / (double)(deg + 1);
      const Real w = (Real)wd;
      const bool sendR = hasR && (!aniso || fwd), sendL = hasL && (!aniso || !fwd);
      const bool sendD = hasD && (!aniso || fwd), sendU = hasU && (!aniso || !fwd);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const Real d = w * th[k];   // every send uses the same snapshot value
        Real x = th[k];
        if (sendD) { ownD[k] = ownD[k] + d; x = x - d; }
        if (sendR) { ownR[k] = ownR[k] + d; x = x - d; }
        if (sendL) { ownL[k] = ownL[k] + d; x = x - d; }
        if (sendU) { ownU[k] = ownU[k] + d; x = x - d; }
        th[k] = x;
      }

      // ---- hand the freshly written vectors to the next nodes ----
      Real out_col[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        carry_row[k] = fwd ? ownR[k] : ownL[k];
        out_col[k] = fwd ? ownD[k] : ownU[k];
      }
#pragma unroll
      for (int k = 0; k < VPL; ++k) carry_col[k] = __shfl_up_sync(0xffffffffu, out_col[k], G);

      // ---- write back ----
      if (act) {
        Real* dst = A.theta + node * L + gl * VPL;
#pragma unroll
        for (int k = 0; k < VPL; ++k) dst[k] = th[k];
        if (hasR) { Real* d2 = A.ah + node * L + gl * VPL;
#pragma unroll
          for (int k = 0; k < VPL; ++k) d2[k] = ownR[k]; }
        if (hasL) { Real* d2 = A.bh + (node - 1) * L + gl * VPL;
#pragma unroll
          for (int k = 0; k < VPL; ++k) d2[k] = ownL[k]; }
        if (hasD) { Real* d2 = ((r + 1 < H) ? A.av + node * L : A.own_bot + (uint64_t)c * L) + gl * VPL;
#pragma unroll
          for (int k = 0; k < VPL; ++k) d2[k] = ownD[k]; }
        if (hasU) { Real* d2 = ((r > 0) ? A.bv + (node - W) * L : A.own_top + (uint64_t)c * L) + gl * VPL;
#pragma unroll
          for (int k = 0; k < VPL; ++k) d2[k] = ownU[k]; }
      }
      // ---- publish: the last row of the band has finished column q_last = s - t_last ----
      const int q_last = s - t_last;
      if (q_last >= 0 && q_last < (int)W && kb + 1 < A.num_bands) {
        __syncwarp();
        if (lane == t_last * G) {
          __threadfence();
          *reinterpret_cast<volatile unsigned int*>(A.progress + kb) = (unsigned)(q_last + 1);
        }
      }
    }
    __pipeline_wait_prior(0);
    __syncwarp();
  }
  (void)in_row_own_p; (void)out_row_own_p; (void)in_col_own_p; (void)out_col_own_p;
}

}  // namespace lpmp

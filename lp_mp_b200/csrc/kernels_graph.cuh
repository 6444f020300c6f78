// General graphs (level / graph-colouring schedule): node-record storage and its kernels.
//
// Replaces, for the SRMP simplex configuration on arbitrary graphs:
//   LP::ComputePass(it...) + FactorContainer::UpdateFactor / UpdateFactorPrimal
//        (/root/reference/include/LP_MP.h:981-1005,1592-1602; include/factors_messages.hxx:2256-2261,2332-2373)
//   LP::LowerBound, LP::EvaluatePrimal (include/LP_MP.h:1507-1536)
//
// Layout (graph_engine.cuh builds it).  Nodes are renumbered in SCHEDULE order k (level by level; inside a level the order is
// free), so a level / colour class is one contiguous range of records:
//   vec : per node k one record  [theta_k | own_0 | own_1 | ...]   -- its unary and ITS side of every incident edge (a_e if it
//         is the edge's first variable, b_e otherwise), slots in the reference's per-node message order (LIFO / FIFO);
//   tab : per incidence q the pairwise parameters in RECEIVER orientation (dense: tab[x_own][y_other], i.e. each table is stored
//         once per endpoint; Potts / truncated linear: the 1 / 2 parameters), in the same order as the slots;
//   per incidence: oth_off[q] = where the neighbour's side of that edge lives in vec, flags[q] (bit 0 side, bit 1 "other endpoint
//         has the higher rank"), oth_node[q] (the neighbour, for the primal cost).
// A node update therefore streams one record (read + write) and the tables of the edges it receives on, and gathers one vector per
// received edge -- the only access of a pass that is not a contiguous stream.
#pragma once
#include "device_common.cuh"

namespace lpmp {

template <typename Real>
struct GraphArgs {
  Real* vec;
  const Real* tab;
  const uint32_t* inc_off;    // [n + 1] first incidence of node k
  const uint32_t* oth_off;    // [Q]
  const uint8_t* flags;       // [Q]
  const uint32_t* oth_node;   // [Q] original index of the neighbour
  const uint32_t* node_of;    // [n] original index of node k
  const uint4* desc;          // [n][2] fast kernel: {first incidence, original index, degree, -} {oth_off | flags of slots 0..3}: everything the
                              // loads of a node depend on in ONE 32-byte read indexed by k (two dependent round trips per node instead of three)
  // mixed label counts (null for one label count per LP)
  const uint32_t* voff;       // [n + 1] record offsets
  const uint64_t* toff;       // [Q + 1] parameter offsets
  const uint32_t* lab_k;      // [n] label count of node k
  const uint32_t* lab_oth;    // [Q] label count of the neighbour
  uint32_t* labels;           // [N] by original index
  double* partial;            // bound / primal: one double per CTA
  uint32_t first, count;      // nodes [first, first + count) (schedule order)
  uint32_t L, P;              // label count (the largest one when mixed), parameters per incidence (uniform L)
  int pw, repam, forward, primal;
  int residual;               // update_factor_residual (include/factors_messages.hxx:2270-2279) instead of UpdateFactor in the plain passes
  // sweep over a subsequence of the nodes (partition passes, MODE 2 of the generic kernel): the t-th node of the launch is plan_node[t],
  // its incidences carry plan_code[plan_base[t] + slot] (bit 0 receive, bit 1 send), every send has weight plan_w[t]
  Real wtab[8];               // wtab[n] = (Real)(1.0 / (double)n), wtab[0] = 0: the send weights of the register kernel (n <= 5) without a
                              // double-precision division per node
  const uint32_t* plan_node;
  const uint32_t* plan_base;
  const uint8_t* plan_code;
  const Real* plan_w;
};

// ---- fast path: one label count L = LT (power of two <= 32), dense tables, at most DMAX incident edges ---------------------------
// LT lanes per node, lane x owns label x of every vector of the record and row x of every table: the min-plus of a receive is LT
// shuffles, adds and mins, no shared memory.  All loads of a node are issued before the first use.
// MODE 0: node update; MODE 1: lower-bound terms (read only).
// R = the largest number of edges a node of this launch receives on.  What bounds the light launches (few receives per node) is the
// number of bytes in flight per SM, i.e. the number of resident threads: the host sorts the nodes of a level by their in / out
// counts and launches each range with the instance that holds registers for exactly its receives (R = 0: ~30 registers, full
// occupancy; R = 4: ~62).
// LBF (anisotropic backward passes): LP::LowerBound accumulated on the fly -- in a backward pass a node is the last writer of its theta
// and of its side of the edges it receives on (their other side was finished by the later endpoint earlier in the pass), and h(x)
// is in registers: min theta + sum over the received edges of min_x [own(x) + h(x)]; every edge is received exactly once per pass.
template <typename Real, int LT, int DMAX, int MODE, int R = DMAX, bool LBF = false>
__global__ void __launch_bounds__(256, (R == 0 ? 8 : (R == 1 ? 6 : (LT * sizeof(Real) <= 32 ? 3 : 1)))) graph_node_kernel(const GraphArgs<Real> A) {
  static_assert(DMAX == 4, "the node descriptor holds four slots");
  constexpr int NPC = 256 / LT;   // nodes per CTA
  constexpr unsigned FULLW = 0xffffffffu;
  const int x = threadIdx.x % LT;
  const bool aniso = A.repam == LPMP_REPAM_ANISOTROPIC || A.repam == LPMP_REPAM_ANISOTROPIC2;
  const Real INF = RealTraits<Real>::inf();
  double acc = 0.0;
  // The loop is uniform over the CTA (a group without a node runs it with degree 0), so every shuffle is a full-warp shuffle with
  // width LT -- no convergence bookkeeping around the sub-warp groups.  A group serves several nodes in turn; the descriptor of the
  // next one is fetched while this one is processed: the dependent chain per node is one round trip (descriptor -> record /
  // tables / neighbour vectors).
  const uint32_t stride = gridDim.x * NPC;
  uint32_t kk = blockIdx.x * NPC + threadIdx.x / LT;
  uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
  if (kk < A.count) { n0 = __ldg(A.desc + 2 * (uint64_t)(A.first + kk)); n1 = __ldg(A.desc + 2 * (uint64_t)(A.first + kk) + 1); }
  for (uint32_t base = blockIdx.x * NPC; base < A.count; base += stride, kk += stride) {
    const bool valid = kk < A.count;
    const uint32_t k = A.first + kk;
    const uint4 d0 = n0, d1 = n1;
    n0 = make_uint4(0, 0, 0, 0); n1 = n0;
    if (kk + stride < A.count) {
      n0 = __ldg(A.desc + 2 * (uint64_t)(k + stride));
      n1 = __ldg(A.desc + 2 * (uint64_t)(k + stride) + 1);
    }
    const uint32_t q0 = d0.x, deg = valid ? d0.z : 0u;
    const uint32_t meta[DMAX] = {d1.x, d1.y, d1.z, d1.w};
    Real* rec = A.vec + ((uint64_t)k + q0) * LT;
    Real th = valid ? rec[x] : Real(0);
    Real own[DMAX];
    uint32_t fl[DMAX];
    bool recv[DMAX];
    int ord[DMAX];              // the how-manieth received edge slot s is
    uint32_t n_in = 0, n_out = 0;
    int n_recv = 0;
#pragma unroll
    for (int s = 0; s < DMAX; ++s) {
      fl[s] = meta[s] & 3u;
      const bool later = (fl[s] >> 1) & 1u;
      if (s < (int)deg) { if (later) ++n_out; else ++n_in; }
      const bool earlier = A.forward ? !later : later;
      // MODE 1: every edge once, at its first variable
      recv[s] = s < (int)deg && (MODE == 1 ? (fl[s] & 1u) == 0u : (!aniso || earlier)) && n_recv < R;
      ord[s] = n_recv;
      n_recv += recv[s] ? 1 : 0;
    }
#pragma unroll
    for (int s = 0; s < DMAX; ++s) own[s] = s < (int)deg ? rec[(1 + s) * LT + x] : Real(0);
    // the j-th received edge: label x of the neighbour's vector (the group shuffles it around) and row x of its table
    Real hm[R > 0 ? R : 1];
    if constexpr (R > 0) {
      Real oth[R], row[R][LT];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        int sj = 0;
        uint32_t mj = 0;
#pragma unroll
        for (int s = 0; s < DMAX; ++s) if (recv[s] && ord[s] == j) { sj = s; mj = meta[s]; }
        oth[j] = Real(0);
#pragma unroll
        for (int y = 0; y < LT; ++y) row[j][y] = Real(0);
        if (j < n_recv) {
          oth[j] = A.vec[(mj & ~3u) + x];
          const Real* t = A.tab + ((uint64_t)(q0 + sj) * LT + x) * LT;
          if constexpr (sizeof(Real) == 4 && LT % 4 == 0) {
#pragma unroll
            for (int y = 0; y < LT; y += 4) {
              const float4 tv = __ldg(reinterpret_cast<const float4*>(t + y));
              row[j][y] = tv.x; row[j][y + 1] = tv.y; row[j][y + 2] = tv.z; row[j][y + 3] = tv.w;
            }
          } else {
#pragma unroll
            for (int y = 0; y < LT; y += 2) {
              const double2 tv = __ldg(reinterpret_cast<const double2*>(t + y));
              row[j][y] = tv.x; row[j][y + 1] = tv.y;
            }
          }
        }
      }
      // h_j(x) = min_y [ c_j(x,y) + other_j(y) ] for the ordinals some group of the warp receives on (warp-uniform branch)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        hm[j] = INF;
        if (__any_sync(FULLW, j < n_recv)) {
#pragma unroll
          for (int y = 0; y < LT; ++y) hm[j] = rmin(hm[j], Real(row[j][y] + __shfl_sync(FULLW, oth[j], y, LT)));
        }
      }
    } else {
      hm[0] = INF;
    }
    auto h_of = [&](const int s) {   // min-marginal of slot s (static selects over the ordinals)
      Real h = hm[0];
#pragma unroll
      for (int j = 1; j < R; ++j) h = ord[s] == j ? hm[j] : h;
      return h;
    };
    if constexpr (MODE == 1) {
      // LP::LowerBound (include/LP_MP.h:1507-1518): min theta_i + sum over the edges whose first variable this node is of
      // min_x [a(x) + min_y (c(x,y) + b(y))]
      Real m = valid ? th : INF;
#pragma unroll
      for (int o = LT >> 1; o > 0; o >>= 1) m = rmin(m, __shfl_xor_sync(FULLW, m, o));
      double term = valid ? (double)m : 0.0;
#pragma unroll
      for (int s = 0; s < DMAX; ++s) {
        if (__any_sync(FULLW, recv[s])) {
          Real h = recv[s] ? Real(h_of(s) + own[s]) : INF;
#pragma unroll
          for (int o = LT >> 1; o > 0; o >>= 1) h = rmin(h, __shfl_xor_sync(FULLW, h, o));
          if (recv[s]) term += (double)h;
        }
      }
      if (x == 0) acc += term;
    } else {
      // ReceiveMessages (include/factors_messages.hxx:2440-2471), slots in the reference's message order
#pragma unroll
      for (int s = 0; s < DMAX; ++s) {
        if (recv[s]) {
          const Real m = h_of(s) + own[s];
          th = th + m;          // RepamLeft(+)   :498-500
          own[s] = own[s] - m;  // RepamRight(-)
        }
      }
      if (A.primal) {   // MaximizePotentialAndComputePrimal: first minimum
        Real v = th;
        uint32_t idx = (uint32_t)x;
#pragma unroll
        for (int o = LT >> 1; o > 0; o >>= 1) {
          const Real ov = __shfl_xor_sync(FULLW, v, o);
          const uint32_t oi = __shfl_xor_sync(FULLW, idx, o);
          if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (valid && x == 0) A.labels[d0.y] = idx;
      }
      if (deg > 0) {
        uint32_t nw;
        if (A.repam == LPMP_REPAM_ANISOTROPIC) nw = n_in > n_out ? n_in : n_out;         // include/LP_MP.h:1397
        else if (A.repam == LPMP_REPAM_ANISOTROPIC2) nw = A.forward ? n_out : n_in;      // :1134
        else if (A.repam == LPMP_REPAM_UNIFORM) nw = n_in + n_out;                       // :1437, leave_weight 0
        else nw = n_in + n_out + 1;                                                      // :1437, leave_weight 1
        const Real w = A.wtab[nw & 7u];   // (deg <= DMAX = 4: nw <= 5)
        const Real d = w * th;   // tmp_factor snapshot (:2801-2805): every send uses the theta left by the receives
        // CallSendMessages (:2698-2744)
#pragma unroll
        for (int s = 0; s < DMAX; ++s) {
          const bool later = (fl[s] >> 1) & 1u;
          const bool is_later = A.forward ? later : !later;
          if (s < (int)deg && (!aniso || is_later)) {
            own[s] = own[s] + d;  // RepamRight(+) first, :501-503
            th = th - d;          // RepamLeft(-)
          }
        }
        if (A.residual && !A.primal) {   // send_messages_residual (:2957-3007): the same messages again, cumulative weight, live theta
          Real r = Real(0);
#pragma unroll
          for (int s = 0; s < DMAX; ++s) {
            const bool later = (fl[s] >> 1) & 1u;
            const bool is_later = A.forward ? later : !later;
            if (s < (int)deg && (!aniso || is_later)) {
              r = r + w;
              const Real dr = r * th;
              own[s] = own[s] + dr;
              th = th - dr;
            }
          }
        }
      }
      if (valid) rec[x] = th;
#pragma unroll
      for (int s = 0; s < DMAX; ++s) if (s < (int)deg) rec[(1 + s) * LT + x] = own[s];
      if constexpr (LBF) {   // after the stores
        Real m = valid ? th : INF;
#pragma unroll
        for (int o = LT >> 1; o > 0; o >>= 1) m = rmin(m, __shfl_xor_sync(FULLW, m, o));
        double term = valid ? (double)m : 0.0;
#pragma unroll
        for (int s = 0; s < DMAX; ++s) {
          if (__any_sync(FULLW, recv[s])) {
            Real h = recv[s] ? Real(h_of(s) + own[s]) : INF;
#pragma unroll
            for (int o = LT >> 1; o > 0; o >>= 1) h = rmin(h, __shfl_xor_sync(FULLW, h, o));
            if (recv[s]) term += (double)h;
          }
        }
        if (x == 0) acc += term;
      }
    }
  }
  if constexpr (MODE == 1 || LBF) {
    const double sum = block_sum(acc);
    if (threadIdx.x == 0) A.partial[blockIdx.x] = sum;
  }
}

// h(x) = min_y [ c(x,y) + other(y) ] with the parameters in receiver orientation; Lk labels on this side, Lo on the other.
// One label count: the O(L) forms of group_min_plus (Potts two-min, truncated-linear window); mixed counts: all pairs.
template <typename Real>
__device__ __forceinline__ void graph_min_plus(int pw, const Real* __restrict__ p, uint32_t Lk, uint32_t Lo, bool mixed, const Real* oth, Real* hs,
                                               int G, int gl, unsigned mask) {
  if (!mixed) { group_min_plus<Real>(pw, p, Lk, 0u, oth, hs, G, gl, mask); return; }
  const Real INF = RealTraits<Real>::inf();
  for (uint32_t x = gl; x < Lk; x += G) {
    Real m = INF;
    for (uint32_t y = 0; y < Lo; ++y) {
      Real c;
      if (pw == LPMP_PAIRWISE_DENSE) c = p[(uint64_t)x * Lo + y];
      else if (pw == LPMP_PAIRWISE_POTTS) c = x == y ? Real(0) : p[0];
      else { const uint32_t d = x > y ? x - y : y - x; c = rmin(Real(p[0] * Real(d)), p[1]); }
      m = rmin(m, Real(c + oth[y]));
    }
    hs[x] = m;
  }
}

// ---- generic path: any label count (also mixed), any pairwise family, any degree -----------------------------------------------
// One sub-warp group of G lanes per node, vectors staged in shared memory (per group: th[L], oth[L], hs[L]).
// MODE 0: node update; MODE 1: lower-bound terms; MODE 2: node update of a planned sweep (weights and masks of
// ComputeAnisotropicWeights on a subsequence, /root/reference/include/LP_MP.h:1230-1415, evaluated on the host).
template <typename Real, int MODE>
__global__ void __launch_bounds__(256) graph_node_generic_kernel(const GraphArgs<Real> A, int G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Real* smem = reinterpret_cast<Real*>(smem_raw);
  const int groups_per_cta = blockDim.x / G;
  const int g = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned mask = group_mask(G);
  Real* th = smem + (size_t)g * 3 * A.L;
  Real* oth = th + A.L;
  Real* hs = oth + A.L;
  const bool aniso = A.repam == LPMP_REPAM_ANISOTROPIC || A.repam == LPMP_REPAM_ANISOTROPIC2;
  const bool mixed = A.voff != nullptr;
  const Real INF = RealTraits<Real>::inf();
  double acc = 0.0;
  for (uint32_t kk = blockIdx.x * groups_per_cta + g; kk < A.count; kk += gridDim.x * groups_per_cta) {
    const uint32_t k = MODE == 2 ? A.plan_node[A.first + kk] : A.first + kk;
    const uint8_t* code = MODE == 2 ? A.plan_code + A.plan_base[A.first + kk] : nullptr;
    const uint32_t q0 = A.inc_off[k], q1 = A.inc_off[k + 1];
    const uint32_t Lk = mixed ? A.lab_k[k] : A.L;
    Real* rec = A.vec + (mixed ? (uint64_t)A.voff[k] : ((uint64_t)k + q0) * A.L);
    __syncwarp(mask);
    for (uint32_t x = gl; x < Lk; x += G) th[x] = rec[x];
    uint32_t n_in = 0, n_out = 0;
    double term = 0.0;
    for (uint32_t q = q0; q < q1; ++q) {
      const uint32_t fl = A.flags[q];
      const bool later = (fl >> 1) & 1u;
      if (later) ++n_out; else ++n_in;
      const bool earlier = A.forward ? !later : later;
      if (MODE == 2 ? (code[q - q0] & 1u) == 0u : (MODE == 1 ? (fl & 1u) != 0u : (aniso && !earlier))) continue;
      Real* own = rec + (uint64_t)(1 + q - q0) * Lk;
      const Real* other = A.vec + A.oth_off[q];
      const uint32_t Lo = mixed ? A.lab_oth[q] : A.L;
      const Real* par = A.tab + (mixed ? A.toff[q] : (uint64_t)q * A.P);
      __syncwarp(mask);
      for (uint32_t y = gl; y < Lo; y += G) oth[y] = other[y];
      __syncwarp(mask);
      graph_min_plus<Real>(A.pw, par, Lk, Lo, mixed, oth, hs, G, gl, mask);
      if (MODE == 1) {
        Real m = INF;
        for (uint32_t x = gl; x < Lk; x += G) m = rmin(m, Real(hs[x] + own[x]));
        m = group_min(m, G, mask);
        term += (double)m;
      } else {
        for (uint32_t x = gl; x < Lk; x += G) {
          const Real ow = own[x];
          const Real m = hs[x] + ow;
          th[x] = th[x] + m;   // RepamLeft(+)   include/factors_messages.hxx:498-500
          own[x] = ow - m;     // RepamRight(-)
        }
      }
    }
    if (MODE == 1) {
      Real m = INF;
      for (uint32_t x = gl; x < Lk; x += G) m = rmin(m, th[x]);
      m = group_min(m, G, mask);
      if (gl == 0) acc += (double)m + term;
      continue;
    }
    if (A.primal) {  // MaximizePotentialAndComputePrimal: first minimum
      Real v = INF;
      uint32_t idx = 0xffffffffu;
      for (uint32_t x = gl; x < Lk; x += G) if (th[x] < v) { v = th[x]; idx = x; }
      group_argmin(v, idx, G, mask);
      if (gl == 0) A.labels[A.node_of[k]] = idx;
    }
    if (q1 > q0) {
      const uint32_t deg = n_in + n_out;
      double wd;
      if (A.repam == LPMP_REPAM_ANISOTROPIC) wd = 1.0 / (double)(n_in > n_out ? n_in : n_out);   // include/LP_MP.h:1397
      else if (A.repam == LPMP_REPAM_ANISOTROPIC2) wd = 1.0 / (double)(A.forward ? n_out : n_in);   // :1134
      else if (A.repam == LPMP_REPAM_UNIFORM) wd = 1.0 / (double)deg;         // :1437, leave_weight 0
      else wd = 1.0 / (double)(deg + 1);                                      // :1437, leave_weight 1
      const Real w = MODE == 2 ? A.plan_w[A.first + kk] : (Real)wd;
      for (uint32_t x = gl; x < Lk; x += G) hs[x] = th[x];  // tmp_factor snapshot, :2801-2805 (each lane reads back its own labels)
      // CallSendMessages (:2698-2744)
      for (uint32_t q = q0; q < q1; ++q) {
        const bool later = (A.flags[q] >> 1) & 1u;
        const bool is_later = A.forward ? later : !later;
        if (MODE == 2 ? (code[q - q0] & 2u) == 0u : (aniso && !is_later)) continue;
        Real* own = rec + (uint64_t)(1 + q - q0) * Lk;
        for (uint32_t x = gl; x < Lk; x += G) {
          const Real d = w * hs[x];
          own[x] = own[x] + d;  // RepamRight(+) first, :501-503
          th[x] = th[x] - d;    // RepamLeft(-)
        }
      }
      if (MODE != 2 && A.residual && !A.primal) {   // send_messages_residual (:2957-3007)
        Real r = Real(0);
        for (uint32_t q = q0; q < q1; ++q) {
          const bool later = (A.flags[q] >> 1) & 1u;
          const bool is_later = A.forward ? later : !later;
          if (aniso && !is_later) continue;
          r = r + w;
          Real* own = rec + (uint64_t)(1 + q - q0) * Lk;
          for (uint32_t x = gl; x < Lk; x += G) {
            const Real d = r * th[x];
            own[x] = own[x] + d;
            th[x] = th[x] - d;
          }
        }
      }
    }
    for (uint32_t x = gl; x < Lk; x += G) rec[x] = th[x];
  }
  if (MODE == 1) {
    const double sum = block_sum(acc);
    if (threadIdx.x == 0) A.partial[blockIdx.x] = sum;
  }
}

// LP::EvaluatePrimal (include/LP_MP.h:1520-1536): sum_i theta_i(x_i) + sum_e [(c_e(x_i,x_j) + a_e(x_i)) + b_e(x_j)]; one thread
// per node, every edge at its first variable
template <typename Real>
__global__ void __launch_bounds__(256) graph_primal_kernel(const GraphArgs<Real> A) {
  const bool mixed = A.voff != nullptr;
  double acc = 0.0;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < A.count; k += gridDim.x * blockDim.x) {
    const uint32_t q0 = A.inc_off[k], q1 = A.inc_off[k + 1];
    const uint32_t Lk = mixed ? A.lab_k[k] : A.L;
    const Real* rec = A.vec + (mixed ? (uint64_t)A.voff[k] : ((uint64_t)k + q0) * A.L);
    const uint32_t x = A.labels[A.node_of[k]];
    acc += (double)rec[x];
    for (uint32_t q = q0; q < q1; ++q) {
      if (A.flags[q] & 1u) continue;
      const uint32_t y = A.labels[A.oth_node[q]];
      const uint32_t Lo = mixed ? A.lab_oth[q] : A.L;
      const Real* p = A.tab + (mixed ? A.toff[q] : (uint64_t)q * A.P);
      Real c;
      if (A.pw == LPMP_PAIRWISE_DENSE) c = p[(uint64_t)x * Lo + y];
      else if (A.pw == LPMP_PAIRWISE_POTTS) c = x == y ? Real(0) : p[0];
      else { const uint32_t d = x > y ? x - y : y - x; c = rmin(Real(p[0] * Real(d)), p[1]); }
      acc += (double)Real(Real(c + rec[(uint64_t)(1 + q - q0) * Lk + x]) + A.vec[A.oth_off[q] + y]);
    }
  }
  const double s = block_sum(acc);
  if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

// deterministic sum of many partial sums (fused bound of a pass: one slot per CTA of every launch): one CTA, fixed strides, fixed tree
__global__ void __launch_bounds__(1024) wide_sum_kernel(const double* __restrict__ partial, uint32_t n, double constant, double* out) {
  double v = 0.0;
  for (uint32_t k = threadIdx.x; k < n; k += 1024) v += partial[k];
  const double s = block_sum(v);
  if (threadIdx.x == 0) *out = constant + s;
}

// CheckPrimalConsistency with one label count per variable: labels that are unset or out of range
__global__ void count_bad_labels_var_kernel(const uint32_t* __restrict__ labels, const uint32_t* __restrict__ counts, uint64_t n, unsigned int* count) {
  unsigned int bad = 0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) bad += labels[k] >= counts[k] ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

}  // namespace lpmp

// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path
// (lp_mp_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
//
// Standalone CPU restatement of the LP_MP message-passing hot path for simplex (unary/pairwise) factors in
// the SRMP configuration, behind the same C ABI as the CUDA library
// (include/lpmp_b200.h, symbols prefixed lpmp_oracle_).  Dependency-free C++17, strictly sequential like the
// reference.  Compile with -ffp-contract=off: bit-parity statements assume one rounding per + and *.
//
// Parity status: PINNED.  (1) every golden vector the reference's tests hold for this path
// (test/simplex.cpp, test/simplex_marginalization.cpp, test/potts_factor.cpp, test/vector.cpp,
// test/test_model.cpp, test/graphical_model.cpp:98-136) is checked in tests/test_oracle_goldens.py;
// (2) on random instances it is compared bit-for-bit (theta, a, b, labels) with oracle/_ref -- the
// reference's own scheduler/containers/solver/visitor compiled unmodified (oracle/ref_driver.cpp).
//
// What follows which reference lines:
//   node update (receive, argmin, snapshot, send)   include/factors_messages.hxx:2256-2261,2332-2373,2440-2471,
//                                                   2698-2744,2787-2814; sign conventions :486-510
//   per-node message order (types, LIFO/FIFO)       include/factors_messages.hxx:2005-2077,1772-1930,3402-3419
//   weights / receive masks                         include/LP_MP.h:1230-1415 (anisotropic), 1420-1449 (uniform),
//                                                   1037-1063,1487-1505 (damped, full receive mask)
//   pass / iteration                                include/LP_MP.h:869-911,981-1005,1592-1602
//   lower bound, primal cost                        include/LP_MP.h:1507-1536
//   min-marginal arithmetic                         include/vector.hxx:289-443,850-859,889-944
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

#define LPMP_FN(name) lpmp_oracle_##name
#include "../include/lpmp_b200.h"
#include "../include/lpmp_synthetic.h"

namespace {

constexpr uint32_t NO_LABEL = 0xffffffffu;

struct Incidence {
  uint64_t edge;
  uint32_t side;   // 0: node is the first variable (owns a_e), 1: second variable (owns b_e)
  uint32_t later;  // other endpoint has higher rank (is processed later in the forward pass)
};

struct OracleBase {
  virtual ~OracleBase() {}
  std::string err;
};

template <typename Real>
struct Oracle : OracleBase {
  uint64_t N = 0, E = 0;
  uint32_t L = 0;
  int pw = -1;
  bool finalized = false, force_dense = false;
  int repam = -1, edge_order = LPMP_EDGE_ORDER_LIFO, send_mode = LPMP_SEND_SHARED;
  double constant = 0.0;
  bool labels_valid = false;
  uint32_t gridH = 0, gridW = 0;
  std::vector<Real> theta, a, b, params;  // params: [E][P]
  std::vector<uint64_t> ei, ej, rank, order, inc_off;
  std::vector<Incidence> inc;
  std::vector<uint32_t> n_in, n_out, labels, kept_labels;
  std::vector<Real> h_, t_;  // scratch
  uint64_t message_updates = 0;
  // mixed label counts (add_unaries_var): every per-variable / per-edge array is packed; L = the largest count
  bool ragged = false;
  std::vector<uint32_t> Lv;
  std::vector<uint64_t> toff, aoff, boff, poff;
  // factors in insertion order as runs {kind (0 unary, 1 pairwise), first, count}: the order of the reference's f_ (include/LP_MP.h:214)
  std::vector<std::array<uint64_t, 3>> runs;
  void note_run(uint64_t kind, uint64_t first, uint64_t count) {
    if (count == 0) return;
    if (!runs.empty() && runs.back()[0] == kind && runs.back()[1] + runs.back()[2] == first) runs.back()[2] += count;
    else runs.push_back({kind, first, count});
  }

  uint32_t P() const { return pw == LPMP_PAIRWISE_DENSE ? L * L : (pw == LPMP_PAIRWISE_POTTS ? 1u : 2u); }
  uint32_t Ln(uint64_t i) const { return ragged ? Lv[i] : L; }
  uint64_t to(uint64_t i) const { return ragged ? toff[i] : i * L; }
  uint64_t ao(uint64_t e) const { return ragged ? aoff[e] : e * L; }
  uint64_t bo(uint64_t e) const { return ragged ? boff[e] : e * L; }
  uint64_t po(uint64_t e) const { return ragged ? poff[e] : e * P(); }
  uint64_t theta_len() const { return ragged ? toff[N] : N * L; }

  // c_e(x,y) of the materialised table (defines Potts / truncated linear: test/potts_factor.cpp:8-19)
  Real cost(uint64_t e, uint32_t x, uint32_t y) const {
    const Real* p = &params[po(e)];
    if (pw == LPMP_PAIRWISE_DENSE) return p[x * Ln(ej[e]) + y];
    if (pw == LPMP_PAIRWISE_POTTS) return x == y ? Real(0) : p[0];
    const uint32_t d = x > y ? x - y : y - x;
    return std::min(Real(p[0] * Real(d)), p[1]);
  }

  // h(x) = min_y [ c_e(.,.) + other_side(y) ], x on `side`.  matrix::min1(v)/min2(v): include/vector.hxx:850-859,889-913
  void min_plus(uint64_t e, uint32_t side, Real* h) const {
    const Real* other = side == 0 ? &b[bo(e)] : &a[ao(e)];
    const Real INF = std::numeric_limits<Real>::infinity();
    const bool tl_fast_ok = pw != LPMP_PAIRWISE_TRUNCLIN || (params[2 * e] >= Real(0) && params[2 * e + 1] >= Real(0));
    if (pw == LPMP_PAIRWISE_DENSE || force_dense || ragged || !tl_fast_ok) {
      const uint32_t Lown = Ln(side == 0 ? ei[e] : ej[e]), Loth = Ln(side == 0 ? ej[e] : ei[e]);
      for (uint32_t x = 0; x < Lown; ++x) {
        Real m = INF;
        for (uint32_t y = 0; y < Loth; ++y) {
          const Real c = side == 0 ? cost(e, x, y) : cost(e, y, x);
          m = std::min(m, Real(c + other[y]));
        }
        h[x] = m;
      }
    } else if (pw == LPMP_PAIRWISE_POTTS) {
      // two smallest entries: vector::two_min / min_except, include/vector.hxx:315-323,348-443
      Real m1 = INF, m2 = INF;
      uint32_t i1 = 0;
      for (uint32_t y = 0; y < L; ++y) {
        const Real v = other[y];
        if (v < m1) { m2 = m1; m1 = v; i1 = y; } else if (v < m2) { m2 = v; }
      }
      const Real lambda = params[e];
      for (uint32_t x = 0; x < L; ++x) h[x] = std::min(other[x], Real(lambda + (x == i1 ? m2 : m1)));
    } else {
      // truncated linear: terms with slope*d >= trunc collapse to trunc + global min (rounding is monotone)
      const Real slope = params[2 * e], trunc = params[2 * e + 1];
      Real gmin = INF;
      for (uint32_t y = 0; y < L; ++y) gmin = std::min(gmin, other[y]);
      const Real far = trunc + gmin;
      for (uint32_t x = 0; x < L; ++x) {
        Real m = std::min(far, other[x]);  // d = 0: c = min(0, trunc) = 0
        for (uint32_t d = 1; d < L; ++d) {
          const Real cd = slope * Real(d);
          if (!(cd < trunc)) break;
          if (x >= d) m = std::min(m, Real(cd + other[x - d]));
          if (x + d < L) m = std::min(m, Real(cd + other[x + d]));
        }
        h[x] = m;
      }
    }
  }

  Real omega(uint64_t i, bool forward) const {
    const uint32_t deg = n_in[i] + n_out[i];
    double w;
    if (repam == LPMP_REPAM_ANISOTROPIC) w = 1.0 / double(std::max(n_in[i], n_out[i]));  // include/LP_MP.h:1397
    else if (repam == LPMP_REPAM_ANISOTROPIC2) w = 1.0 / double(forward ? n_out[i] : n_in[i]);   // :1134 (no_send_messages_later)
    else if (repam == LPMP_REPAM_UNIFORM) w = 1.0 / double(deg);                            // :1437 leave_weight 0
    else w = 1.0 / double(deg + 1);                                                          // :1437 leave_weight 1
    return Real(w);
  }

  // FactorContainer::UpdateFactor / UpdateFactorPrimal, include/factors_messages.hxx:2256-2261,2332-2373
  void update_node(uint64_t i, bool forward, bool primal) {
    const uint64_t k0 = inc_off[i], k1 = inc_off[i + 1];
    if (k0 == k1 && !primal) return;
    Real* th = &theta[to(i)];
    Real* h = h_.data();
    const uint32_t L = Ln(i);   // (shadows the LP-wide count: everything below is per variable)
    const bool aniso = repam == LPMP_REPAM_ANISOTROPIC || repam == LPMP_REPAM_ANISOTROPIC2;
    for (uint64_t k = k0; k < k1; ++k) {  // ReceiveMessages, :2440-2471
      const Incidence& c = inc[k];
      const bool earlier = forward ? !c.later : c.later;
      if (aniso && !earlier) continue;
      min_plus(c.edge, c.side, h);
      Real* own = c.side == 0 ? &a[ao(c.edge)] : &b[bo(c.edge)];
      for (uint32_t x = 0; x < L; ++x) {
        const Real m = h[x] + own[x];
        th[x] += m;   // RepamLeft(+)  :498-500
        own[x] -= m;  // RepamRight(-)
      }
      ++message_updates;
    }
    if (primal) {  // MaximizePotentialAndComputePrimal: first minimum
      uint32_t arg = 0;
      for (uint32_t x = 1; x < L; ++x) if (th[x] < th[arg]) arg = x;
      labels[i] = arg;
    }
    if (k0 == k1) return;
    const Real w = omega(i, forward);
    Real* t = t_.data();
    for (uint32_t x = 0; x < L; ++x) t[x] = th[x];  // tmp_factor snapshot, :2801-2805
    for (uint64_t k = k0; k < k1; ++k) {             // CallSendMessages, :2698-2744
      const Incidence& c = inc[k];
      const bool later = forward ? c.later : !c.later;
      if (aniso && !later) continue;
      Real* own = c.side == 0 ? &a[ao(c.edge)] : &b[bo(c.edge)];
      for (uint32_t x = 0; x < L; ++x) {
        const Real d = w * t[x];
        own[x] += d;  // RepamRight(+) first, :501-503
        th[x] -= d;   // RepamLeft(-)
      }
      ++message_updates;
    }
    // (rounding passes go through UpdateFactorPrimal, which always sends in shared mode: include/LP_MP.h:1592-1602)
    if (send_mode == LPMP_SEND_RESIDUAL && !primal) {   // send_messages_residual, :2957-3007: the same messages again, cumulative weight, live theta
      Real residual_omega = Real(0);
      for (uint64_t k = k0; k < k1; ++k) {
        const Incidence& c = inc[k];
        const bool later = forward ? c.later : !c.later;
        if (aniso && !later) continue;
        residual_omega += w;
        Real* own = c.side == 0 ? &a[ao(c.edge)] : &b[bo(c.edge)];
        for (uint32_t x = 0; x < L; ++x) {
          const Real d = residual_omega * th[x];
          own[x] += d;
          th[x] -= d;
        }
      }
    }
  }

  // ---- staged optimisation: LP::put_in_same_partition / construct_factor_partition / compute_partition_pass /
  // compute_overlapping_partition_pass (include/LP_MP.h:463-465,1716-2051) ---------------------------------------------------------
  std::vector<std::array<uint64_t, 2>> partition_graph;
  std::vector<std::vector<uint64_t>> parts;   // variables of every partition in rank order, partitions in the reference's order
  bool partition_valid = false;

  void construct_partition() {
    if (partition_valid) return;
    partition_valid = true;
    // the reference's union-find (include/union_find.hxx:22-50): union by size, the root of x's set wins ties
    std::vector<uint64_t> id(N), sz(N, 1);
    std::iota(id.begin(), id.end(), uint64_t(0));
    auto find = [&](uint64_t p) { uint64_t root = p; while (root != id[root]) root = id[root]; while (p != root) { const uint64_t nx = id[p]; id[p] = root; p = nx; } return root; };
    for (const auto& pr : partition_graph) {
      const uint64_t i = find(pr[0]), j = find(pr[1]);
      if (i == j) continue;
      if (sz[i] < sz[j]) { id[i] = j; sz[j] += sz[i]; } else { id[j] = i; sz[i] += sz[j]; }
    }
    // partitions in the order of their roots (get_contiguous_ids, :70-92: unaries keep their relative order among all factors)
    std::vector<uint64_t> part_of_root(N, ~uint64_t(0));
    uint64_t np = 0;
    for (uint64_t d = 0; d < N; ++d) if (find(d) == d) part_of_root[d] = np++;
    parts.assign(np, {});
    for (uint64_t k = 0; k < N; ++k) parts[part_of_root[find(order[k])]].push_back(order[k]);   // rank order inside a partition
  }

  // one sweep over the subsequence `seq` with the weights / receive masks ComputeAnisotropicWeights derives for it
  // (include/LP_MP.h:1230-1415 with the pairwise factors outside the iterated list, :1262-1301,1319-1345)
  std::vector<uint64_t> pos_;
  void sequence_pass(const std::vector<uint64_t>& seq) {
    const uint64_t NONE = ~uint64_t(0);
    pos_.assign(N, NONE);
    for (uint64_t p = 0; p < seq.size(); ++p) pos_[seq[p]] = p;
    Real* h = h_.data();
    Real* t = t_.data();
    for (uint64_t p = 0; p < seq.size(); ++p) {
      const uint64_t i = seq[p];
      const uint64_t k0 = inc_off[i], k1 = inc_off[i + 1];
      if (k0 == k1) continue;
      Real* th = &theta[to(i)];
      const uint32_t Li = Ln(i);
      uint32_t n_send = 0;
      auto other_of = [&](const Incidence& c) { return c.side == 0 ? ej[c.edge] : ei[c.edge]; };
      // receive: the other variable is in the list and before this one; or it is outside the list and this is not the first of the list
      for (uint64_t k = k0; k < k1; ++k) {
        const Incidence& c = inc[k];
        const uint64_t po = pos_[other_of(c)];
        if (po != NONE && po > p) ++n_send;
        if (!(po != NONE ? po < p : p > 0)) continue;
        min_plus(c.edge, c.side, h);
        Real* own = c.side == 0 ? &a[ao(c.edge)] : &b[bo(c.edge)];
        for (uint32_t x = 0; x < Li; ++x) {
          const Real m = h[x] + own[x];
          th[x] += m;
          own[x] -= m;
        }
        ++message_updates;
      }
      if (n_send == 0) continue;
      const uint32_t deg = uint32_t(k1 - k0);
      const Real w = Real(1.0 / double(std::max(n_send, deg - n_send)));   // srmp_weight, :1386 (no receiving factor is in the list)
      for (uint32_t x = 0; x < Li; ++x) t[x] = th[x];
      for (uint64_t k = k0; k < k1; ++k) {
        const Incidence& c = inc[k];
        const uint64_t po = pos_[other_of(c)];
        if (!(po != NONE && po > p)) continue;
        Real* own = c.side == 0 ? &a[ao(c.edge)] : &b[bo(c.edge)];
        for (uint32_t x = 0; x < Li; ++x) {
          const Real d = w * t[x];
          own[x] += d;
          th[x] -= d;
        }
        ++message_updates;
      }
    }
  }
  static std::vector<uint64_t> concat_reversed(const std::vector<uint64_t>& first, const std::vector<uint64_t>& second) {
    std::vector<uint64_t> f(first);
    f.insert(f.end(), second.rbegin(), second.rend());
    return f;
  }
  void partition_pass(uint64_t no_passes, bool overlapping) {
    construct_partition();
    const size_t np = parts.size();
    auto reversed = [](const std::vector<uint64_t>& v) { return std::vector<uint64_t>(v.rbegin(), v.rend()); };
    if (!overlapping) {   // compute_partition_pass, :1961-1996
      for (size_t i = 0; i < np; ++i) {
        for (uint64_t it = 0; it < no_passes; ++it) { sequence_pass(parts[i]); sequence_pass(reversed(parts[i])); }
        if (i + 1 < np) sequence_pass(concat_reversed(parts[i], parts[i + 1]));
      }
      for (size_t ri = 0; ri < np; ++ri) {
        const size_t i = np - ri - 1;
        for (uint64_t it = 0; it < no_passes; ++it) { sequence_pass(parts[i]); sequence_pass(reversed(parts[i])); }
        if (i != 0) sequence_pass(concat_reversed(parts[i], parts[i - 1]));
      }
    } else {              // compute_overlapping_partition_pass, :2050-2070
      for (size_t i = 0; i + 1 < np; ++i) {
        const auto f_forward = concat_reversed(parts[i], parts[i + 1]), f_backward = concat_reversed(parts[i + 1], parts[i]);
        for (uint64_t it = 0; it < no_passes; ++it) { sequence_pass(f_forward); sequence_pass(f_backward); }
        sequence_pass(f_forward);
      }
      for (size_t ri = 1; ri < np; ++ri) {
        const size_t i = np - ri - 1;
        const auto f_forward = concat_reversed(parts[i], parts[i + 1]), f_backward = concat_reversed(parts[i + 1], parts[i]);
        for (uint64_t it = 0; it < no_passes; ++it) { sequence_pass(f_backward); sequence_pass(f_forward); }
        sequence_pass(f_backward);
      }
    }
  }

  void pass(bool forward, bool primal) {
    if (forward) for (uint64_t k = 0; k < N; ++k) update_node(order[k], true, primal);
    else for (uint64_t k = N; k-- > 0;) update_node(order[k], false, primal);
    if (primal) labels_valid = true;
  }

  double lower_bound() {  // LP::LowerBound, include/LP_MP.h:1507-1518: unaries first, then pairwise, in double
    double lb = constant;
    for (uint64_t i = 0; i < N; ++i) {
      Real m = theta[to(i)];
      for (uint32_t x = 1; x < Ln(i); ++x) m = std::min(m, theta[to(i) + x]);
      lb += double(m);
    }
    Real* h = h_.data();
    for (uint64_t e = 0; e < E; ++e) {
      min_plus(e, 0, h);
      Real m = std::numeric_limits<Real>::infinity();
      for (uint32_t x = 0; x < Ln(ei[e]); ++x) m = std::min(m, Real(h[x] + a[ao(e) + x]));
      lb += double(m);
    }
    return lb;
  }

  double evaluate_primal() const {  // LP::EvaluatePrimal, include/LP_MP.h:1520-1536
    if (!labels_valid) return std::numeric_limits<double>::infinity();
    double cost_sum = constant;
    for (uint64_t i = 0; i < N; ++i) cost_sum += double(theta[to(i) + labels[i]]);
    for (uint64_t e = 0; e < E; ++e) {
      const uint32_t x = labels[ei[e]], y = labels[ej[e]];
      cost_sum += double(Real(Real(cost(e, x, y) + a[ao(e) + x]) + b[bo(e) + y]));
    }
    return cost_sum;
  }

  int add_unaries(uint64_t n, uint32_t L_, const double* th) {
    if (finalized) { err = "add_unaries after finalize"; return LPMP_ERR_STATE; }
    if (ragged) { std::vector<uint32_t> Ls(n, L_); return add_unaries_var(n, Ls.data(), th); }
    if (L_ == 0 || (L != 0 && L != L_)) { err = "one label count per LP (add_unaries_var takes mixed counts)"; return LPMP_ERR_UNSUPPORTED; }
    L = L_;
    theta.reserve((N + n) * L);
    for (uint64_t k = 0; k < n * L; ++k) theta.push_back(Real(th[k]));
    note_run(0, N, n);
    N += n;
    return LPMP_OK;
  }
  int add_unaries_var(uint64_t n, const uint32_t* Ls, const double* th) {
    if (finalized) { err = "add_unaries after finalize"; return LPMP_ERR_STATE; }
    if (gridH) { err = "add_unaries on a grid LP"; return LPMP_ERR_STATE; }
    if (E != 0 && !ragged) { err = "add_unaries_var after pairwise factors of a uniform LP"; return LPMP_ERR_STATE; }
    if (!ragged) {   // what has been added so far had one label count
      ragged = true;
      Lv.assign(N, L);
      toff.resize(N + 1);
      for (uint64_t i = 0; i <= N; ++i) toff[i] = i * L;
    }
    uint64_t k = 0;
    for (uint64_t q = 0; q < n; ++q) {
      if (Ls[q] == 0) { err = "variable without labels"; return LPMP_ERR_INVALID; }
      Lv.push_back(Ls[q]);
      L = std::max(L, Ls[q]);
      for (uint32_t x = 0; x < Ls[q]; ++x) theta.push_back(Real(th[k++]));
      toff.push_back(theta.size());
    }
    note_run(0, N, n);
    N += n;
    return LPMP_OK;
  }

  int add_pairwise(int type, uint64_t m, const uint64_t* i, const uint64_t* j, const double* p0, const double* p1) {
    if (finalized) { err = "add_pairwise after finalize"; return LPMP_ERR_STATE; }
    if (pw != -1 && pw != type) { err = "one pairwise family per LP"; return LPMP_ERR_UNSUPPORTED; }
    pw = type;
    const uint32_t Pn = P();
    uint64_t src = 0;   // read position in p0 (dense tables of a ragged LP are packed)
    for (uint64_t k = 0; k < m; ++k) {
      if (i[k] >= N || j[k] >= N || i[k] == j[k]) { err = "edge endpoint out of range"; return LPMP_ERR_INVALID; }
      ei.push_back(i[k]);
      ej.push_back(j[k]);
      if (ragged) {
        if (poff.empty()) { poff.push_back(0); aoff.push_back(0); boff.push_back(0); }
        aoff.push_back(aoff.back() + Lv[i[k]]);
        boff.push_back(boff.back() + Lv[j[k]]);
        poff.push_back(poff.back() + (type == LPMP_PAIRWISE_DENSE ? uint64_t(Lv[i[k]]) * Lv[j[k]] : (type == LPMP_PAIRWISE_POTTS ? 1 : 2)));
      }
      const uint64_t Pe = ragged && type == LPMP_PAIRWISE_DENSE ? uint64_t(Lv[i[k]]) * Lv[j[k]] : Pn;
      if (type == LPMP_PAIRWISE_DENSE) for (uint64_t q = 0; q < Pe; ++q) params.push_back(Real(p0[src++]));
      else if (type == LPMP_PAIRWISE_POTTS) params.push_back(Real(p0[k]));
      else { params.push_back(Real(p0[k])); params.push_back(Real(p1[k])); }
    }
    note_run(1, E, m);
    E += m;
    return LPMP_OK;
  }

  // dual variables in the reference's archive layout (include/factor_archive.hxx:34-48, include/serialization.hxx:226-425): per factor
  // in insertion order its serialize_dual() members as doubles -- unary: theta; pairwise: msg1 (a), msg2 (b)
  uint64_t dual_bytes() const { return sizeof(double) * (theta.size() + a.size() + b.size()); }
  template <bool SAVE>
  void dual_archive(double* ar) {
    for (const auto& r : runs)
      for (uint64_t k = r[1]; k < r[1] + r[2]; ++k) {
        if (r[0] == 0) {
          for (uint32_t x = 0; x < Ln(k); ++x, ++ar) { if (SAVE) *ar = double(theta[to(k) + x]); else theta[to(k) + x] = Real(*ar); }
        } else {
          for (uint32_t x = 0; x < Ln(ei[k]); ++x, ++ar) { if (SAVE) *ar = double(a[ao(k) + x]); else a[ao(k) + x] = Real(*ar); }
          for (uint32_t x = 0; x < Ln(ej[k]); ++x, ++ar) { if (SAVE) *ar = double(b[bo(k) + x]); else b[bo(k) + x] = Real(*ar); }
        }
      }
  }

  int finalize() {
    if (finalized) { err = "finalize called twice"; return LPMP_ERR_STATE; }
    if (N == 0) { err = "no factors"; return LPMP_ERR_STATE; }
    if (pw == -1) pw = LPMP_PAIRWISE_POTTS;
    if (rank.empty()) { rank.resize(N); std::iota(rank.begin(), rank.end(), uint64_t(0)); }
    if (ragged && poff.empty()) { poff.push_back(0); aoff.push_back(0); boff.push_back(0); }
    a.assign(ragged ? aoff[E] : E * L, Real(0));
    b.assign(ragged ? boff[E] : E * L, Real(0));
    labels.assign(N, NO_LABEL);
    h_.resize(L);
    t_.resize(L);
    order.resize(N);
    std::iota(order.begin(), order.end(), uint64_t(0));
    std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) { return rank[x] < rank[y]; });
    for (uint64_t k = 1; k < N; ++k) if (rank[order[k]] == rank[order[k - 1]]) { err = "node ranks must be distinct"; return LPMP_ERR_INVALID; }
    // per-node message order: first-variable message type, then second-variable type (dispatcher list order,
    // include/factors_messages.hxx:3402-3419); inside a type LIFO (:2030-2041) or FIFO (:1772-1930)
    std::vector<std::vector<uint64_t>> first(N), second(N);
    for (uint64_t e = 0; e < E; ++e) { first[ei[e]].push_back(e); second[ej[e]].push_back(e); }
    inc_off.assign(N + 1, 0);
    inc.clear();
    n_in.assign(N, 0);
    n_out.assign(N, 0);
    for (uint64_t i = 0; i < N; ++i) {
      for (int s = 0; s < 2; ++s) {
        std::vector<uint64_t>& v = s == 0 ? first[i] : second[i];
        if (edge_order == LPMP_EDGE_ORDER_LIFO) std::reverse(v.begin(), v.end());
        for (uint64_t e : v) {
          const uint64_t other = s == 0 ? ej[e] : ei[e];
          const uint32_t later = rank[other] > rank[i] ? 1u : 0u;
          inc.push_back({e, uint32_t(s), later});
          if (later) ++n_out[i]; else ++n_in[i];
        }
      }
      inc_off[i + 1] = inc.size();
    }
    finalized = true;
    return LPMP_OK;
  }
};

struct Handle {
  int dtype;
  Oracle<float>* f = nullptr;
  Oracle<double>* d = nullptr;
  std::string err;
  ~Handle() { delete f; delete d; }
};

#define DISPATCH(h, expr) ((h)->dtype == LPMP_F32 ? ([&](auto& o) { return expr; })(*(h)->f) : ([&](auto& o) { return expr; })(*(h)->d))
#define GET_ERR(h) ((h)->dtype == LPMP_F32 ? (h)->f->err : (h)->d->err)

inline Handle* H(lpmp_solver* h) { return reinterpret_cast<Handle*>(h); }
inline const Handle* H(const lpmp_solver* h) { return reinterpret_cast<const Handle*>(h); }

template <typename O>
int need_final(O& o) {
  if (!o.finalized) { o.err = "finalize() has not been called"; return LPMP_ERR_STATE; }
  return LPMP_OK;
}
template <typename O>
int need_repam(O& o) {
  if (int r = need_final(o)) return r;
  if (o.repam < 0) { o.err = "no reparametrization mode set"; return LPMP_ERR_STATE; }  // include/LP_MP.h:458
  return LPMP_OK;
}

template <typename O>
int grid_create(O& o, uint32_t Hh, uint32_t W, uint32_t L, int type) {
  if (o.N != 0 || o.finalized) { o.err = "grid_create on a non-empty LP"; return LPMP_ERR_STATE; }
  if (Hh == 0 || W == 0 || L == 0 || type < 0 || type > 2) { o.err = "bad grid shape"; return LPMP_ERR_INVALID; }
  o.gridH = Hh; o.gridW = W; o.L = L; o.pw = type;
  o.N = uint64_t(Hh) * W;
  o.theta.assign(o.N * L, 0);
  // per node: horizontal edge before vertical edge
  for (uint32_t r = 0; r < Hh; ++r)
    for (uint32_t c = 0; c < W; ++c) {
      const uint64_t i = uint64_t(r) * W + c;
      if (c + 1 < W) { o.ei.push_back(i); o.ej.push_back(i + 1); }
      if (r + 1 < Hh) { o.ei.push_back(i); o.ej.push_back(i + W); }
    }
  o.E = o.ei.size();
  o.params.assign(o.E * o.P(), 0);
  o.note_run(0, 0, o.N);
  o.note_run(1, 0, o.E);
  return LPMP_OK;
}

// edge index (insertion order) of the horizontal / vertical edge leaving node (r,c)
template <typename O>
uint64_t grid_edge(const O& o, uint32_t r, uint32_t c, bool vertical) {
  // edges before row r: r rows x (W-1) horizontal + r*W vertical (all rows < r have a row below, since r < H)
  const uint64_t W = o.gridW, Hh = o.gridH;
  uint64_t base = uint64_t(r) * (W - 1) + uint64_t(r) * W;
  const bool has_v = r + 1 < Hh;
  // inside row r: nodes 0..c-1 each contribute (c'<W-1) + has_v
  base += uint64_t(c) * (1 + (has_v ? 1 : 0));  // all c' < c satisfy c' < W-1
  if (!vertical) return base;
  return base + (c + 1 < W ? 1 : 0);
}

template <typename O, typename Src>
void grid_set_pairwise_t(O& o, const Src* hz, const Src* vt) {
  using Real = typename std::remove_reference<decltype(o.params[0])>::type;
  const uint32_t P = o.P();
  for (uint32_t r = 0; r < o.gridH; ++r)
    for (uint32_t c = 0; c < o.gridW; ++c) {
      if (c + 1 < o.gridW && hz) {
        const uint64_t e = grid_edge(o, r, c, false), s = uint64_t(r) * (o.gridW - 1) + c;
        for (uint32_t q = 0; q < P; ++q) o.params[e * P + q] = Real(hz[s * P + q]);
      }
      if (r + 1 < o.gridH && vt) {
        const uint64_t e = grid_edge(o, r, c, true), s = uint64_t(r) * o.gridW + c;
        for (uint32_t q = 0; q < P; ++q) o.params[e * P + q] = Real(vt[s * P + q]);
      }
    }
}

template <typename O>
int grid_fill_synthetic(O& o, uint64_t seed, double p_lo, double p_hi) {
  using Real = typename std::remove_reference<decltype(o.params[0])>::type;
  if (o.gridH == 0) { o.err = "not a grid"; return LPMP_ERR_STATE; }
  const uint32_t P = o.P(), L = o.L;
  const uint64_t W = o.gridW;
  for (uint64_t k = 0; k < o.N * L; ++k) o.theta[k] = Real(lpmp_synth_u01(seed, LPMP_SYNTH_STREAM_UNARY, k));
  for (uint32_t r = 0; r < o.gridH; ++r)
    for (uint32_t c = 0; c < o.gridW; ++c)
      for (int v = 0; v < 2; ++v) {
        if (v == 0 ? c + 1 >= o.gridW : r + 1 >= o.gridH) continue;
        const uint64_t e = grid_edge(o, r, c, v == 1);
        const uint64_t slot = uint64_t(r) * W + c;  // device plane slot (pitch W for both planes)
        for (uint32_t q = 0; q < P; ++q) {
          double val;
          if (o.pw == LPMP_PAIRWISE_TRUNCLIN) val = q == 0 ? p_lo : p_hi;
          else val = lpmp_synth_param(seed, v == 0 ? LPMP_SYNTH_STREAM_HORIZONTAL : LPMP_SYNTH_STREAM_VERTICAL, slot * P + q, p_lo, p_hi);
          o.params[e * P + q] = Real(val);
        }
      }
  return LPMP_OK;
}

}  // namespace

extern "C" {

int LPMP_FN(create)(int dtype, int device, lpmp_solver** out) {
  (void)device;
  if (!out || (dtype != LPMP_F32 && dtype != LPMP_F64)) return LPMP_ERR_INVALID;
  Handle* h = new Handle();
  h->dtype = dtype;
  if (dtype == LPMP_F32) h->f = new Oracle<float>(); else h->d = new Oracle<double>();
  *out = reinterpret_cast<lpmp_solver*>(h);
  return LPMP_OK;
}
int LPMP_FN(destroy)(lpmp_solver* h) { delete H(h); return LPMP_OK; }
const char* LPMP_FN(last_error)(const lpmp_solver* h) { return h ? GET_ERR(H(h)).c_str() : "null handle"; }

int LPMP_FN(add_unaries)(lpmp_solver* h, uint64_t n, uint32_t L, const double* theta) { return DISPATCH(H(h), o.add_unaries(n, L, theta)); }
int LPMP_FN(add_unaries_var)(lpmp_solver* h, uint64_t n, const uint32_t* L, const double* theta) { return DISPATCH(H(h), o.add_unaries_var(n, L, theta)); }
int LPMP_FN(add_pairwise_dense)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* t) { return DISPATCH(H(h), o.add_pairwise(LPMP_PAIRWISE_DENSE, m, i, j, t, nullptr)); }
int LPMP_FN(add_pairwise_potts)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* l) { return DISPATCH(H(h), o.add_pairwise(LPMP_PAIRWISE_POTTS, m, i, j, l, nullptr)); }
int LPMP_FN(add_pairwise_trunclin)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j, const double* s, const double* t) { return DISPATCH(H(h), o.add_pairwise(LPMP_PAIRWISE_TRUNCLIN, m, i, j, s, t)); }

int LPMP_FN(set_node_ranks)(lpmp_solver* h, const uint64_t* rank) {
  return DISPATCH(H(h), ([&] {
    if (o.finalized) { o.err = "set_node_ranks after finalize"; return int(LPMP_ERR_STATE); }
    o.rank.assign(rank, rank + o.N);
    return int(LPMP_OK);
  })());
}
int LPMP_FN(set_edge_order)(lpmp_solver* h, int eo) {
  return DISPATCH(H(h), ([&] {
    if (o.finalized) { o.err = "set_edge_order after finalize"; return int(LPMP_ERR_STATE); }
    if (eo != LPMP_EDGE_ORDER_LIFO && eo != LPMP_EDGE_ORDER_FIFO) { o.err = "bad edge order"; return int(LPMP_ERR_INVALID); }
    o.edge_order = eo;
    return int(LPMP_OK);
  })());
}
int LPMP_FN(set_schedule)(lpmp_solver* h, int) { (void)h; return LPMP_OK; }  // the oracle is always sequential
int LPMP_FN(set_tuning)(lpmp_solver* h, const char*, int) { (void)h; return LPMP_OK; }

int LPMP_FN(greedy_colouring)(uint64_t n, uint64_t m, const uint64_t* i, const uint64_t* j, uint32_t* colour, uint32_t* num_colours) {
  std::vector<std::vector<uint64_t>> adj(n);
  for (uint64_t e = 0; e < m; ++e) {
    if (i[e] >= n || j[e] >= n) return LPMP_ERR_INVALID;
    adj[i[e]].push_back(j[e]);
    adj[j[e]].push_back(i[e]);
  }
  uint32_t nc = 0;
  std::vector<uint32_t> used;
  for (uint64_t v = 0; v < n; ++v) {
    used.clear();
    for (uint64_t u : adj[v]) if (u < v) used.push_back(colour[u]);
    std::sort(used.begin(), used.end());
    uint32_t c = 0;
    for (uint32_t x : used) { if (x == c) ++c; else if (x > c) break; }
    colour[v] = c;
    nc = std::max(nc, c + 1);
  }
  if (num_colours) *num_colours = nc;
  return LPMP_OK;
}

int LPMP_FN(grid_create)(lpmp_solver* h, uint32_t Hh, uint32_t W, uint32_t L, int type) { return DISPATCH(H(h), grid_create(o, Hh, W, L, type)); }
int LPMP_FN(grid_set_unaries)(lpmp_solver* h, const void* theta, int src) {
  return DISPATCH(H(h), ([&] {
    if (o.gridH == 0) { o.err = "not a grid"; return int(LPMP_ERR_STATE); }
    using Real = typename std::remove_reference<decltype(o.theta[0])>::type;
    for (uint64_t k = 0; k < o.N * o.L; ++k) o.theta[k] = src == LPMP_F32 ? Real(static_cast<const float*>(theta)[k]) : Real(static_cast<const double*>(theta)[k]);
    return int(LPMP_OK);
  })());
}
int LPMP_FN(grid_set_pairwise)(lpmp_solver* h, const void* hz, const void* vt, int src) {
  return DISPATCH(H(h), ([&] {
    if (o.gridH == 0) { o.err = "not a grid"; return int(LPMP_ERR_STATE); }
    if (src == LPMP_F32) grid_set_pairwise_t(o, static_cast<const float*>(hz), static_cast<const float*>(vt));
    else grid_set_pairwise_t(o, static_cast<const double*>(hz), static_cast<const double*>(vt));
    return int(LPMP_OK);
  })());
}
int LPMP_FN(grid_fill_synthetic)(lpmp_solver* h, uint64_t seed, double lo, double hi) { return DISPATCH(H(h), grid_fill_synthetic(o, seed, lo, hi)); }
int LPMP_FN(grid_set_stripe)(lpmp_solver* h, uint32_t, uint32_t) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_set_boundary_pairwise)(lpmp_solver* h, const void*, const void*, int) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_boundary_ptrs)(lpmp_solver* h, int, void**, uint64_t*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_set_peer_boundary)(lpmp_solver* h, int, void*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_get_boundary_vectors)(lpmp_solver* h, int, void*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_lower_bound_partial)(lpmp_solver* h, const void*, double*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_lower_bound_partial_device)(lpmp_solver* h, const void*, double*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_evaluate_primal_partial_device)(lpmp_solver* h, const void*, const uint32_t*, double*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_get_boundary_labels)(lpmp_solver* h, int, uint32_t*) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(grid_push_boundary)(lpmp_solver* h, int, void*, int) { GET_ERR(H(h)) = "stripes: device only"; return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(get_trace)(lpmp_solver*, uint64_t*, uint64_t) { return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(ipc_export)(const void*, void*) { return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(ipc_open)(const void*, void**) { return LPMP_ERR_UNSUPPORTED; }
int LPMP_FN(ipc_close)(void*) { return LPMP_ERR_UNSUPPORTED; }

int LPMP_FN(finalize)(lpmp_solver* h) { return DISPATCH(H(h), o.finalize()); }
int LPMP_FN(begin)(lpmp_solver* h) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; o.repam = -1; return int(LPMP_OK); })()); }
int LPMP_FN(end)(lpmp_solver* h) { return DISPATCH(H(h), need_final(o)); }
int LPMP_FN(set_reparametrization)(lpmp_solver* h, int mode) {
  return DISPATCH(H(h), ([&] {
    if (mode < 0 || mode > 3) { o.err = "reparametrization mode unknown"; return int(LPMP_ERR_INVALID); }
    o.repam = mode;
    return int(LPMP_OK);
  })());
}
int LPMP_FN(set_send_mode)(lpmp_solver* h, int mode) {
  return DISPATCH(H(h), ([&] {
    if (mode == LPMP_SEND_ADAPTIVE) { o.err = "adaptive sending needs send_message_to_right_improvement of the marginalisation message (absent from the reference snapshot)"; return int(LPMP_ERR_UNSUPPORTED); }
    if (mode != LPMP_SEND_SHARED && mode != LPMP_SEND_RESIDUAL) { o.err = "send mode unknown"; return int(LPMP_ERR_INVALID); }
    o.send_mode = mode;
    return int(LPMP_OK);
  })());
}
int LPMP_FN(compute_pass)(lpmp_solver* h, uint64_t) {
  return DISPATCH(H(h), ([&] { if (int r = need_repam(o)) return r; o.pass(true, false); o.pass(false, false); return int(LPMP_OK); })());
}
int LPMP_FN(compute_forward_pass)(lpmp_solver* h) { return DISPATCH(H(h), ([&] { if (int r = need_repam(o)) return r; o.pass(true, false); return int(LPMP_OK); })()); }
int LPMP_FN(compute_backward_pass)(lpmp_solver* h) { return DISPATCH(H(h), ([&] { if (int r = need_repam(o)) return r; o.pass(false, false); return int(LPMP_OK); })()); }
int LPMP_FN(compute_forward_pass_and_primal)(lpmp_solver* h, uint64_t) { return DISPATCH(H(h), ([&] { if (int r = need_repam(o)) return r; o.pass(true, true); return int(LPMP_OK); })()); }
int LPMP_FN(compute_backward_pass_and_primal)(lpmp_solver* h, uint64_t) { return DISPATCH(H(h), ([&] { if (int r = need_repam(o)) return r; o.pass(false, true); return int(LPMP_OK); })()); }
int LPMP_FN(put_in_same_partition)(lpmp_solver* h, uint64_t m, const uint64_t* i, const uint64_t* j) {
  return DISPATCH(H(h), ([&] {
    for (uint64_t k = 0; k < m; ++k) if (i[k] >= o.N || j[k] >= o.N) { o.err = "variable out of range"; return int(LPMP_ERR_INVALID); }
    for (uint64_t k = 0; k < m; ++k) o.partition_graph.push_back({i[k], j[k]});
    o.partition_valid = false;
    return int(LPMP_OK);
  })());
}
int LPMP_FN(compute_partition_pass)(lpmp_solver* h, uint64_t inner, int overlapping) {
  return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; o.partition_pass(inner, overlapping != 0); return int(LPMP_OK); })());
}
int LPMP_FN(lower_bound)(lpmp_solver* h, double* out) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; *out = o.lower_bound(); return int(LPMP_OK); })()); }
int LPMP_FN(evaluate_primal)(lpmp_solver* h, double* out) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; *out = o.evaluate_primal(); return int(LPMP_OK); })()); }
int LPMP_FN(check_primal_consistency)(lpmp_solver* h, int* out) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; *out = 1; return int(LPMP_OK); })()); }
int LPMP_FN(add_to_constant)(lpmp_solver* h, double x) { return DISPATCH(H(h), ([&] { o.constant += x; return int(LPMP_OK); })()); }
int LPMP_FN(num_factors)(const lpmp_solver* h, uint64_t* out) { return DISPATCH(H(h), ([&] { *out = o.N + o.E; return int(LPMP_OK); })()); }
int LPMP_FN(num_messages)(const lpmp_solver* h, uint64_t* out) { return DISPATCH(H(h), ([&] { *out = 2 * o.E; return int(LPMP_OK); })()); }

int LPMP_FN(get_labeling)(lpmp_solver* h, uint32_t* out) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; std::copy(o.labels.begin(), o.labels.end(), out); return int(LPMP_OK); })()); }
int LPMP_FN(keep_labeling)(lpmp_solver* h) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; o.kept_labels = o.labels; return int(LPMP_OK); })()); }
int LPMP_FN(get_kept_labeling)(lpmp_solver* h, uint32_t* out) { return DISPATCH(H(h), ([&] { std::copy(o.kept_labels.begin(), o.kept_labels.end(), out); return int(LPMP_OK); })()); }
int LPMP_FN(get_unaries)(lpmp_solver* h, double* th) { return DISPATCH(H(h), ([&] { for (uint64_t k = 0; k < o.theta.size(); ++k) th[k] = double(o.theta[k]); return int(LPMP_OK); })()); }
int LPMP_FN(get_messages)(lpmp_solver* h, double* a, double* b) {
  return DISPATCH(H(h), ([&] {
    if (int r = need_final(o)) return r;
    for (uint64_t k = 0; k < o.a.size(); ++k) if (a) a[k] = double(o.a[k]);
    for (uint64_t k = 0; k < o.b.size(); ++k) if (b) b[k] = double(o.b[k]);
    return int(LPMP_OK);
  })());
}
int LPMP_FN(set_messages)(lpmp_solver* h, const double* a, const double* b) {
  return DISPATCH(H(h), ([&] {
    if (int r = need_final(o)) return r;
    using Real = typename std::remove_reference<decltype(o.a[0])>::type;
    for (uint64_t k = 0; k < o.a.size(); ++k) if (a) o.a[k] = Real(a[k]);
    for (uint64_t k = 0; k < o.b.size(); ++k) if (b) o.b[k] = Real(b[k]);
    return int(LPMP_OK);
  })());
}
int LPMP_FN(set_unaries)(lpmp_solver* h, const double* th) {
  return DISPATCH(H(h), ([&] {
    using Real = typename std::remove_reference<decltype(o.theta[0])>::type;
    for (uint64_t k = 0; k < o.theta.size(); ++k) o.theta[k] = Real(th[k]);
    return int(LPMP_OK);
  })());
}
int LPMP_FN(dual_size)(lpmp_solver* h, uint64_t* bytes) { return DISPATCH(H(h), ([&] { if (int r = need_final(o)) return r; *bytes = o.dual_bytes(); return int(LPMP_OK); })()); }
int LPMP_FN(save_dual)(lpmp_solver* h, void* archive, uint64_t bytes) {
  return DISPATCH(H(h), ([&] {
    if (int r = need_final(o)) return r;
    if (!archive || bytes != o.dual_bytes()) { o.err = "archive size does not match this LP"; return int(LPMP_ERR_INVALID); }
    o.template dual_archive<true>(static_cast<double*>(archive));
    return int(LPMP_OK);
  })());
}
int LPMP_FN(load_dual)(lpmp_solver* h, const void* archive, uint64_t bytes) {
  return DISPATCH(H(h), ([&] {
    if (int r = need_final(o)) return r;
    if (!archive || bytes != o.dual_bytes()) { o.err = "archive size does not match this LP"; return int(LPMP_ERR_INVALID); }
    o.template dual_archive<false>(const_cast<double*>(static_cast<const double*>(archive)));
    return int(LPMP_OK);
  })());
}
int LPMP_FN(get_stats)(lpmp_solver* h, lpmp_stats* out) {
  return DISPATCH(H(h), ([&] {
    std::memset(out, 0, sizeof(*out));
    out->num_nodes = o.N; out->num_edges = o.E; out->num_labels = o.L; out->num_levels = o.N;
    out->message_updates = o.message_updates;
    const double s = sizeof(o.theta[0]);
    out->algorithmic_bytes_per_pass = s * (2.0 * o.N * o.L + double(o.E) * (o.P() + 5.0 * o.L));
    return int(LPMP_OK);
  })());
}
int LPMP_FN(synchronize)(lpmp_solver*) { return LPMP_OK; }
int LPMP_FN(stream)(lpmp_solver*, void** out) { if (out) *out = nullptr; return LPMP_OK; }

// oracle-only switch: evaluate Potts / truncated-linear min-marginals through the materialised dense table
int lpmp_oracle_force_dense(lpmp_solver* h, int on) { return DISPATCH(H(h), ([&] { o.force_dense = on != 0; return int(LPMP_OK); })()); }

}  // extern "C"

namespace {
template <typename Real>
int minmarginal(int type, uint64_t n, uint32_t L1, uint32_t L2, const double* p0, const double* p1, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  if (type != LPMP_PAIRWISE_DENSE && L1 != L2) return LPMP_ERR_INVALID;
  const Real INF = std::numeric_limits<Real>::infinity();
  for (uint64_t k = 0; k < n; ++k) {
    auto c = [&](uint32_t x, uint32_t y) -> Real {
      if (type == LPMP_PAIRWISE_DENSE) return Real(p0[(k * L1 + x) * L2 + y]);
      if (type == LPMP_PAIRWISE_POTTS) return x == y ? Real(0) : Real(p0[k]);
      const uint32_t d = x > y ? x - y : y - x;
      return std::min(Real(Real(p0[k]) * Real(d)), Real(p1[k]));
    };
    auto m1 = [&](uint32_t x) { return msg1 ? Real(msg1[k * L1 + x]) : Real(0); };
    auto m2 = [&](uint32_t y) { return msg2 ? Real(msg2[k * L2 + y]) : Real(0); };
    if (side == 0) {
      for (uint32_t x = 0; x < L1; ++x) { Real m = INF; for (uint32_t y = 0; y < L2; ++y) m = std::min(m, Real(c(x, y) + m2(y))); out[k * L1 + x] = double(Real(m + m1(x))); }
    } else {
      for (uint32_t y = 0; y < L2; ++y) { Real m = INF; for (uint32_t x = 0; x < L1; ++x) m = std::min(m, Real(c(x, y) + m1(x))); out[k * L2 + y] = double(Real(m + m2(y))); }
    }
    if (lb) {
      Real best = INF;
      for (uint32_t x = 0; x < L1; ++x) { Real m = INF; for (uint32_t y = 0; y < L2; ++y) m = std::min(m, Real(c(x, y) + m2(y))); best = std::min(best, Real(m + m1(x))); }
      lb[k] = double(best);
    }
  }
  return LPMP_OK;
}
}  // namespace

extern "C" {

}  // extern "C"

namespace {
template <typename Real>
int node_update_t(uint64_t n, uint32_t L, uint32_t deg, int pw, int repam, int forward, const uint8_t* later, const double* theta, const double* own,
                  const double* other, const double* params, double* theta_out, double* own_out, uint32_t* label) {
  const uint32_t P = pw == LPMP_PAIRWISE_DENSE ? L * L : (pw == LPMP_PAIRWISE_POTTS ? 1u : 2u);
  const bool aniso = repam == LPMP_REPAM_ANISOTROPIC || repam == LPMP_REPAM_ANISOTROPIC2;
  const Real INF = std::numeric_limits<Real>::infinity();
  std::vector<Real> th(L), t(L), h(L), ow(size_t(deg) * L);
  for (uint64_t k = 0; k < n; ++k) {
    for (uint32_t x = 0; x < L; ++x) th[x] = Real(theta[k * L + x]);
    for (uint64_t q = 0; q < uint64_t(deg) * L; ++q) ow[q] = Real(own[k * deg * L + q]);
    uint32_t n_in = 0, n_out = 0;
    for (uint32_t s = 0; s < deg; ++s) { if (later[k * deg + s]) ++n_out; else ++n_in; }
    for (uint32_t s = 0; s < deg; ++s) {   // ReceiveMessages
      const bool lat = later[k * deg + s] != 0;
      if (aniso && !(forward ? !lat : lat)) continue;
      const double* p = params + (k * deg + s) * P;
      const double* ot = other + (k * deg + s) * L;
      for (uint32_t x = 0; x < L; ++x) {
        Real m = INF;
        for (uint32_t y = 0; y < L; ++y) {
          Real c;
          if (pw == LPMP_PAIRWISE_DENSE) c = Real(p[x * L + y]);
          else if (pw == LPMP_PAIRWISE_POTTS) c = x == y ? Real(0) : Real(p[0]);
          else { const uint32_t d = x > y ? x - y : y - x; c = std::min(Real(Real(p[0]) * Real(d)), Real(p[1])); }
          m = std::min(m, Real(c + Real(ot[y])));
        }
        h[x] = m;
      }
      for (uint32_t x = 0; x < L; ++x) { const Real m = h[x] + ow[s * L + x]; th[x] += m; ow[s * L + x] -= m; }
    }
    uint32_t arg = 0;
    for (uint32_t x = 1; x < L; ++x) if (th[x] < th[arg]) arg = x;
    if (label) label[k] = arg;
    if (deg > 0) {
      double wd;
      if (repam == LPMP_REPAM_ANISOTROPIC) wd = 1.0 / double(std::max(n_in, n_out));
      else if (repam == LPMP_REPAM_ANISOTROPIC2) wd = 1.0 / double(forward ? n_out : n_in);
      else if (repam == LPMP_REPAM_UNIFORM) wd = 1.0 / double(deg);
      else wd = 1.0 / double(deg + 1);
      const Real w = Real(wd);
      for (uint32_t x = 0; x < L; ++x) t[x] = th[x];
      for (uint32_t s = 0; s < deg; ++s) {   // SendMessages
        const bool lat = later[k * deg + s] != 0;
        if (aniso && !(forward ? lat : !lat)) continue;
        for (uint32_t x = 0; x < L; ++x) { const Real d = w * t[x]; ow[s * L + x] += d; th[x] -= d; }
      }
    }
    for (uint32_t x = 0; x < L; ++x) theta_out[k * L + x] = double(th[x]);
    for (uint64_t q = 0; q < uint64_t(deg) * L; ++q) own_out[k * deg * L + q] = double(ow[q]);
  }
  return LPMP_OK;
}
}  // namespace

extern "C" {

int LPMP_FN(node_update)(int dtype, uint64_t n, uint32_t L, uint32_t deg, int pw, int repam, int forward, const uint8_t* later, const double* theta,
                         const double* own, const double* other, const double* params, double* theta_out, double* own_out, uint32_t* label) {
  if (pw < 0 || pw > 2 || repam < 0 || repam > 3 || L == 0) return LPMP_ERR_INVALID;
  return dtype == LPMP_F32 ? node_update_t<float>(n, L, deg, pw, repam, forward, later, theta, own, other, params, theta_out, own_out, label)
                           : node_update_t<double>(n, L, deg, pw, repam, forward, later, theta, own, other, params, theta_out, own_out, label);
}
int LPMP_FN(minmarginal_dense)(int dtype, uint64_t n, uint32_t L1, uint32_t L2, const double* t, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? minmarginal<float>(LPMP_PAIRWISE_DENSE, n, L1, L2, t, nullptr, msg1, msg2, side, out, lb) : minmarginal<double>(LPMP_PAIRWISE_DENSE, n, L1, L2, t, nullptr, msg1, msg2, side, out, lb);
}
int LPMP_FN(minmarginal_potts)(int dtype, uint64_t n, uint32_t L, const double* lambda, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? minmarginal<float>(LPMP_PAIRWISE_POTTS, n, L, L, lambda, nullptr, msg1, msg2, side, out, lb) : minmarginal<double>(LPMP_PAIRWISE_POTTS, n, L, L, lambda, nullptr, msg1, msg2, side, out, lb);
}
int LPMP_FN(minmarginal_trunclin)(int dtype, uint64_t n, uint32_t L, const double* slope, const double* trunc, const double* msg1, const double* msg2, int side, double* out, double* lb) {
  return dtype == LPMP_F32 ? minmarginal<float>(LPMP_PAIRWISE_TRUNCLIN, n, L, L, slope, trunc, msg1, msg2, side, out, lb) : minmarginal<double>(LPMP_PAIRWISE_TRUNCLIN, n, L, L, slope, trunc, msg1, msg2, side, out, lb);
}
int LPMP_FN(unary_min)(int dtype, uint64_t n, uint32_t L, const double* theta, double* lb, uint32_t* arg) {
  for (uint64_t k = 0; k < n; ++k) {
    uint32_t best = 0;
    for (uint32_t x = 1; x < L; ++x) {
      const bool less = dtype == LPMP_F32 ? float(theta[k * L + x]) < float(theta[k * L + best]) : theta[k * L + x] < theta[k * L + best];
      if (less) best = x;
    }
    if (lb) lb[k] = dtype == LPMP_F32 ? double(float(theta[k * L + best])) : theta[k * L + best];
    if (arg) arg[k] = best;
  }
  return LPMP_OK;
}

}  // extern "C"

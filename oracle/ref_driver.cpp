// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into, imported by or called from the product path
// (lp_mp_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
//
// Reference-core oracle: the reference's OWN scheduler, containers, solver loop and visitor
//   /root/reference/include/{LP_MP.h, factors_messages.hxx, solver.hxx, vector.hxx,
//                            visitors/standard_visitor.hxx}
// are compiled UNMODIFIED from where they lie (see oracle/Makefile; -I oracle/shims supplies stand-ins
// for the four un-vendored third-party headers).  The only pieces of the hot path that are absent from
// the reference snapshot -- the simplex factors and the marginalisation message, which live in the
// sibling repository LP_MP-MRF (README.md:10-11, no pinned version) -- are RESTATED below from their
// pinned behaviour:
//   * UnarySimplexFactor::LowerBound            test/simplex.cpp:7-12
//   * PairwiseSimplexFactor::LowerBound         test/simplex.cpp:50-66
//   * UnaryPairwiseMessage::send_message_to_left (msg -= omega * min-marginal)
//                                               test/simplex_marginalization.cpp:22-41
//   * sign conventions of send/Repam            include/factors/labeling_list_factor.hxx:479-506,
//                                               test/test_model.hxx:66-100,
//                                               include/factors_messages.hxx:486-510
//   * min-marginal arithmetic = matrix::min1(v) / min2(v)   include/vector.hxx:850-859,889-913
// Everything else (topological sort, omega / receive masks, receive-then-send node update with the
// snapshot copy, rounding and primal propagation, lower bound, StandardVisitor control law) is executed
// by the reference's code.  Output of the build goes to oracle/_ref/ (git-ignored).
#include "config.hxx"
#include "factors_messages.hxx"
#include "solver.hxx"
#include "visitors/standard_visitor.hxx"

#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace LP_MP {

constexpr INDEX no_label = std::numeric_limits<INDEX>::max();

// theta in R^L; LowerBound = min theta (test/simplex.cpp:7-12); one label as primal.
struct UnarySimplexFactor {
  UnarySimplexFactor(const REAL* begin, const INDEX L) : theta(begin, begin + L), primal(no_label) {}
  REAL LowerBound() const { return theta.min(); }
  REAL EvaluatePrimal() const {
    if (primal >= theta.size()) return std::numeric_limits<REAL>::infinity();
    return theta[primal];
  }
  // first minimum, cf. std::min_element in test/test_model.hxx:27-33
  void MaximizePotentialAndComputePrimal() {
    if (primal == no_label) primal = std::min_element(theta.begin(), theta.end()) - theta.begin();
  }
  void init_primal() { primal = no_label; }
  template <typename ARCHIVE> void serialize_dual(ARCHIVE& ar) { ar(theta); }
  template <typename ARCHIVE> void serialize_primal(ARCHIVE& ar) { ar(primal); }
  auto export_variables() { return std::tie(theta); }
  INDEX size() const { return theta.size(); }

  vector<REAL> theta;
  INDEX primal;
};

// theta(x1,x2) = cost(x1,x2) + msg1(x1) + msg2(x2)   (test/simplex.cpp:50-66, test/potts_factor.cpp:50-55)
struct PairwiseSimplexFactor {
  PairwiseSimplexFactor(const INDEX d1, const INDEX d2, const REAL* table) : cost(d1, d2), msg1(d1, 0.0), msg2(d2, 0.0) {
    for (INDEX x1 = 0; x1 < d1; ++x1)
      for (INDEX x2 = 0; x2 < d2; ++x2) cost(x1, x2) = table[x1 * d2 + x2];
    init_primal();
  }
  // msg1(x1) + min_{x2}[cost(x1,x2) + msg2(x2)]; arithmetic of matrix::min1(v), include/vector.hxx:850-859,931-944
  vector<REAL> min_marginal_1() const {
    vector<REAL> m = cost.min1(msg2);
    for (INDEX x1 = 0; x1 < m.size(); ++x1) m[x1] += msg1[x1];
    return m;
  }
  // msg2(x2) + min_{x1}[cost(x1,x2) + msg1(x1)]; arithmetic of matrix::min2(v), include/vector.hxx:889-913
  vector<REAL> min_marginal_2() const {
    vector<REAL> m = cost.min2(msg1);
    for (INDEX x2 = 0; x2 < m.size(); ++x2) m[x2] += msg2[x2];
    return m;
  }
  REAL LowerBound() const { return min_marginal_1().min(); }
  REAL EvaluatePrimal() const {
    if (primal[0] >= cost.dim1() || primal[1] >= cost.dim2()) return std::numeric_limits<REAL>::infinity();
    return cost(primal[0], primal[1]) + msg1[primal[0]] + msg2[primal[1]];
  }
  void init_primal() { primal[0] = no_label; primal[1] = no_label; }
  template <typename ARCHIVE> void serialize_dual(ARCHIVE& ar) { ar(msg1); ar(msg2); }
  template <typename ARCHIVE> void serialize_primal(ARCHIVE& ar) { ar(primal); }
  auto export_variables() { return std::tie(msg1, msg2); }

  matrix<REAL> cost;
  vector<REAL> msg1, msg2;
  std::array<INDEX, 2> primal;
};

// SIDE = which variable of the pairwise factor the unary is tied to.  The unary is the LEFT factor of the
// message, the pairwise the RIGHT one (message_passing_schedule::left: the unary receives and sends).
template <Chirality SIDE>
struct UnaryPairwiseMessage {
  // msg -= omega * min-marginal  (test/simplex_marginalization.cpp:26-40)
  template <typename RIGHT_FACTOR, typename MSG>
  void send_message_to_left(const RIGHT_FACTOR& r, MSG& msg, const REAL omega) {
    if (SIDE == Chirality::left) {
      const vector<REAL> m = r.min_marginal_1();
      msg -= omega * m;
    } else {
      const vector<REAL> m = r.min_marginal_2();
      msg -= omega * m;
    }
  }
  // msg -= omega * unary   (cf. include/factors/labeling_list_factor.hxx:501-506)
  template <typename LEFT_FACTOR, typename MSG>
  void send_message_to_right(const LEFT_FACTOR& l, MSG& msg, const REAL omega) {
    msg -= omega * l.theta;
  }
  template <typename LEFT_FACTOR, typename MSG>
  void RepamLeft(LEFT_FACTOR& l, const MSG& m) {
    for (INDEX i = 0; i < l.theta.size(); ++i) l.theta[i] += m[i];
  }
  template <typename LEFT_FACTOR>
  void RepamLeft(LEFT_FACTOR& l, const REAL v, const INDEX i) { l.theta[i] += v; }
  template <typename RIGHT_FACTOR, typename MSG>
  void RepamRight(RIGHT_FACTOR& r, const MSG& m) {
    auto& s = (SIDE == Chirality::left) ? r.msg1 : r.msg2;
    for (INDEX i = 0; i < s.size(); ++i) s[i] += m[i];
  }
  template <typename RIGHT_FACTOR>
  void RepamRight(RIGHT_FACTOR& r, const REAL v, const INDEX i) {
    if (SIDE == Chirality::left) r.msg1[i] += v; else r.msg2[i] += v;
  }
  template <typename LEFT_FACTOR, typename RIGHT_FACTOR>
  void ComputeRightFromLeftPrimal(const LEFT_FACTOR& l, RIGHT_FACTOR& r) {
    r.primal[SIDE == Chirality::left ? 0 : 1] = l.primal;
  }
  template <typename LEFT_FACTOR, typename RIGHT_FACTOR>
  bool CheckPrimalConsistency(const LEFT_FACTOR& l, const RIGHT_FACTOR& r) const {
    return l.primal == r.primal[SIDE == Chirality::left ? 0 : 1];
  }
};

#ifndef LPMP_REF_NO_LEFT
#define LPMP_REF_NO_LEFT variableMessageNumber
#endif
#ifndef LPMP_REF_NO_RIGHT
#define LPMP_REF_NO_RIGHT 1
#endif

struct FMC_SRMP_RESTATED {
  constexpr static const char* name = "SRMP (restated simplex factors on the reference core)";
  using UnaryFactor = FactorContainer<UnarySimplexFactor, FMC_SRMP_RESTATED, 0, true>;
  using PairwiseFactor = FactorContainer<PairwiseSimplexFactor, FMC_SRMP_RESTATED, 1, false>;
  using UnaryPairwiseMessageLeftContainer =
      MessageContainer<UnaryPairwiseMessage<Chirality::left>, 0, 1, message_passing_schedule::left, LPMP_REF_NO_LEFT,
                       LPMP_REF_NO_RIGHT, FMC_SRMP_RESTATED, 0>;
  using UnaryPairwiseMessageRightContainer =
      MessageContainer<UnaryPairwiseMessage<Chirality::right>, 0, 1, message_passing_schedule::left, LPMP_REF_NO_LEFT,
                       LPMP_REF_NO_RIGHT, FMC_SRMP_RESTATED, 1>;
  using FactorList = meta::list<UnaryFactor, PairwiseFactor>;
  using MessageList = meta::list<UnaryPairwiseMessageLeftContainer, UnaryPairwiseMessageRightContainer>;
  using ProblemDecompositionList = meta::list<>;
};

// StandardVisitor whose visit() also records what the solver reported (the reference only prints it).
class RecordingVisitor : public StandardVisitor {
public:
  using StandardVisitor::StandardVisitor;
  LpControl visit(const LpControl c, const REAL lowerBound, const REAL primalBound) {
    lb_hist.push_back(lowerBound);
    ub_hist.push_back(primalBound);
    flags.push_back((c.computeLowerBound ? 1 : 0) | (c.computePrimal ? 2 : 0) | (int(c.repam) << 2));
    return StandardVisitor::visit(c, lowerBound, primalBound);
  }
  std::vector<REAL> lb_hist, ub_hist;
  std::vector<int> flags;
};

#ifdef LPMP_DRIVE_LP_CUDA
}  // namespace LP_MP
// the SAME driver around the product's reference-side binding (include/LP_cuda.hxx): the reference's unmodified Solver::Solve()
// loop, MpRoundingSolver::Iterate and StandardVisitor control law drive the CUDA library instead of LP<FMC>
#include "LP_cuda.hxx"
namespace LP_MP {
using DrivenLP = LP_cuda;
#else
using DrivenLP = LP<FMC_SRMP_RESTATED>;
#endif
using RefSolver = MpRoundingSolver<Solver<DrivenLP, RecordingVisitor>>;

struct RefSolverAccess : public RefSolver {
  using RefSolver::RefSolver;
  RecordingVisitor& visitor() { return this->visitor_; }
};

struct RefHandle {
  std::unique_ptr<RefSolverAccess> solver;
#ifdef LPMP_DRIVE_LP_CUDA
  std::vector<LP_cuda::factor*> unaries, pairwise;
#else
  std::vector<FMC_SRMP_RESTATED::UnaryFactor*> unaries;
  std::vector<FMC_SRMP_RESTATED::PairwiseFactor*> pairwise;
#endif
  std::vector<std::array<INDEX, 2>> edges;
};

// row-major table view with the call operator AddPairwiseFactor expects (matrix<REAL>, include/vector.hxx:833)
struct TableView {
  const double* p; INDEX d2;
  double operator()(const INDEX x, const INDEX y) const { return p[x * d2 + y]; }
};

}  // namespace LP_MP

using namespace LP_MP;

extern "C" {

// options are argv-style; options[0] is the program name (cf. test/graphical_model.cpp:32-42)
void* lpmp_ref_create(int n_options, const char* const* options) {
  std::vector<std::string> opts;
  for (int i = 0; i < n_options; ++i) opts.emplace_back(options[i]);
  if (opts.empty()) opts.emplace_back("lpmp_ref");
  auto* h = new RefHandle();
  h->solver.reset(new RefSolverAccess(opts));
  return h;
}

void lpmp_ref_destroy(void* hv) { delete static_cast<RefHandle*>(hv); }

int64_t lpmp_ref_add_unary(void* hv, uint64_t L, const double* theta) {
  auto* h = static_cast<RefHandle*>(hv);
  auto& lp = h->solver->GetLP();
#ifdef LPMP_DRIVE_LP_CUDA
  auto* f = lp.AddUnaryFactor(std::vector<REAL>(theta, theta + L));
#else
  auto* f = lp.template add_factor<FMC_SRMP_RESTATED::UnaryFactor>(theta, INDEX(L));
#endif
  h->unaries.push_back(f);
  return int64_t(h->unaries.size()) - 1;
}

// table is row-major L_i x L_j.  Relations u_i -> p -> u_j as the MRF constructor of LP_MP-MRF adds them
// (include/LP_MP.h:697-702 AddFactorRelation).
int64_t lpmp_ref_add_pairwise(void* hv, uint64_t i, uint64_t j, const double* table) {
  auto* h = static_cast<RefHandle*>(hv);
  if (i >= h->unaries.size() || j >= h->unaries.size() || i == j) return -1;
  auto& lp = h->solver->GetLP();
#ifdef LPMP_DRIVE_LP_CUDA
  h->pairwise.push_back(lp.AddPairwiseFactor(i, j, TableView{table, lp.GetNumberOfLabels(j)}));
  h->edges.push_back({INDEX(i), INDEX(j)});
  return int64_t(h->pairwise.size()) - 1;
#else
  auto* ui = h->unaries[i];
  auto* uj = h->unaries[j];
  const INDEX d1 = ui->GetFactor()->size(), d2 = uj->GetFactor()->size();
  auto* p = lp.template add_factor<FMC_SRMP_RESTATED::PairwiseFactor>(d1, d2, table);
  lp.template add_message<FMC_SRMP_RESTATED::UnaryPairwiseMessageLeftContainer>(ui, p);
  lp.template add_message<FMC_SRMP_RESTATED::UnaryPairwiseMessageRightContainer>(uj, p);
  lp.AddFactorRelation(ui, p);
  lp.AddFactorRelation(p, uj);
  h->pairwise.push_back(p);
  h->edges.push_back({INDEX(i), INDEX(j)});
  return int64_t(h->pairwise.size()) - 1;
#endif
}

#ifdef LPMP_DRIVE_LP_CUDA
// re-orient edge e = (i, j): AddFactorRelation(u_j, p), AddFactorRelation(p, u_i) (include/LP_MP.h:697-702)
void lpmp_ref_reverse_edge(void* hv, uint64_t e) {
  auto* h = static_cast<RefHandle*>(hv);
  auto& lp = h->solver->GetLP();
  lp.AddFactorRelation(h->unaries[h->edges[e][1]], h->pairwise[e]);
  lp.AddFactorRelation(h->pairwise[e], h->unaries[h->edges[e][0]]);
}
#endif

// LP::put_in_same_partition on the unary factors of variables i and j (include/LP_MP.h:465)
void lpmp_ref_put_in_same_partition(void* hv, uint64_t i, uint64_t j) {
  auto* h = static_cast<RefHandle*>(hv);
  h->solver->GetLP().put_in_same_partition(h->unaries[i], h->unaries[j]);
}

void lpmp_ref_add_to_constant(void* hv, double x) { static_cast<RefHandle*>(hv)->solver->GetLP().add_to_constant(x); }

void lpmp_ref_begin(void* hv) { static_cast<RefHandle*>(hv)->solver->GetLP().Begin(); }

// 0 anisotropic, 1 uniform, 2 damped_uniform, 3 anisotropic2
void lpmp_ref_set_reparametrization(void* hv, int mode) {
  const LPReparametrizationMode m = mode == 0 ? LPReparametrizationMode::Anisotropic
                                   : mode == 1 ? LPReparametrizationMode::Uniform
                                   : mode == 2 ? LPReparametrizationMode::DampedUniform
                                               : LPReparametrizationMode::Anisotropic2;
  static_cast<RefHandle*>(hv)->solver->GetLP().set_reparametrization(m);
}

void lpmp_ref_compute_pass(void* hv, uint64_t iter) { static_cast<RefHandle*>(hv)->solver->GetLP().ComputePass(iter); }
void lpmp_ref_forward_pass(void* hv) { static_cast<RefHandle*>(hv)->solver->GetLP().ComputeForwardPass(); }
void lpmp_ref_backward_pass(void* hv) { static_cast<RefHandle*>(hv)->solver->GetLP().ComputeBackwardPass(); }
void lpmp_ref_forward_pass_and_primal(void* hv, uint64_t iter) {
  static_cast<RefHandle*>(hv)->solver->GetLP().ComputeForwardPassAndPrimal(iter);
}
void lpmp_ref_backward_pass_and_primal(void* hv, uint64_t iter) {
  static_cast<RefHandle*>(hv)->solver->GetLP().ComputeBackwardPassAndPrimal(iter);
}
double lpmp_ref_lower_bound(void* hv) { return static_cast<RefHandle*>(hv)->solver->GetLP().LowerBound(); }
double lpmp_ref_evaluate_primal(void* hv) { return static_cast<RefHandle*>(hv)->solver->GetLP().EvaluatePrimal(); }

uint64_t lpmp_ref_num_factors(void* hv) { return static_cast<RefHandle*>(hv)->solver->GetLP().GetNumberOfFactors(); }
uint64_t lpmp_ref_num_messages(void* hv) { return static_cast<RefHandle*>(hv)->solver->GetLP().GetNumberOfMessages(); }
#ifdef LPMP_DRIVE_LP_CUDA
// state of the CUDA LP through the C ABI (packed layouts, include/lpmp_b200.h)
static std::vector<double> cuda_unaries(RefHandle* h) {
  auto& lp = h->solver->GetLP();
  INDEX n = 0;
  for (INDEX i = 0; i < h->unaries.size(); ++i) n += lp.GetNumberOfLabels(i);
  std::vector<double> th(n);
  if (lpmp_get_unaries(lp.handle(), th.data()) != LPMP_OK) throw std::runtime_error(lpmp_last_error(lp.handle()));
  return th;
}
uint64_t lpmp_ref_no_send_messages(void*, uint64_t) { return 0; }
void lpmp_ref_get_unary(void* hv, uint64_t i, double* out) {
  auto* h = static_cast<RefHandle*>(hv);
  auto& lp = h->solver->GetLP();
  const std::vector<double> th = cuda_unaries(h);
  INDEX off = 0;
  for (INDEX k = 0; k < i; ++k) off += lp.GetNumberOfLabels(k);
  for (INDEX x = 0; x < lp.GetNumberOfLabels(i); ++x) out[x] = th[off + x];
}
void lpmp_ref_get_pairwise_msgs(void* hv, uint64_t e, double* msg1, double* msg2) {
  auto* h = static_cast<RefHandle*>(hv);
  auto& lp = h->solver->GetLP();
  INDEX na = 0, nb = 0, oa = 0, ob = 0;
  for (INDEX k = 0; k < h->edges.size(); ++k) {
    if (k == e) { oa = na; ob = nb; }
    na += lp.GetNumberOfLabels(h->edges[k][0]); nb += lp.GetNumberOfLabels(h->edges[k][1]);
  }
  std::vector<double> a(na), b(nb);
  if (lpmp_get_messages(lp.handle(), a.data(), b.data()) != LPMP_OK) throw std::runtime_error(lpmp_last_error(lp.handle()));
  for (INDEX x = 0; x < lp.GetNumberOfLabels(h->edges[e][0]); ++x) msg1[x] = a[oa + x];
  for (INDEX x = 0; x < lp.GetNumberOfLabels(h->edges[e][1]); ++x) msg2[x] = b[ob + x];
}
void lpmp_ref_get_labels(void* hv, uint64_t* out) {
  auto* h = static_cast<RefHandle*>(hv);
  const std::vector<uint32_t> l = h->solver->GetLP().labeling();
  for (std::size_t i = 0; i < l.size(); ++i) out[i] = l[i] == 0xffffffffu ? no_label : l[i];
}
uint64_t lpmp_ref_dual_size(void* hv) { return static_cast<RefHandle*>(hv)->solver->GetLP().dual_size_in_bytes(); }
void lpmp_ref_save_dual(void* hv, void* out) { static_cast<RefHandle*>(hv)->solver->GetLP().save_dual(out); }
void lpmp_ref_load_dual(void* hv, const void* in) { static_cast<RefHandle*>(hv)->solver->GetLP().load_dual(in); }
#else
uint64_t lpmp_ref_no_send_messages(void* hv, uint64_t i) { return static_cast<RefHandle*>(hv)->unaries[i]->no_send_messages(); }

void lpmp_ref_get_unary(void* hv, uint64_t i, double* out) {
  const auto& t = static_cast<RefHandle*>(hv)->unaries[i]->GetFactor()->theta;
  for (INDEX x = 0; x < t.size(); ++x) out[x] = t[x];
}
void lpmp_ref_get_pairwise_msgs(void* hv, uint64_t e, double* msg1, double* msg2) {
  auto* f = static_cast<RefHandle*>(hv)->pairwise[e]->GetFactor();
  for (INDEX x = 0; x < f->msg1.size(); ++x) msg1[x] = f->msg1[x];
  for (INDEX x = 0; x < f->msg2.size(); ++x) msg2[x] = f->msg2[x];
}
void lpmp_ref_get_labels(void* hv, uint64_t* out) {
  auto* h = static_cast<RefHandle*>(hv);
  for (std::size_t i = 0; i < h->unaries.size(); ++i) out[i] = h->unaries[i]->GetFactor()->primal;
}

// The dual variables of ALL factors in the reference's own archive: exactly what factor_archive<serialization_functor::dual>
// (include/factor_archive.hxx:34-48) does -- an allocate_archive pass for the size, then save_archive / load_archive over the
// factors in f_ order (include/serialization.hxx:226-425; FactorContainer::serialize_dual, include/factors_messages.hxx:3196-3209)
uint64_t lpmp_ref_dual_size(void* hv) {
  auto& lp = static_cast<RefHandle*>(hv)->solver->GetLP();
  allocate_archive aa;
  for (INDEX i = 0; i < lp.GetNumberOfFactors(); ++i) lp.GetFactor(i)->serialize_dual(aa);
  return aa.size();
}
void lpmp_ref_save_dual(void* hv, void* out) {
  auto& lp = static_cast<RefHandle*>(hv)->solver->GetLP();
  allocate_archive aa;
  for (INDEX i = 0; i < lp.GetNumberOfFactors(); ++i) lp.GetFactor(i)->serialize_dual(aa);
  serialization_archive ar(aa);
  save_archive sa(ar);
  for (INDEX i = 0; i < lp.GetNumberOfFactors(); ++i) lp.GetFactor(i)->serialize_dual(sa);
  std::memcpy(out, ar.begin(), ar.size());
}
void lpmp_ref_load_dual(void* hv, const void* in) {
  auto& lp = static_cast<RefHandle*>(hv)->solver->GetLP();
  allocate_archive aa;
  for (INDEX i = 0; i < lp.GetNumberOfFactors(); ++i) lp.GetFactor(i)->serialize_dual(aa);
  serialization_archive ar(aa);
  std::memcpy(ar.begin(), in, ar.size());
  load_archive la(ar);
  for (INDEX i = 0; i < lp.GetNumberOfFactors(); ++i) lp.GetFactor(i)->serialize_dual(la);
}
#endif

// single-factor entry points (goldens of test/simplex.cpp, test/simplex_marginalization.cpp)
double lpmp_ref_unary_lower_bound(uint64_t L, const double* theta) {
  UnarySimplexFactor f(theta, INDEX(L));
  return f.LowerBound();
}
// side 0: marginalise out x2 (result over x1); side 1: marginalise out x1.  out = min-marginal, lb = LowerBound.
void lpmp_ref_pairwise_min_marginal(uint64_t d1, uint64_t d2, const double* table, const double* msg1, const double* msg2,
                                    int side, double* out, double* lb) {
  PairwiseSimplexFactor f(d1, d2, table);
  for (INDEX x = 0; x < d1; ++x) f.msg1[x] = msg1 ? msg1[x] : 0.0;
  for (INDEX x = 0; x < d2; ++x) f.msg2[x] = msg2 ? msg2[x] : 0.0;
  const vector<REAL> m = side == 0 ? f.min_marginal_1() : f.min_marginal_2();
  for (INDEX x = 0; x < m.size(); ++x) out[x] = m[x];
  if (lb) *lb = f.LowerBound();
}

// Runs the reference's own Solver::Solve() loop (include/solver.hxx:230-257) with the StandardVisitor
// control law; copies what the solver reported at each visit.  Returns number of visits.
uint64_t lpmp_ref_solve(void* hv, double* lb_hist, double* ub_hist, int32_t* flags, uint64_t cap, double* final_lb,
                        double* final_ub, double* seconds) {
  auto* h = static_cast<RefHandle*>(hv);
  const auto t0 = std::chrono::steady_clock::now();
  h->solver->Solve();
  const auto t1 = std::chrono::steady_clock::now();
  auto& v = h->solver->visitor();
  const uint64_t n = v.lb_hist.size();
  for (uint64_t k = 0; k < n && k < cap; ++k) {
    if (lb_hist) lb_hist[k] = v.lb_hist[k];
    if (ub_hist) ub_hist[k] = v.ub_hist[k];
    if (flags) flags[k] = v.flags[k];
  }
  if (final_lb) *final_lb = h->solver->lower_bound();
  if (final_ub) *final_ub = h->solver->primal_cost();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return n;
}

}  // extern "C"

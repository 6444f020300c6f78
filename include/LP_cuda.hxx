// LP_cuda.hxx -- the reference-side binding of the B200 message-passing path.
//
// Goes next to include/LP_MP.h in the reference tree (it needs nothing of this repository but lpmp_b200.h and
// liblpmp_b200.so).  LP_cuda is an LP_TYPE for the reference's UNMODIFIED
//     Solver<LP_TYPE, VISITOR>              /root/reference/include/solver.hxx:29-30
//     MpRoundingSolver<Solver<...>>         /root/reference/include/solver.hxx:381-400
//     StandardVisitor                       /root/reference/include/visitors/standard_visitor.hxx:48-199
// i.e. it provides exactly what those templates touch: `using FMC` with `name` / `ProblemDecompositionList`
// (solver.hxx:34-35,56), the constructor LP(TCLAP::CmdLine&) (LP_MP.h:212,586-594), Begin / End / set_reparametrization /
// ComputePass / Compute{Forward,Backward}Pass[AndPrimal] / LowerBound / EvaluatePrimal / CheckPrimalConsistency
// (solver.hxx:185,233,248,263,269,274,281,306,323,390-392), GetNumberOfFactors (standard_visitor.hxx:294), plus the construction
// calls a problem constructor makes: add factors, add the u_i -> p -> u_j relations (LP_MP.h:697-702), add_to_constant,
// put_in_same_partition (LP_MP.h:462-465).  Factors are the SRMP configuration's: unary and pairwise simplex factors tied by
// marginalisation messages (LP_MP-MRF; surface of its MRF problem constructor as used in test/graphical_model.cpp:98-108).
//
// Errors cross the C ABI as status codes and are re-thrown as std::runtime_error, which is what the reference throws
// (LP_MP.h:458).  One handle = one device = one host thread (the reference is single-threaded).
#ifndef LP_MP_LP_CUDA_HXX
#define LP_MP_LP_CUDA_HXX

#include "config.hxx"        // REAL, INDEX, LPReparametrizationMode, TCLAP
#include "meta/meta.hpp"
#include "lpmp_b200.h"

#include <algorithm>
#include <cstdint>
#include <deque>
#include <limits>
#include <queue>
#include <stdexcept>
#include <string>
#include <vector>

namespace LP_MP {

struct FMC_SRMP_CUDA {   // what Solver reads from LP_TYPE::FMC (include/solver.hxx:34-35,56)
  constexpr static const char* name = "SRMP on B200 (liblpmp_b200)";
  using ProblemDecompositionList = meta::list<>;
};

class LP_cuda {
public:
  using FMC = FMC_SRMP_CUDA;

  // what add_factor returns in the reference is a FactorContainer*; here a factor is a (kind, index) record with a stable address
  struct factor {
    enum kind_t : unsigned char { unary, pairwise } kind;
    INDEX index;      // unary: node number; pairwise: edge number (insertion order)
  };

  // LP<FMC>::LP(cmd), include/LP_MP.h:586-594: the options the reference LP registers (same names), plus device selection
  explicit LP_cuda(TCLAP::CmdLine& cmd)
      : reparametrization_type_arg_("", "reparametrizationType", "message sending type: ", false, "shared", "{shared|residual|partition|overlapping_partition|adaptive}", cmd),
        inner_iteration_number_arg_("", "innerIteration", "number of iterations in inner loop in partition reparamtrization, default = 5", false, 5, "positive integer", cmd),
        device_arg_("", "cudaDevice", "CUDA device ordinal", false, 0, "integer", cmd),
        precision_arg_("", "cudaPrecision", "storage / arithmetic type on the device: f64 reproduces REAL = double bit for bit, f32 is the throughput mode", false, "f64", "{f64|f32}", cmd) {}
  ~LP_cuda() { if (h_) lpmp_destroy(h_); }
  LP_cuda(const LP_cuda&) = delete;
  LP_cuda& operator=(const LP_cuda&) = delete;

  // ---- problem construction (mrf.AddUnaryFactor / AddPairwiseFactor, test/graphical_model.cpp:98-108) ------------------------
  template <typename VECTOR>
  factor* AddUnaryFactor(const VECTOR& cost) {
    need_open();
    label_count_.push_back((uint32_t)cost.size());
    for (INDEX x = 0; x < cost.size(); ++x) theta_.push_back(cost[x]);
    factors_.push_back({factor::unary, unaries_.size()});
    unaries_.push_back(&factors_.back());
    return &factors_.back();
  }
  // table(x_i, x_j): anything with operator()(x1, x2) (matrix<REAL>, include/vector.hxx:833); adds the factor, its two
  // marginalisation messages and the relations u_i -> p -> u_j (include/LP_MP.h:266-285,697-702)
  template <typename MATRIX>
  factor* AddPairwiseFactor(const INDEX i, const INDEX j, const MATRIX& table) {
    need_open();
    if (i >= unaries_.size() || j >= unaries_.size() || i == j) throw std::runtime_error("AddPairwiseFactor: unknown or identical variables");
    const uint32_t d1 = label_count_[i], d2 = label_count_[j];
    for (uint32_t x = 0; x < d1; ++x)
      for (uint32_t y = 0; y < d2; ++y) tables_.push_back(table(x, y));
    ei_.push_back(i); ej_.push_back(j); reversed_.push_back(0);
    factors_.push_back({factor::pairwise, ei_.size() - 1});
    pairwise_.push_back(&factors_.back());
    return &factors_.back();
  }
  factor* GetUnaryFactor(const INDEX i) const { return unaries_.at(i); }
  factor* GetPairwiseFactor(const INDEX e) const { return pairwise_.at(e); }
  INDEX GetNumberOfVariables() const { return unaries_.size(); }
  INDEX GetNumberOfLabels(const INDEX i) const { return label_count_.at(i); }

  // AddFactorRelation(f1, f2): f1 is updated before f2 in the forward pass (include/LP_MP.h:697-702).  AddPairwiseFactor already
  // added u_i -> p -> u_j; a caller that wants the edge oriented the other way adds u_j -> p (and p -> u_i).
  void AddFactorRelation(factor* f1, factor* f2) {
    need_open();
    factor* u = f1->kind == factor::unary ? f1 : f2;
    factor* p = f1->kind == factor::pairwise ? f1 : f2;
    if (u->kind != factor::unary || p->kind != factor::pairwise) throw std::runtime_error("AddFactorRelation: needs one unary and one pairwise factor");
    const INDEX e = p->index;
    if (u->index != ei_[e] && u->index != ej_[e]) throw std::runtime_error("AddFactorRelation: the unary is not a variable of the pairwise factor");
    const bool unary_first = f1 == u;
    // u before p: u is the tail of the edge; p before u: u is its head
    const bool tail_is_i = unary_first ? u->index == ei_[e] : u->index == ej_[e];
    reversed_[e] = tail_is_i ? 0 : 1;
    custom_relations_ = custom_relations_ || !tail_is_i;
  }
  void ForwardPassFactorRelation(factor* f1, factor* f2) { AddFactorRelation(f1, f2); }
  void BackwardPassFactorRelation(factor* f1, factor* f2) { AddFactorRelation(f2, f1); }
  // include/LP_MP.h:465.  Only updated factors (unaries) make up a partition; a pairwise factor stands for the union of its variables
  void put_in_same_partition(factor* f1, factor* f2) {
    auto var = [&](factor* f) { return f->kind == factor::unary ? f->index : ei_[f->index]; };
    for (factor* f : {f1, f2}) if (f->kind == factor::pairwise) partition_pairs_.push_back({ei_[f->index], ej_[f->index]});
    partition_pairs_.push_back({var(f1), var(f2)});
    if (h_) push_partition_pairs();
  }
  void add_to_constant(const REAL x) { if (h_) ck(lpmp_add_to_constant(h_, x)); else constant_ += x; }   // :462

  INDEX GetNumberOfFactors() const { return factors_.size(); }
  factor* GetFactor(const INDEX i) { return &factors_.at(i); }
  INDEX GetNumberOfMessages() const { return 2 * ei_.size(); }

  // ---- the LP_TYPE surface ---------------------------------------------------------------------------------------------------
  void Begin() {   // include/LP_MP.h:704-728
    const std::string& t = reparametrization_type_arg_.getValue();
    if (t == "adaptive") throw std::runtime_error("LP_cuda: reparametrizationType adaptive is not available on the device (see lpmp_set_send_mode)");
    if (t != "shared" && t != "residual" && t != "partition" && t != "overlapping_partition") throw std::runtime_error("reparametrizationType " + t + " unknown");
    if (!h_) upload();
    ck(lpmp_begin(h_));
    ck(lpmp_set_send_mode(h_, t == "residual" ? LPMP_SEND_RESIDUAL : LPMP_SEND_SHARED));
    staged_ = t == "partition" ? 1 : (t == "overlapping_partition" ? 2 : 0);
  }
  void End() { ck(lpmp_end(h_)); }
  void set_reparametrization(const LPReparametrizationMode m) {   // include/config.hxx:72
    int mode;
    switch (m) {
      case LPReparametrizationMode::Anisotropic: mode = LPMP_REPAM_ANISOTROPIC; break;
      case LPReparametrizationMode::Anisotropic2: mode = LPMP_REPAM_ANISOTROPIC2; break;
      case LPReparametrizationMode::Uniform: mode = LPMP_REPAM_UNIFORM; break;
      case LPReparametrizationMode::DampedUniform: mode = LPMP_REPAM_DAMPED_UNIFORM; break;
      default: throw std::runtime_error("no reparametrization mode set");   // as include/LP_MP.h:458
    }
    ck(lpmp_set_reparametrization(need()->h_, mode));
  }
  void ComputePass(const INDEX iteration) {                                                                                 // :869-887
    need();
    if (staged_) ck(lpmp_compute_partition_pass(h_, inner_iteration_number_arg_.getValue(), staged_ == 2));
    if (staged_ != 1) ck(lpmp_compute_pass(h_, iteration));
  }
  void ComputeForwardPass() { ck(lpmp_compute_forward_pass(need()->h_)); }                                                   // :889-900
  void ComputeBackwardPass() { ck(lpmp_compute_backward_pass(need()->h_)); }                                                 // :902-911
  void ComputeForwardPassAndPrimal(const INDEX iteration) { ck(lpmp_compute_forward_pass_and_primal(need()->h_, iteration)); }   // :913-922
  void ComputeBackwardPassAndPrimal(const INDEX iteration) { ck(lpmp_compute_backward_pass_and_primal(need()->h_, iteration)); } // :924-933
  void ComputePassAndPrimal(const INDEX iteration) { ComputeForwardPassAndPrimal(iteration); ComputeBackwardPassAndPrimal(iteration); }
  REAL LowerBound() {                                                                                                       // :1507-1518
    if (!h_) upload();   // Solver::Solve prints the bound before Begin() in debug mode (solver.hxx:232-234)
    double v;
    ck(lpmp_lower_bound(h_, &v));
    return v;
  }
  REAL EvaluatePrimal() { double v; ck(lpmp_evaluate_primal(need()->h_, &v)); return v; }                                   // :1520-1536
  bool CheckPrimalConsistency() { int ok; ck(lpmp_check_primal_consistency(need()->h_, &ok)); return ok != 0; }             // :1066-1082

  // ---- state access (what the reference reads through GetFactor(i)->GetFactor()) ----------------------------------------------
  std::vector<uint32_t> labeling() { std::vector<uint32_t> l(unaries_.size()); ck(lpmp_get_labeling(need()->h_, l.data())); return l; }
  // dual variables in the layout of a factor_archive over all factors in insertion order (include/factor_archive.hxx:34-48,
  // include/serialization.hxx:226-322): unary -> theta (L_i doubles), pairwise -> msg1 (L_i doubles), msg2 (L_j doubles)
  INDEX dual_size_in_bytes() { uint64_t n; ck(lpmp_dual_size(need()->h_, &n)); return n; }
  void save_dual(void* archive) { ck(lpmp_save_dual(need()->h_, archive, dual_size_in_bytes())); }
  void load_dual(const void* archive) { ck(lpmp_load_dual(need()->h_, archive, dual_size_in_bytes())); }
  lpmp_solver* handle() { return need()->h_; }

private:
  LP_cuda* need() { if (!h_) throw std::runtime_error("LP_cuda: Begin() has not been called"); return this; }
  void need_open() const { if (h_) throw std::runtime_error("LP_cuda: the problem cannot be changed after Begin()"); }
  void ck(const int status) const { if (status != LPMP_OK) throw std::runtime_error(std::string("liblpmp_b200: ") + (h_ ? lpmp_last_error(h_) : "no handle")); }

  // Begin() of a fresh problem: SortFactors / weight computation happen on the other side of the boundary (lpmp_finalize)
  void upload() {
    const std::string& prec = precision_arg_.getValue();
    if (prec != "f64" && prec != "f32") throw std::runtime_error("cudaPrecision must be f64 or f32");
    if (lpmp_create(prec == "f64" ? LPMP_F64 : LPMP_F32, (int)device_arg_.getValue(), &h_) != LPMP_OK) {
      h_ = nullptr;
      throw std::runtime_error("liblpmp_b200: no CUDA device (there is no CPU fallback)");
    }
    const uint64_t n = unaries_.size(), m = ei_.size();
    ck(lpmp_add_unaries_var(h_, n, label_count_.data(), theta_.data()));
    std::vector<uint64_t> a(ei_.begin(), ei_.end()), b(ej_.begin(), ej_.end());
    if (m) ck(lpmp_add_pairwise_dense(h_, m, a.data(), b.data(), tables_.data()));
    if (custom_relations_) {
      // the relations orient every edge; any topological order of the variables is a valid rank
      std::vector<std::vector<INDEX>> out(n);
      std::vector<INDEX> indeg(n, 0);
      for (uint64_t e = 0; e < m; ++e) {
        const INDEX s = reversed_[e] ? ej_[e] : ei_[e], t = reversed_[e] ? ei_[e] : ej_[e];
        out[s].push_back(t); ++indeg[t];
      }
      std::priority_queue<INDEX, std::vector<INDEX>, std::greater<INDEX>> ready;
      for (INDEX i = 0; i < n; ++i) if (indeg[i] == 0) ready.push(i);
      std::vector<uint64_t> rank(n);
      uint64_t next = 0;
      while (!ready.empty()) {
        const INDEX i = ready.top(); ready.pop();
        rank[i] = next++;
        for (const INDEX t : out[i]) if (--indeg[t] == 0) ready.push(t);
      }
      if (next != n) throw std::runtime_error("factor relations do not form a DAG");   // include/topological_sort.hxx:125-127
      ck(lpmp_set_node_ranks(h_, rank.data()));
    }
    if (constant_ != 0.0) ck(lpmp_add_to_constant(h_, constant_));
    ck(lpmp_finalize(h_));
    push_partition_pairs();
    theta_.clear(); theta_.shrink_to_fit(); tables_.clear(); tables_.shrink_to_fit();
  }

  void push_partition_pairs() {
    if (partition_pairs_.empty()) return;
    std::vector<uint64_t> a, b;
    for (const auto& p : partition_pairs_) { a.push_back(p[0]); b.push_back(p[1]); }
    ck(lpmp_put_in_same_partition(h_, a.size(), a.data(), b.data()));
    partition_pairs_.clear();
  }

  TCLAP::ValueArg<std::string> reparametrization_type_arg_;
  TCLAP::ValueArg<INDEX> inner_iteration_number_arg_;
  TCLAP::ValueArg<INDEX> device_arg_;
  TCLAP::ValueArg<std::string> precision_arg_;

  lpmp_solver* h_ = nullptr;
  std::deque<factor> factors_;               // insertion order = the reference's f_ (include/LP_MP.h:214)
  std::vector<factor*> unaries_, pairwise_;
  std::vector<uint32_t> label_count_;
  std::vector<double> theta_, tables_;       // host staging until Begin()
  std::vector<INDEX> ei_, ej_;
  std::vector<unsigned char> reversed_;
  std::vector<std::array<INDEX, 2>> partition_pairs_;
  int staged_ = 0;   // 1: partition, 2: overlapping_partition (--reparametrizationType)
  bool custom_relations_ = false;
  REAL constant_ = 0.0;
};

}  // namespace LP_MP

#endif  // LP_MP_LP_CUDA_HXX

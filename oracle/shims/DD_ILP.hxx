// TEST INFRASTRUCTURE ONLY (oracle/): inert stand-in for pawelswoboda/DD_ILP.
// The reference includes "DD_ILP.hxx" (include/LP_MP.h:29, include/factors_messages.hxx:36) only so
// that the virtual ILP-export members of FactorTypeAdapter/MessageTypeAdapter instantiate
// (include/LP_MP.h:104-110,170-171; include/factors_messages.hxx:1604-1666,3441-3613).
// Exporting to an external ILP solver is out of scope of the message-passing hot path, so every
// operation here is a no-op. It is our own code, not a copy of DD_ILP.
#pragma once
#include <cstddef>
#include <string>

namespace DD_ILP {

struct variable_counters {
  std::size_t variables = 0, vectors = 0, matrices = 0, tensors = 0;
};

struct sat_solver {};
struct problem_export {};

template <typename BASE>
class external_solver_interface : public BASE {
public:
  struct variable {};
  struct vector {
    std::size_t n = 0;
    variable operator[](std::size_t) const { return {}; }
    std::size_t size() const { return n; }
    const variable* begin() const { return nullptr; }
    const variable* end() const { return nullptr; }
  };
  struct matrix {
    variable operator()(std::size_t, std::size_t) const { return {}; }
    std::size_t dim1() const { return 0; }
    std::size_t dim2() const { return 0; }
  };
  struct tensor {
    variable operator()(std::size_t, std::size_t, std::size_t) const { return {}; }
    std::size_t dim1() const { return 0; }
    std::size_t dim2() const { return 0; }
    std::size_t dim3() const { return 0; }
  };

  variable add_variable() { return {}; }
  template <typename V> vector add_vector(const V&) { return {}; }
  template <typename M> matrix add_matrix(const M&) { return {}; }
  template <typename T> tensor add_tensor(const T&) { return {}; }

  variable load_variable() { return {}; }
  vector load_vector() { return {}; }
  matrix load_matrix() { return {}; }
  tensor load_tensor() { return {}; }

  template <typename X> void add_variable_objective(const X&) {}
  template <typename X> void add_vector_objective(const X&) {}
  template <typename X> void add_matrix_objective(const X&) {}
  template <typename X> void add_tensor_objective(const X&) {}

  variable_counters get_variable_counters() const { return counters_; }
  void set_variable_counters(const variable_counters& c) { counters_ = c; }
  void init_variable_loading() {}

  template <typename IT> void add_simplex_constraint(IT, IT) {}
  template <typename IT> void add_at_most_one_constraint(IT, IT) {}
  void make_equal(variable, variable) {}
  template <typename IT> void make_equal(IT, IT, IT, IT) {}
  bool solution(variable) const { return false; }
  template <typename IT> std::size_t first_active(IT, IT) const { return 0; }

  bool solve() { return false; }
  void write_to_file(const std::string&) {}

private:
  variable_counters counters_;
};

}  // namespace DD_ILP

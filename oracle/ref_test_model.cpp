// TEST INFRASTRUCTURE ONLY (oracle/).  Golden (6) of SURVEY.md 8c: the reference's own three-factor test
// model (/root/reference/test/test_model.hxx:10-137, included from where it lies) driven exactly as
// test/test_model.cpp:19-46 does, minus the LP_external_solver wrapper (it does not compile at this commit:
// include/LP_external_interface.hxx:65).  Proves that the shim-compiled reference core behaves like the
// reference: LowerBound() must reach 1.0.  Exit code 0 on success.
#include "config.hxx"
#include "factors_messages.hxx"
#include "solver.hxx"
#include "visitors/standard_visitor.hxx"
#include "test.h"
#include "test_model.hxx"

using namespace LP_MP;

int main() {
  std::vector<std::string> options = {"test solver", "--maxIter", "5", "-v", "0"};
  Solver<LP<test_FMC>, StandardVisitor> s(options);
  auto& lp = s.GetLP();
  auto* f1 = lp.template add_factor<typename test_FMC::factor>(0, 1);
  auto* f2 = lp.template add_factor<typename test_FMC::factor>(1, 0);
  auto* f3 = lp.template add_factor<typename test_FMC::factor>(0, 0);
  lp.template add_message<typename test_FMC::message>(f1, f2);
  lp.template add_message<typename test_FMC::message>(f1, f3);
  test(lp.GetNumberOfFactors() == 3);
  test(lp.GetNumberOfMessages() == 2);
  test(lp.GetFactor(0) == f1 && lp.GetFactor(1) == f2 && lp.GetFactor(2) == f3);
  test(f1->no_messages() == 2 && f1->no_send_messages() == 2);
  test(f2->no_messages() == 1 && f2->no_send_messages() == 0);
  test(f3->no_messages() == 1 && f3->no_send_messages() == 0);
  s.Solve();
  test(std::abs(lp.LowerBound() - 1.0) <= eps);
  std::cout << "test_model lower bound = " << lp.LowerBound() << "\n";
  return 0;
}

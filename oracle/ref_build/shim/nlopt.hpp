// Test infrastructure (oracle/): NLopt is not installed.  The stand-in's optimize() always throws, which drives
// the reference into its own documented fall-back (ci.cpp:70-75,112-116,183-188: fixed weights).
#ifndef XREF_NLOPT_HPP
#define XREF_NLOPT_HPP
#include <stdexcept>
#include <vector>
namespace nlopt {
enum algorithm { LN_COBYLA = 25 };
enum result { FAILURE = -1, INVALID_ARGS = -2, OUT_OF_MEMORY = -3, ROUNDOFF_LIMITED = -4, FORCED_STOP = -5, SUCCESS = 1,
              STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4, MAXEVAL_REACHED = 5, MAXTIME_REACHED = 6 };
typedef double (*vfunc)(const std::vector<double>&, std::vector<double>&, void*);
class opt {
 public:
  opt(algorithm, unsigned) {}
  void set_lower_bounds(double) {}
  void set_upper_bounds(double) {}
  void set_maxtime(double) {}
  void set_min_objective(vfunc, void*) {}
  void add_inequality_constraint(vfunc, void*, double) {}
  void add_equality_constraint(vfunc, void*, double) {}
  void set_ftol_abs(double) {}
  void set_xtol_rel(double) {}
  result optimize(std::vector<double>&, double&) { throw std::runtime_error("nlopt stand-in: optimiser unavailable"); }
};
}  // namespace nlopt
#endif

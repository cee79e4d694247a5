// Test infrastructure (oracle/): chi-squared quantile with the call surface of boost::math (Boost is not
// installed).  quantile(chi_squared(k), p) solves P(k/2, x/2) = p for the regularised lower incomplete gamma
// function (series / continued fraction, Newton with bisection safeguard), to ~1e-14 relative.
#ifndef XREF_BOOST_MATH_DISTRIBUTIONS
#define XREF_BOOST_MATH_DISTRIBUTIONS
#include <cmath>
#include <map>
#include <utility>
namespace boost {
namespace math {
namespace xref_detail {
inline double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
inline double chi2_quantile(double k, double p) {
  static thread_local std::map<std::pair<double, double>, double> cache;
  auto it = cache.find({k, p});
  if (it != cache.end()) return it->second;
  const double a = 0.5 * k;
  // Wilson-Hilferty start
  double z = 0.0;
  {
    // inverse normal (Acklam) is overkill: bracket + bisection/Newton is enough
    double lo = 0.0, hi = std::fmax(4.0 * k, 16.0);
    while (gamma_p(a, 0.5 * hi) < p) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    for (int it2 = 0; it2 < 200; ++it2) {
      const double f = gamma_p(a, 0.5 * x) - p;
      if (f > 0) hi = x; else lo = x;
      const double pdf = 0.5 * std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a));
      double xn = x - f / pdf;
      if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
      if (std::fabs(xn - x) <= 1e-15 * std::fabs(x)) {
        x = xn;
        break;
      }
      x = xn;
    }
    z = x;
  }
  cache[{k, p}] = z;
  return z;
}
}  // namespace xref_detail
template <typename T = double> class chi_squared_distribution {
  T k_;

 public:
  explicit chi_squared_distribution(T k) : k_(k) {}
  T degrees_of_freedom() const { return k_; }
};
typedef chi_squared_distribution<double> chi_squared;
template <typename T> inline T quantile(const chi_squared_distribution<T>& d, const double& p) {
  return T(xref_detail::chi2_quantile(double(d.degrees_of_freedom()), p));
}
}  // namespace math
}  // namespace boost
#endif

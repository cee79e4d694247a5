// Test infrastructure (oracle/): optional BLAS/LAPACK back end for the Eigen stand-in.  The harness binds the
// ILP64 OpenBLAS that ships with numpy (symbols scipy_*_64_) through dlopen at run time; without it the stand-in
// uses its own loops.  Never part of the product path.
#ifndef XREF_BLAS_H
#define XREF_BLAS_H
#include <cstddef>
#include <cstdint>
#include <vector>

namespace xref {
typedef int64_t blasint;
struct Blas {
  void (*dgemm)(const char*, const char*, const blasint*, const blasint*, const blasint*, const double*, const double*,
                const blasint*, const double*, const blasint*, const double*, double*, const blasint*) = nullptr;
  void (*dgeqrf)(const blasint*, const blasint*, double*, const blasint*, double*, double*, const blasint*,
                 blasint*) = nullptr;
  void (*dgetrf)(const blasint*, const blasint*, double*, const blasint*, blasint*, blasint*) = nullptr;
  void (*dgetri)(const blasint*, double*, const blasint*, const blasint*, double*, const blasint*, blasint*) = nullptr;
  void (*set_threads)(int) = nullptr;
};
inline Blas& blas() {
  static Blas b;
  return b;
}
inline void blas_gemm(std::ptrdiff_t m, std::ptrdiff_t n, std::ptrdiff_t k, const double* A, std::ptrdiff_t lda,
                      const double* B, std::ptrdiff_t ldb, double* C, std::ptrdiff_t ldc) {
  const blasint M = m, N = n, K = k, LA = lda, LB = ldb, LC = ldc;
  const double one = 1.0, zero = 0.0;
  blas().dgemm("N", "N", &M, &N, &K, &one, A, &LA, B, &LB, &zero, C, &LC);
}
inline void blas_geqrf(std::ptrdiff_t m, std::ptrdiff_t n, double* A, std::ptrdiff_t lda, double* tau) {
  const blasint M = m, N = n, LA = lda;
  blasint info = 0, lwork = -1;
  double wq = 0;
  blas().dgeqrf(&M, &N, A, &LA, tau, &wq, &lwork, &info);
  lwork = blasint(wq) + 1;
  std::vector<double> work(static_cast<size_t>(lwork));
  blas().dgeqrf(&M, &N, A, &LA, tau, work.data(), &lwork, &info);
}
inline void blas_inverse(std::ptrdiff_t n, double* A) {
  const blasint N = n;
  blasint info = 0, lwork = -1;
  std::vector<blasint> piv(static_cast<size_t>(n));
  blas().dgetrf(&N, &N, A, &N, piv.data(), &info);
  double wq = 0;
  blas().dgetri(&N, A, &N, piv.data(), &wq, &lwork, &info);
  lwork = blasint(wq) + 1;
  std::vector<double> work(static_cast<size_t>(lwork));
  blas().dgetri(&N, A, &N, piv.data(), work.data(), &lwork, &info);
}
}  // namespace xref
#endif

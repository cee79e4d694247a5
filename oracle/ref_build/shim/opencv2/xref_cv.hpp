// Test infrastructure (oracle/): stand-in for the few OpenCV types the reference's filter back end touches
// (cv::Mat as a small double matrix, cv::triangulatePoints as the two-view DLT it documents, no-op drawing).
// Written from scratch; OpenCV's C++ headers are not installed in this container.
#ifndef XREF_CV_HPP
#define XREF_CV_HPP
#include <cmath>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32F 5
#define CV_64F 6
#define CV_64FC1 6

namespace cv {
template <typename T> struct Point_ {
  T x{}, y{};
  Point_() {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x(T(o.x)), y(T(o.y)) {}
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
struct Scalar {
  double v[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {}
};
struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};
struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
};

template <typename T> class MatCommaInitializer_;

class Mat {
 public:
  int rows = 0, cols = 0;
  std::shared_ptr<std::vector<double>> d;  // row-major, shared like cv::Mat headers
  Mat() {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), d(std::make_shared<std::vector<double>>(size_t(r) * c, 0.0)) {}
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  Mat clone() const {
    Mat m;
    m.rows = rows;
    m.cols = cols;
    if (d) m.d = std::make_shared<std::vector<double>>(*d);
    return m;
  }
  bool empty() const { return rows == 0 || cols == 0 || !d; }
  int type() const { return CV_64F; }
  Size size() const { return Size(cols, rows); }
  template <typename T> T& at(int i) { return reinterpret_cast<T&>((*d)[size_t(i)]); }
  template <typename T> const T& at(int i) const { return reinterpret_cast<const T&>((*d)[size_t(i)]); }
  template <typename T> T& at(int i, int j) { return reinterpret_cast<T&>((*d)[size_t(i) * cols + j]); }
  template <typename T> const T& at(int i, int j) const {
    return reinterpret_cast<const T&>((*d)[size_t(i) * cols + j]);
  }
};
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat c(a.rows, b.cols, CV_64F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0;
      for (int k = 0; k < a.cols; ++k) s += a.at<double>(i, k) * b.at<double>(k, j);
      c.at<double>(i, j) = s;
    }
  return c;
}

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_64F) {}
  MatCommaInitializer_<T> operator<<(T v);
};
template <typename T> class MatCommaInitializer_ {
  Mat_<T> m_;
  size_t n_ = 0;

 public:
  MatCommaInitializer_(const Mat_<T>& m, T v) : m_(m) { (*m_.d)[n_++] = double(v); }
  template <typename U> MatCommaInitializer_& operator,(U v) {
    (*m_.d)[n_++] = double(v);
    return *this;
  }
  operator Mat() const { return m_; }
  operator Mat_<T>() const { return m_; }
};
template <typename T> MatCommaInitializer_<T> Mat_<T>::operator<<(T v) { return MatCommaInitializer_<T>(*this, v); }

// Delaunay lookup of TrackManager::featureTriangleAtPoint (range-finder facet, out of scope): the type exists so that the
// reference's track_manager.cpp compiles unmodified; locate() reports "outside", so the caller gets no facet.
struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
  Rect() {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
  template <typename T> bool contains(const Point_<T>& p) const { return p.x >= x && p.y >= y && p.x < x + width && p.y < y + height; }
};
template <typename T, int N> struct Vec {
  T v[N] = {};
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
typedef Vec<float, 6> Vec6f;
typedef Vec<float, 4> Vec4f;
class Subdiv2D {
 public:
  enum { PTLOC_ERROR = -2, PTLOC_OUTSIDE_RECT = -1, PTLOC_INSIDE = 0, PTLOC_VERTEX = 1, PTLOC_ON_EDGE = 2 };
  Subdiv2D() {}
  explicit Subdiv2D(Rect) {}
  template <typename T> int insert(const Point_<T>&) { return 0; }
  template <typename T> int locate(const Point_<T>&, int& edge, int& vertex) { edge = 0; vertex = 0; return PTLOC_OUTSIDE_RECT; }
  int edgeOrg(int, Point2f* = nullptr) const { return 0; }
  int edgeDst(int, Point2f* = nullptr) const { return 0; }
  int nextEdge(int) const { return 0; }
  Point2f getVertex(int, int* first_edge = nullptr) const { if (first_edge) *first_edge = 0; return Point2f(); }
  void getTriangleList(std::vector<Vec6f>& t) const { t.clear(); }
};

// drawing: no-ops
template <typename... A> inline void rectangle(A&&...) {}
template <typename... A> inline void line(A&&...) {}
template <typename... A> inline void circle(A&&...) {}
template <typename... A> inline void putText(A&&...) {}
template <typename... A> inline void cvtColor(A&&...) {}
template <typename... A> inline Size getTextSize(A&&...) { return Size(); }
enum { FONT_HERSHEY_SIMPLEX = 0, FONT_HERSHEY_PLAIN = 1, LINE_AA = 16, LINE_8 = 8 };

// Two-view linear triangulation (the DLT cv::triangulatePoints documents): for every point the homogeneous
// solution is the right singular vector of the 4x4 design matrix with the smallest singular value.  One-sided
// Jacobi on the columns of A.
inline void triangulatePoints(const Mat& P1, const Mat& P2, const Mat& x1, const Mat& x2, Mat& out) {
  const int n = x1.cols;
  out = Mat(4, n, CV_64F);
  for (int p = 0; p < n; ++p) {
    double A[4][4], V[4][4];
    const Mat* P[2] = {&P1, &P2};
    const Mat* X[2] = {&x1, &x2};
    for (int v = 0; v < 2; ++v) {
      const double x = X[v]->at<double>(0, p), y = X[v]->at<double>(1, p);
      for (int k = 0; k < 4; ++k) {
        A[2 * v][k] = x * P[v]->at<double>(2, k) - P[v]->at<double>(0, k);
        A[2 * v + 1][k] = y * P[v]->at<double>(2, k) - P[v]->at<double>(1, k);
      }
    }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = 0;
      for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 4; ++b) {
          double al = 0, be = 0, ga = 0;
          for (int i = 0; i < 4; ++i) {
            al += A[i][a] * A[i][a];
            be += A[i][b] * A[i][b];
            ga += A[i][a] * A[i][b];
          }
          if (ga == 0.0 || std::fabs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
          off = std::fmax(off, std::fabs(ga) / std::sqrt(al * be));
          const double zeta = (be - al) / (2.0 * ga);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
          for (int i = 0; i < 4; ++i) {
            const double u = A[i][a], w = A[i][b];
            A[i][a] = c * u - s * w;
            A[i][b] = s * u + c * w;
            const double vu = V[i][a], vw = V[i][b];
            V[i][a] = c * vu - s * vw;
            V[i][b] = s * vu + c * vw;
          }
        }
      if (off < 1e-15) break;
    }
    int best = 0;
    double bn = -1;
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int i = 0; i < 4; ++i) s += A[i][j] * A[i][j];
      if (bn < 0 || s < bn) {
        bn = s;
        best = j;
      }
    }
    for (int i = 0; i < 4; ++i) out.at<double>(i, p) = V[i][best];
  }
}
}  // namespace cv
inline int cvRound(double x) { return (int)std::lround(x); }
#ifndef CV_AA
#define CV_AA 16
#endif
#ifndef CV_GRAY2BGR
#define CV_GRAY2BGR 8
#endif
#endif

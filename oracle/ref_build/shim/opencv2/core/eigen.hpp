// Test infrastructure (oracle/): cv::eigen2cv for the stand-ins.
#ifndef XREF_CV_EIGEN_HPP
#define XREF_CV_EIGEN_HPP
#include <Eigen/Core>
#include "../xref_cv.hpp"
namespace cv {
template <typename D> void eigen2cv(const Eigen::MatrixBase<D>& src, Mat& dst) {
  dst = Mat(int(src.rows()), int(src.cols()), CV_64F);
  for (int i = 0; i < dst.rows; ++i)
    for (int j = 0; j < dst.cols; ++j) dst.at<double>(i, j) = src.coeff(i, j);
}
}  // namespace cv
#endif

// opencv2/opencv.hpp -- minimal stand-in (test infrastructure).  TSDFGrid::Integrate
// (reference utils/tsdf/voxel_tsdf.cu:416-441) reads only rows / cols / type() / total() / data of a
// continuous cv::Mat; the OpenCV C++ headers are absent from this image.
#ifndef RSF_REF_SHIM_OPENCV_
#define RSF_REF_SHIM_OPENCV_
#include <cstddef>
#include <vector>
#include <string>
#define CV_8UC3 16
#define CV_32FC1 5
namespace cv {
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), type_(0) {}
  Mat(int rows_, int cols_, int type, void* data_) : rows(rows_), cols(cols_), data(static_cast<unsigned char*>(data_)), type_(type) {}
  int type() const { return type_; }
  size_t total() const { return static_cast<size_t>(rows) * cols; }
  bool isContinuous() const { return true; }
  int rows, cols;
  unsigned char* data;
 private:
  int type_;
};
}  // namespace cv
#endif

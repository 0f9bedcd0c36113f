// opencv2/opencv.hpp -- minimal stand-in (test infrastructure).  TSDFGrid::Integrate
// (reference utils/tsdf/voxel_tsdf.cu:416-441) reads only rows / cols / type() / total() / data of a
// continuous cv::Mat; modules/tsdf_module.cc:22-37 additionally uses empty(), clone() and Mat::ones().
// The OpenCV C++ headers are absent from this image.
#ifndef RSF_REF_SHIM_OPENCV_
#define RSF_REF_SHIM_OPENCV_
#include <cstddef>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#define CV_8UC3 16
#define CV_32FC1 5
namespace cv {
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), type_(0) {}
  Mat(int rows_, int cols_, int type, void* data_) : rows(rows_), cols(cols_), data(static_cast<unsigned char*>(data_)), type_(type) {}
  Mat(int rows_, int cols_, int type) : rows(rows_), cols(cols_), data(nullptr), type_(type) {
    own_ = std::shared_ptr<unsigned char>(new unsigned char[bytes() ? bytes() : 1], std::default_delete<unsigned char[]>());
    data = own_.get();
  }
  int type() const { return type_; }
  size_t total() const { return static_cast<size_t>(rows) * cols; }
  size_t elemSize() const { return type_ == CV_8UC3 ? 3 : 4; }
  bool isContinuous() const { return true; }
  bool empty() const { return data == nullptr || total() == 0; }
  Mat clone() const {
    if (empty()) return Mat();
    Mat m(rows, cols, type_);
    std::memcpy(m.data, data, bytes());
    return m;
  }
  static Mat ones(int rows_, int cols_, int type) {   // CV_32FC1 only (what tsdf_module.cc asks for)
    Mat m(rows_, cols_, type);
    float* f = reinterpret_cast<float*>(m.data);
    for (size_t i = 0; i < m.total(); ++i) f[i] = 1.f;
    return m;
  }
  int rows, cols;
  unsigned char* data;

 private:
  size_t bytes() const { return total() * elemSize(); }
  int type_;
  std::shared_ptr<unsigned char> own_;
};
}  // namespace cv
#endif

// rsfusion/tsdf_grid.hpp -- header-only C++ adapter: the reference's class and method names on
// top of the C ABI (include/rsfusion.h).
//
// Mirrors, with the same names, argument meaning and order:
//   CameraIntrinsics<T>, CameraParams        utils/cuda/camera.cuh:13-68
//   SE3<T>                                   utils/cuda/lie_group.cuh:8-45
//   BoundingCube<T>                          utils/tsdf/voxel_tsdf.cuh:20-34
//   VoxelSpatialTSDF, VoxelSpatialTSDFSEGM   utils/tsdf/voxel_types.cuh:48-70
//   TSDFGrid                                 utils/tsdf/voxel_tsdf.cuh:39-107
//
// Differences a maintainer must know:
//   * images are passed as rsfusion::Image views (data, rows, cols, type) -- define
//     RSFUSION_WITH_OPENCV before including this header to also get the cv::Mat overloads;
//     RSFUSION_WITH_EIGEN adds conversions from Eigen matrices / quaternions.
//   * RayCast writes into caller-provided host buffers instead of GLImage8UC4 textures, and can
//     also return depth and probability images.
//   * failures throw rsfusion::Error (the reference asserts or silently continues).
//   * define RSFUSION_GLOBAL_NAMES to import the names into the global namespace, which is where
//     the reference keeps them, so that TSDFSystem / offline_eval sources recompile unchanged.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rsfusion.h"

#ifdef RSFUSION_WITH_EIGEN
#include <Eigen/Dense>
#include <Eigen/Geometry>
#endif
#ifdef RSFUSION_WITH_OPENCV
#ifndef CV_8UC3          // (an overlay build has included <opencv2/opencv.hpp> already)
#include <opencv2/core.hpp>
#endif
#endif

namespace rsfusion {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("rsfusion error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
  if (rc != RSF_OK) throw Error(rc, rsf_last_error());
}

// ---- utils/cuda/camera.cuh ---------------------------------------------------------------------
template <typename T>
struct CameraIntrinsics {
  T fx, fy, cx, cy;
  CameraIntrinsics(const T& fx_, const T& fy_, const T& cx_, const T& cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}
  CameraIntrinsics<T> Inverse() const {
    const T fx_inv = 1 / fx, fy_inv = 1 / fy;
    return CameraIntrinsics<T>(fx_inv, fy_inv, -cx * fx_inv, -cy * fy_inv);
  }
  std::array<T, 3> operator*(const std::array<T, 3>& v) const { return {fx * v[0] + cx * v[2], fy * v[1] + cy * v[2], v[2]}; }
};

class CameraParams {
 public:
  CameraIntrinsics<float> intrinsics;
  CameraIntrinsics<float> intrinsics_inv;
  int img_h;
  int img_w;
  CameraParams(const CameraIntrinsics<float>& intrinsics_, int img_h_, int img_w_)
      : intrinsics(intrinsics_), intrinsics_inv(intrinsics_.Inverse()), img_h(img_h_), img_w(img_w_) {}
};

// ---- utils/cuda/lie_group.cuh ------------------------------------------------------------------
// Quaternion (w, x, y, z) + translation.  Float operations follow Eigen 3.3's order (the reference
// stores an Eigen::Quaternion): q*v = v + w*uv + q.vec x uv, uv = 2 (q.vec x v); Quaternion(Matrix3)
// is the trace-based conversion without normalisation.
template <typename T>
class SE3 {
 public:
  SE3() : q_{1, 0, 0, 0}, t_{0, 0, 0} {}
  SE3(const std::array<T, 4>& quat_wxyz, const std::array<T, 3>& trans) : q_(quat_wxyz), t_(trans) {}
  // row-major 4x4 or 3x4 [R | t]
  explicit SE3(const T* mat, int row_stride = 4) {
    T R[3][3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) R[i][j] = mat[i * row_stride + j];
      t_[i] = mat[i * row_stride + 3];
    }
    const T tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > T(0)) {
      T s = std::sqrt(tr + T(1));
      q_[0] = T(0.5) * s;
      s = T(0.5) / s;
      q_[1] = (R[2][1] - R[1][2]) * s;
      q_[2] = (R[0][2] - R[2][0]) * s;
      q_[3] = (R[1][0] - R[0][1]) * s;
    } else {
      int i = 0;
      if (R[1][1] > R[0][0]) i = 1;
      if (R[2][2] > R[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      T s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + T(1));
      q_[1 + i] = T(0.5) * s;
      s = T(0.5) / s;
      q_[0] = (R[k][j] - R[j][k]) * s;
      q_[1 + j] = (R[j][i] + R[i][j]) * s;
      q_[1 + k] = (R[k][i] + R[i][k]) * s;
    }
  }
#ifdef RSFUSION_WITH_EIGEN
  SE3(const Eigen::Quaternion<T>& rot, const Eigen::Matrix<T, 3, 1>& trans)
      : q_{rot.w(), rot.x(), rot.y(), rot.z()}, t_{trans[0], trans[1], trans[2]} {}
  explicit SE3(const Eigen::Matrix<T, 4, 4>& m) : SE3(Eigen::Quaternion<T>(m.template topLeftCorner<3, 3>()), m.template topRightCorner<3, 1>()) {}
  explicit SE3(const Eigen::Matrix<T, 3, 4>& m) : SE3(Eigen::Quaternion<T>(m.template topLeftCorner<3, 3>()), m.template topRightCorner<3, 1>()) {}
#endif
  static SE3<T> Identity() { return SE3<T>(); }

  SE3<T> Inverse() const {
    const T n2 = (q_[1] * q_[1] + q_[2] * q_[2]) + (q_[3] * q_[3] + q_[0] * q_[0]);  // Eigen redux order (x,y | z,w)
    std::array<T, 4> qi{0, 0, 0, 0};
    if (n2 > T(0)) qi = {q_[0] / n2, -q_[1] / n2, -q_[2] / n2, -q_[3] / n2};
    return SE3<T>(qi, rotate(qi, {-t_[0], -t_[1], -t_[2]}));
  }
  std::array<T, 4> GetR() const { return q_; }
  std::array<T, 3> GetT() const { return t_; }
  std::array<T, 3> Apply(const std::array<T, 3>& v) const {
    const auto r = rotate(q_, v);
    return {r[0] + t_[0], r[1] + t_[1], r[2] + t_[2]};
  }
  SE3<T> operator*(const SE3<T>& o) const {
    const auto &a = q_, &b = o.q_;   // Hamilton product
    const std::array<T, 4> q{a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                             a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                             a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                             a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]};
    return SE3<T>(q, Apply(o.t_));
  }
  // {qw,qx,qy,qz,tx,ty,tz}: the `pose` argument of the C ABI
  std::array<float, 7> as_pose() const {
    return {(float)q_[0], (float)q_[1], (float)q_[2], (float)q_[3], (float)t_[0], (float)t_[1], (float)t_[2]};
  }

 private:
  static std::array<T, 3> rotate(const std::array<T, 4>& q, const std::array<T, 3>& v) {
    const T ux = q[2] * v[2] - q[3] * v[1], uy = q[3] * v[0] - q[1] * v[2], uz = q[1] * v[1] - q[2] * v[0];
    const T uvx = ux + ux, uvy = uy + uy, uvz = uz + uz;
    const T cx = q[2] * uvz - q[3] * uvy, cy = q[3] * uvx - q[1] * uvz, cz = q[1] * uvy - q[2] * uvx;
    return {(v[0] + q[0] * uvx) + cx, (v[1] + q[0] * uvy) + cy, (v[2] + q[0] * uvz) + cz};
  }
  std::array<T, 4> q_;
  std::array<T, 3> t_;
};

// ---- utils/tsdf/voxel_tsdf.cuh:20-34 -----------------------------------------------------------
template <typename T>
struct BoundingCube {
  T xmin, xmax, ymin, ymax, zmin, zmax;
};

// ---- utils/tsdf/voxel_types.cuh:48-70 (bit-compatible row formats of the exports) ---------------
struct VoxelSpatialTSDF {
  float position[3];
  float tsdf;
};
struct VoxelSpatialTSDFSEGM {
  float position[3];
  float tsdf;
  float probability;
};
static_assert(sizeof(VoxelSpatialTSDF) == 16 && sizeof(VoxelSpatialTSDFSEGM) == 20, "row formats of the reference");

// cv::Mat stand-in: type uses OpenCV's numeric values so the reference's asserts read the same
constexpr int kType8UC3 = 16;    // CV_8UC3
constexpr int kType32FC1 = 5;    // CV_32FC1
struct Image {
  const void* data = nullptr;
  int rows = 0, cols = 0, type = 0;
  Image() = default;
  Image(const void* d, int r, int c, int t) : data(d), rows(r), cols(c), type(t) {}
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
#ifdef RSFUSION_WITH_OPENCV
  Image(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), type(m.type()) {  // NOLINT (implicit on purpose)
    if (!m.empty() && !m.isContinuous()) throw Error(RSF_ERR_INVALID, "cv::Mat must be continuous");
  }
#endif
};

// ---- utils/tsdf/voxel_tsdf.cuh:39-107 ----------------------------------------------------------
class TSDFGrid {
 public:
  TSDFGrid(float voxel_size, float truncation, const rsf_options* extra = nullptr) {
    rsf_options o;
    rsf_default_options(&o);
    if (extra) o = *extra;
    o.voxel_size = voxel_size;
    o.truncation = truncation;
    check(rsf_create(&o, &g_));
  }
  ~TSDFGrid() { rsf_destroy(g_); }
  TSDFGrid(const TSDFGrid&) = delete;
  TSDFGrid& operator=(const TSDFGrid&) = delete;

  // voxel_tsdf.cu:416-452 (same type / size contract as the asserts at :419-428)
  void Integrate(const Image& img_rgb, const Image& img_depth, const Image& img_ht, const Image& img_lt, float max_depth,
                 const CameraIntrinsics<float>& intrinsics, const SE3<float>& cam_T_world) {
    validate(img_rgb, img_depth, img_ht, img_lt);
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    check(rsf_integrate(g_, (const uint8_t*)img_rgb.data, (const float*)img_depth.data, (const float*)img_ht.data,
                        (const float*)img_lt.data, img_depth.rows, img_depth.cols, max_depth, K, pose.data()));
  }
  // Pipelined form (rsf_integrate_async): the buffers must stay valid until Synchronize().
  void IntegrateAsync(const Image& img_rgb, const Image& img_depth, const Image& img_ht, const Image& img_lt,
                      float max_depth, const CameraIntrinsics<float>& intrinsics, const SE3<float>& cam_T_world) {
    validate(img_rgb, img_depth, img_ht, img_lt);
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    check(rsf_integrate_async(g_, (const uint8_t*)img_rgb.data, (const float*)img_depth.data, (const float*)img_ht.data,
                              (const float*)img_lt.data, img_depth.rows, img_depth.cols, max_depth, K, pose.data()));
  }
  // Raw 16-bit images (rsf_integrate_u16 / _async): what the reference's callers convert on the CPU before Integrate
  // (examples/tsdf/offline.cc:62-76).  img_ht / img_lt may both be null pointers ("no segmentation": ht 0, lt 1).
  void IntegrateU16(const uint8_t* img_rgb, const uint16_t* img_depth, const uint16_t* img_ht, const uint16_t* img_lt, int rows,
                    int cols, float depth_factor, float max_depth, const CameraIntrinsics<float>& intrinsics,
                    const SE3<float>& cam_T_world, bool asynchronous = false) {
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    check((asynchronous ? rsf_integrate_u16_async : rsf_integrate_u16)(g_, img_rgb, img_depth, img_ht, img_lt, rows, cols,
                                                                        depth_factor, max_depth, K, pose.data()));
  }
  // K-class semantic extension (grid created with rsf_options.num_classes = K > 2): class_prob = K planes of
  // rows*cols float probabilities; rows of 4 + K floats (x, y, z, tsdf, p_0 .. p_{K-1}) on export.
  void IntegrateClasses(const uint8_t* img_rgb, const float* img_depth, const float* class_prob, int rows, int cols, float max_depth,
                        const CameraIntrinsics<float>& intrinsics, const SE3<float>& cam_T_world) {
    const float K[4] = {intrinsics.fx, intrinsics.fy, intrinsics.cx, intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    check(rsf_integrate_classes(g_, img_rgb, img_depth, class_prob, rows, cols, max_depth, K, pose.data()));
  }
  std::vector<float> GatherValidClasses(int num_classes) {
    int64_t n = 0;
    check(rsf_gather_valid_classes(g_, nullptr, 0, &n));
    std::vector<float> rows((size_t)n * (4 + num_classes));
    if (n) check(rsf_gather_valid_classes(g_, rows.data(), n, &n));
    return rows;
  }
  void Synchronize() { check(rsf_synchronize(g_)); }

#ifdef RSFUSION_WITH_GL
  // The reference's signature (voxel_tsdf.cuh:78-79): render into the grid's device images and hand them to the OpenGL
  // bridge, GLImage8UC4::LoadCuda(device pointer, stream), exactly as voxel_tsdf.cu:895-901 does; then synchronise.
  // GLImage8UC4 comes from the caller's utils/gl/image.h, which must be included before this header.
  void RayCast(float max_depth, const CameraParams& virtual_cam, const SE3<float>& cam_T_world, GLImage8UC4* tsdf_rgba = NULL,
               GLImage8UC4* tsdf_normal = NULL) {
    const float K[4] = {virtual_cam.intrinsics.fx, virtual_cam.intrinsics.fy, virtual_cam.intrinsics.cx, virtual_cam.intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    const uint8_t *d_rgba = nullptr, *d_normal = nullptr;
    void* stream = nullptr;
    check(rsf_raycast_internal_device(g_, max_depth, virtual_cam.img_h, virtual_cam.img_w, K, pose.data(), &d_rgba, &d_normal, &stream));
    if (tsdf_rgba != NULL) tsdf_rgba->LoadCuda(d_rgba, (cudaStream_t)stream);
    if (tsdf_normal != NULL) tsdf_normal->LoadCuda(d_normal, (cudaStream_t)stream);
    check(rsf_synchronize(g_));
  }
  // host-buffer form (no defaults here, so that RayCast(max_depth, cam, pose) is the reference's call)
  void RayCast(float max_depth, const CameraParams& virtual_cam, const SE3<float>& cam_T_world, uint8_t* tsdf_rgba,
               uint8_t* tsdf_normal, float* depth = nullptr, float* prob = nullptr) {
#else
  // voxel_tsdf.cu:885-902.  rgba / normal: img_h*img_w*4 bytes each; depth / prob: img_h*img_w floats.
  void RayCast(float max_depth, const CameraParams& virtual_cam, const SE3<float>& cam_T_world, uint8_t* tsdf_rgba = nullptr,
               uint8_t* tsdf_normal = nullptr, float* depth = nullptr, float* prob = nullptr) {
#endif
    const float K[4] = {virtual_cam.intrinsics.fx, virtual_cam.intrinsics.fy, virtual_cam.intrinsics.cx, virtual_cam.intrinsics.cy};
    const auto pose = cam_T_world.as_pose();
    check(rsf_raycast(g_, max_depth, virtual_cam.img_h, virtual_cam.img_w, K, pose.data(), tsdf_rgba, tsdf_normal, depth, prob));
  }

  std::vector<VoxelSpatialTSDF> GatherValid() { return gather<VoxelSpatialTSDF>(0, nullptr); }                 // :476-502
  std::vector<VoxelSpatialTSDFSEGM> GatherValidSemantic() { return gather<VoxelSpatialTSDFSEGM>(1, nullptr); }  // :504-530
  std::vector<VoxelSpatialTSDF> GatherVoxels(const BoundingCube<float>& volumn) {                              // :532-559
    const float b[6] = {volumn.xmin, volumn.xmax, volumn.ymin, volumn.ymax, volumn.zmin, volumn.zmax};
    return gather<VoxelSpatialTSDF>(2, b);
  }
  // :736-845; vertices / indices as plain arrays (Eigen::Vector3f / Vector3i are layout-compatible)
  void GatherValidMesh(std::vector<std::array<float, 3>>* vertex_buffer, std::vector<std::array<int32_t, 3>>* index_buffer,
                       std::vector<float>* vertex_prob_buffer) {
    int64_t nv = 0, nt = 0;
    check(rsf_gather_valid_mesh(g_, nullptr, 0, &nv, nullptr, 0, &nt, nullptr));
    vertex_buffer->resize((size_t)nv);
    vertex_prob_buffer->resize((size_t)nv);
    index_buffer->resize((size_t)nt);
    if (nv || nt)
      check(rsf_gather_valid_mesh(g_, (float*)vertex_buffer->data(), nv, &nv, (int32_t*)index_buffer->data(), nt, &nt,
                                  vertex_prob_buffer->data()));
    vertex_buffer->resize((size_t)nv);
    vertex_prob_buffer->resize((size_t)nv);
    index_buffer->resize((size_t)nt);
  }

#ifdef RSFUSION_WITH_EIGEN
  // The reference's signature (voxel_tsdf.cuh:104-106); Eigen::Vector3f / Vector3i are three packed 32-bit words.
  void GatherValidMesh(std::vector<Eigen::Vector3f>* vertex_buffer, std::vector<Eigen::Vector3i>* index_buffer,
                       std::vector<float>* vertex_prob_buffer) {
    static_assert(sizeof(Eigen::Vector3f) == 12 && sizeof(Eigen::Vector3i) == 12, "packed 3-vectors expected");
    int64_t nv = 0, nt = 0;
    check(rsf_gather_valid_mesh(g_, nullptr, 0, &nv, nullptr, 0, &nt, nullptr));
    vertex_buffer->resize((size_t)nv);
    vertex_prob_buffer->resize((size_t)nv);
    index_buffer->resize((size_t)nt);
    if (nv || nt)
      check(rsf_gather_valid_mesh(g_, reinterpret_cast<float*>(vertex_buffer->data()), nv, &nv,
                                  reinterpret_cast<int32_t*>(index_buffer->data()), nt, &nt, vertex_prob_buffer->data()));
    vertex_buffer->resize((size_t)nv);
    vertex_prob_buffer->resize((size_t)nv);
    index_buffer->resize((size_t)nt);
  }
#endif

  int NumActiveBlock() {   // hash_table_.NumActiveBlock(), voxel_hash.cu:225
    int64_t n = 0;
    check(rsf_num_active_blocks(g_, &n));
    return (int)n;
  }
  rsf_frame_stats Stats() {
    rsf_frame_stats s;
    check(rsf_get_stats(g_, &s));
    return s;
  }
  rsf_grid_t handle() const { return g_; }

 private:
  static void validate(const Image& rgb, const Image& depth, const Image& ht, const Image& lt) {
    if (rgb.type != kType8UC3 || depth.type != kType32FC1 || ht.type != kType32FC1 || lt.type != kType32FC1)
      throw Error(RSF_ERR_INVALID, "Integrate expects CV_8UC3 rgb and CV_32FC1 depth / ht / lt");
    if (rgb.cols != depth.cols || rgb.rows != depth.rows || depth.cols != ht.cols || depth.rows != ht.rows ||
        depth.cols != lt.cols || depth.rows != lt.rows)
      throw Error(RSF_ERR_INVALID, "Integrate: image sizes differ");
  }
  template <class Row>
  std::vector<Row> gather(int which, const float* bounds) {
    int64_t n = 0;
    auto call = [&](float* rows, int64_t cap) {
      if (which == 0) return rsf_gather_valid(g_, rows, cap, &n);
      if (which == 1) return rsf_gather_valid_semantic(g_, rows, cap, &n);
      return rsf_gather_voxels(g_, bounds, rows, cap, &n);
    };
    check(call(nullptr, 0));
    std::vector<Row> ret((size_t)n);
    if (n) check(call(reinterpret_cast<float*>(ret.data()), n));
    ret.resize((size_t)n);
    return ret;
  }
  rsf_grid_t g_ = nullptr;
};

}  // namespace rsfusion

#ifdef RSFUSION_GLOBAL_NAMES
using rsfusion::BoundingCube;
using rsfusion::CameraIntrinsics;
using rsfusion::CameraParams;
using rsfusion::SE3;
using rsfusion::TSDFGrid;
using rsfusion::VoxelSpatialTSDF;
using rsfusion::VoxelSpatialTSDFSEGM;
#endif

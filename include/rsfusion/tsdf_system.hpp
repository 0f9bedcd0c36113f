// rsfusion/tsdf_system.hpp -- re-host of the reference's TSDFSystem (modules/tsdf_module.{h,cc}):
// a worker thread that drains an input queue into a TSDFGrid, serialised against readers
// (Query / Render / DownloadAll) by a read mutex.  Same public surface:
//   TSDFSystem(voxel_size, truncation, max_depth, intrinsics, extrinsics = Identity)
//   Integrate / Query / Render x2 / DownloadAll / DownloadAllMesh / is_terminated / terminate / SetPause
//
// What changed underneath:
//   * Integrate copies the caller's images once, straight into a page-locked staging slot
//     (the reference clones four cv::Mat into pageable memory, modules/tsdf_module.cc:22-37, and
//     the grid then copies them again from pageable memory), so the H2D copy is a true async DMA.
//   * The worker uses the grid's pipelined entry point (rsf_integrate_async): the copy of frame k+1
//     overlaps the kernels of frame k.  It sleeps on a condition variable instead of spinning on an
//     empty queue (modules/tsdf_module.cc:101).
//   * Render writes RGBA images into host buffers instead of GLImage8UC4 textures.
#pragma once

#include <condition_variable>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include "tsdf_grid.hpp"

namespace rsfusion {

class TSDFSystem {
 public:
  TSDFSystem(float voxel_size, float truncation, float max_depth, const CameraIntrinsics<float>& intrinsics,
             const SE3<float>& extrinsics = SE3<float>::Identity(), const rsf_options* extra = nullptr, int queue_slots = 4)
      : tsdf_(voxel_size, truncation, extra), max_depth_(max_depth), intrinsics_(intrinsics), cam_T_posecam_(extrinsics) {
    slots_.resize(queue_slots < 2 ? 2 : queue_slots);
    for (size_t i = 0; i < slots_.size(); ++i) free_.push_back((int)i);
    t_ = std::thread(&TSDFSystem::Run, this);
  }
  ~TSDFSystem() {
    terminate();
    for (auto& s : slots_)
      if (s.base) rsf_host_free(s.base);
  }

  // modules/tsdf_module.cc:22-37.  img_ht / img_lt may be empty: all-ones maps are substituted
  // (ht == lt => the fused probability stays exactly 0.5).
  void Integrate(const SE3<float>& posecam_T_world, const Image& img_rgb, const Image& img_depth, const Image& img_ht = Image(),
                 const Image& img_lt = Image()) {
    {
      std::unique_lock<std::mutex> pause_lock(mtx_pause_);
      cv_pause_.wait(pause_lock, [&] { return !pause_; });
    }
    if (img_rgb.type != kType8UC3 || img_depth.type != kType32FC1 || img_rgb.rows != img_depth.rows || img_rgb.cols != img_depth.cols)
      throw Error(RSF_ERR_INVALID, "TSDFSystem::Integrate expects CV_8UC3 rgb and CV_32FC1 depth of equal size");
    int slot;
    {
      std::unique_lock<std::mutex> lock(mtx_queue_);
      cv_free_.wait(lock, [&] { return !free_.empty() || terminate_; });   // back-pressure instead of an unbounded queue
      if (terminate_) return;
      slot = free_.front();
      free_.pop_front();
    }
    Slot& s = slots_[slot];
    const size_t P = (size_t)img_depth.rows * img_depth.cols;
    ensure(s, P);
    s.rows = img_depth.rows;
    s.cols = img_depth.cols;
    s.cam_T_world = cam_T_posecam_ * posecam_T_world;   // modules/tsdf_module.cc:29
    std::memcpy(s.rgb, img_rgb.data, P * 3);
    std::memcpy(s.depth, img_depth.data, P * sizeof(float));
    if (img_ht.empty() || img_lt.empty()) {
      for (size_t i = 0; i < P; ++i) s.ht[i] = s.lt[i] = 1.f;
    } else {
      std::memcpy(s.ht, img_ht.data, P * sizeof(float));
      std::memcpy(s.lt, img_lt.data, P * sizeof(float));
    }
    {
      std::lock_guard<std::mutex> lock(mtx_queue_);
      inputs_.push_back(slot);
    }
    cv_work_.notify_one();
  }

  std::vector<VoxelSpatialTSDF> Query(const BoundingCube<float>& volumn) {   // :39-43
    std::lock_guard<std::mutex> lock(mtx_read_);
    return tsdf_.GatherVoxels(volumn);
  }
  // :45-55.  img_rgba / img_normal: img_h*img_w*4 bytes each (may be nullptr).
  void Render(const CameraParams& virtual_cam, const SE3<float> cam_T_world, uint8_t* img_rgba, uint8_t* img_normal) {
    Render(virtual_cam, cam_T_world, img_rgba, img_normal, max_depth_ * 2);
  }
  void Render(const CameraParams& virtual_cam, const SE3<float> cam_T_world, uint8_t* img_rgba, uint8_t* img_normal,
              float max_depth, float* depth = nullptr, float* prob = nullptr) {
    std::lock_guard<std::mutex> lock(mtx_read_);
    tsdf_.RayCast(max_depth, virtual_cam, cam_T_world, img_rgba, img_normal, depth, prob);
  }
  void DownloadAll(const std::string& file_path) {   // :57-64: raw N x 5 float32 rows
    std::lock_guard<std::mutex> lock(mtx_read_);
    const auto rows = tsdf_.GatherValidSemantic();
    std::ofstream fout(file_path, std::ios::out | std::ios::binary);
    fout.write((const char*)rows.data(), (std::streamsize)(rows.size() * sizeof(VoxelSpatialTSDFSEGM)));
  }
  void DownloadAllMesh(const std::string& vertices_path, const std::string& indices_path, const std::string& prob_path) {   // :66-86
    std::lock_guard<std::mutex> lock(mtx_read_);
    std::vector<std::array<float, 3>> v;
    std::vector<std::array<int32_t, 3>> idx;
    std::vector<float> p;
    tsdf_.GatherValidMesh(&v, &idx, &p);
    std::ofstream(vertices_path, std::ios::binary).write((const char*)v.data(), (std::streamsize)(v.size() * 12));
    std::ofstream(prob_path, std::ios::binary).write((const char*)p.data(), (std::streamsize)(p.size() * 4));
    std::ofstream(indices_path, std::ios::binary).write((const char*)idx.data(), (std::streamsize)(idx.size() * 12));
  }

  bool is_terminated() {
    std::lock_guard<std::mutex> lock(mtx_queue_);
    return terminate_;
  }
  // Unlike the reference (which drops whatever is still queued) pending frames are integrated first.
  void terminate() {
    {
      std::lock_guard<std::mutex> lock(mtx_queue_);
      terminate_ = true;
    }
    cv_work_.notify_all();
    cv_free_.notify_all();
    if (t_.joinable()) t_.join();
  }
  void SetPause(bool pause) {
    std::unique_lock<std::mutex> pause_lock(mtx_pause_);
    pause_ = pause;
    cv_pause_.notify_all();
  }
  // Block until every queued frame has been integrated (new; the reference has no such call).  Throws if a frame was
  // dropped because of an error since the last Flush (pool exhausted, CUDA error): the reference would have aborted.
  void Flush() {
    std::unique_lock<std::mutex> lock(mtx_queue_);
    cv_free_.wait(lock, [&] { return (inputs_.empty() && free_.size() >= slots_.size()) || worker_dead_; });
    if (dropped_frames_ > 0) {
      const std::string msg = last_error_;
      dropped_frames_ = 0;
      throw Error(RSF_ERR_CUDA, "TSDFSystem: frame(s) dropped: " + msg);
    }
  }
  // message of the most recent worker-thread error ("" if none) and the number of frames dropped since the last Flush
  std::string last_error() {
    std::lock_guard<std::mutex> lock(mtx_queue_);
    return last_error_;
  }
  int dropped_frames() {
    std::lock_guard<std::mutex> lock(mtx_queue_);
    return dropped_frames_;
  }
  TSDFGrid& grid() { return tsdf_; }

 private:
  struct Slot {
    void* base = nullptr;
    size_t cap = 0;
    uint8_t* rgb = nullptr;
    float *depth = nullptr, *ht = nullptr, *lt = nullptr;
    int rows = 0, cols = 0;
    SE3<float> cam_T_world;
  };
  static void ensure(Slot& s, size_t P) {
    if (s.cap >= P) return;
    if (s.base) rsf_host_free(s.base);
    const size_t off_d = (P * 3 + 255) & ~(size_t)255;
    check(rsf_host_alloc(&s.base, off_d + 3 * P * sizeof(float)));
    s.rgb = (uint8_t*)s.base;
    s.depth = (float*)((uint8_t*)s.base + off_d);
    s.ht = s.depth + P;
    s.lt = s.ht + P;
    s.cap = P;
  }
  // modules/tsdf_module.cc:88-115
  void Run() {
    std::deque<int> in_flight;   // slots whose H2D may still be running (at most 2: the grid double-buffers)
    for (;;) {
      int slot = -1;
      {
        std::unique_lock<std::mutex> lock(mtx_queue_);
        cv_work_.wait(lock, [&] { return !inputs_.empty() || terminate_; });
        if (inputs_.empty()) break;   // terminate requested and queue drained
        slot = inputs_.front();
        inputs_.pop_front();
      }
      bool queued = false;     // has `slot` joined in_flight?  (decides who recycles it on an error)
      try {
        std::lock_guard<std::mutex> lock(mtx_read_);
        Slot& s = slots_[slot];
        tsdf_.IntegrateAsync(Image(s.rgb, s.rows, s.cols, kType8UC3), Image(s.depth, s.rows, s.cols, kType32FC1),
                             Image(s.ht, s.rows, s.cols, kType32FC1), Image(s.lt, s.rows, s.cols, kType32FC1), max_depth_,
                             intrinsics_, s.cam_T_world);
        in_flight.push_back(slot);
        queued = true;
        bool idle;
        {
          std::lock_guard<std::mutex> ql(mtx_queue_);
          idle = inputs_.empty();
        }
        // keep one frame in flight while more work is queued; drain when idle so readers see it
        if (in_flight.size() >= 2 || idle) {
          tsdf_.Synchronize();
          std::lock_guard<std::mutex> ql(mtx_queue_);
          for (int f : in_flight) free_.push_back(f);
          in_flight.clear();
          cv_free_.notify_all();
        }
      } catch (const std::exception& e) {
        // A pool-full / CUDA error usually surfaces in Synchronize(), i.e. after `slot` has joined in_flight: recycle every
        // slot exactly once, and only after the copy engine can no longer be reading any of them (best effort: the
        // stream may itself be in error).  The reference aborts here (CUDA_SAFE_CALL); we keep running and keep the
        // message for last_error() / Flush().
        rsf_synchronize(tsdf_.handle());
        std::lock_guard<std::mutex> ql(mtx_queue_);
        last_error_ = e.what();
        ++dropped_frames_;
        if (!queued) free_.push_back(slot);
        for (int f : in_flight) free_.push_back(f);
        in_flight.clear();
        cv_free_.notify_all();
      }
    }
    std::lock_guard<std::mutex> ql(mtx_queue_);
    worker_dead_ = true;
    cv_free_.notify_all();
  }

  TSDFGrid tsdf_;
  float max_depth_;
  const CameraIntrinsics<float> intrinsics_;
  const SE3<float> cam_T_posecam_;
  std::vector<Slot> slots_;
  std::mutex mtx_queue_;
  std::deque<int> inputs_, free_;
  std::condition_variable cv_work_, cv_free_;
  std::mutex mtx_read_;
  bool terminate_ = false, worker_dead_ = false;
  std::mutex mtx_pause_;
  std::condition_variable cv_pause_;
  bool pause_ = false;
  std::string last_error_;   // guarded by mtx_queue_
  int dropped_frames_ = 0;
  std::thread t_;
};

}  // namespace rsfusion

#ifdef RSFUSION_GLOBAL_NAMES
using rsfusion::TSDFSystem;
#endif

// ref_harness.cu -- C-ABI harness around RA-SLAM's OWN TSDFGrid (test infrastructure).
//
// This file is NOT part of the product.  It is linked with the reference's unmodified
// utils/tsdf/{voxel_types,voxel_mem,voxel_hash,voxel_tsdf}.cu, compiled where they lie under
// /root/reference against the stand-in headers in oracle/ref_shim/ (Eigen 3.3.7 subset, cv::Mat,
// GLImage8UC4), into oracle/_ref/libref_tsdf.so (git-ignored; travels to the GPU box).  Only tests/,
// __graft_entry__.smoke() and bench.py's reference arm load it, as the checker / the baseline.
//
// A subclass reaches TSDFGrid's protected device buffers (utils/tsdf/voxel_tsdf.cuh:110-144) to read
// the raycast images and the voxel state back; no reference behaviour is altered.
#include <cstdint>
#include <cstring>
#include <vector>

#include "utils/tsdf/voxel_tsdf.cuh"

namespace {

// Collect every live hash entry (idx >= 0) and its 512 voxels.  Output order is table order with
// atomics, i.e. arbitrary: callers sort by coordinate.
__global__ void ref_export_kernel(const VoxelHashTable table, int16_t* coords, float* tsdf, uint8_t* rgbw,
                                  float* prob, int* count, int cap) {
  __shared__ int slot;
  const int entry = blockIdx.x;
  const VoxelBlock& block = table.GetBlock(entry);
  if (block.idx < 0) return;
  if (threadIdx.x == 0) slot = atomicAdd(count, 1);
  __syncthreads();
  if (slot >= cap) return;
  if (threadIdx.x < 3) coords[slot * 3 + threadIdx.x] = block.position[threadIdx.x];
  if (!tsdf) return;  // coordinates only
  for (int i = threadIdx.x; i < BLOCK_VOLUME; i += blockDim.x) {
    const VoxelRGBW& c = table.mem.GetVoxel<VoxelRGBW>(i, block);
    const size_t o = (size_t)slot * BLOCK_VOLUME + i;
    tsdf[o] = table.mem.GetVoxel<VoxelTSDF>(i, block).tsdf;
    prob[o] = table.mem.GetVoxel<VoxelSEGM>(i, block).probability;
    rgbw[o * 4 + 0] = c.rgb[0];
    rgbw[o * 4 + 1] = c.rgb[1];
    rgbw[o * 4 + 2] = c.rgb[2];
    rgbw[o * 4 + 3] = c.weight;
  }
}

__global__ void ref_count_kernel(const VoxelHashTable table, int* count) {
  const int entry = blockIdx.x * blockDim.x + threadIdx.x;
  if (entry < NUM_ENTRY && table.GetBlock(entry).idx >= 0) atomicAdd(count, 1);
}

class RefGrid : public TSDFGrid {
 public:
  RefGrid(float voxel_size, float truncation) : TSDFGrid(voxel_size, truncation) {}

  void ReadRaycast(int n, uint8_t* rgba, uint8_t* normal) {
    if (rgba) cudaMemcpy(rgba, img_tsdf_rgba_, sizeof(uchar4) * n, cudaMemcpyDeviceToHost);
    if (normal) cudaMemcpy(normal, img_tsdf_normal_, sizeof(uchar4) * n, cudaMemcpyDeviceToHost);
  }
  int64_t NumActive() const { return hash_table_.NumActiveBlock(); }
  int64_t NumLiveEntries() {
    int* d;
    cudaMalloc(&d, sizeof(int));
    cudaMemset(d, 0, sizeof(int));
    ref_count_kernel<<<NUM_ENTRY / 256, 256>>>(hash_table_, d);
    int h = 0;
    cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return h;
  }
  int64_t Export(int16_t* coords, float* tsdf, uint8_t* rgbw, float* prob, int64_t cap) {
    int16_t* dc;
    float *dt = nullptr, *dp = nullptr;
    uint8_t* dr = nullptr;
    int* dn;
    cudaMalloc(&dc, cap * 3 * sizeof(int16_t));
    if (tsdf) {
      cudaMalloc(&dt, cap * BLOCK_VOLUME * sizeof(float));
      cudaMalloc(&dp, cap * BLOCK_VOLUME * sizeof(float));
      cudaMalloc(&dr, cap * BLOCK_VOLUME * 4);
    }
    cudaMalloc(&dn, sizeof(int));
    cudaMemset(dn, 0, sizeof(int));
    ref_export_kernel<<<NUM_ENTRY, 128>>>(hash_table_, dc, dt, dr, dp, dn, (int)cap);
    int n = 0;
    cudaMemcpy(&n, dn, sizeof(int), cudaMemcpyDeviceToHost);
    const int64_t m = n < cap ? n : cap;
    cudaMemcpy(coords, dc, m * 3 * sizeof(int16_t), cudaMemcpyDeviceToHost);
    if (tsdf) {
      cudaMemcpy(tsdf, dt, m * BLOCK_VOLUME * sizeof(float), cudaMemcpyDeviceToHost);
      cudaMemcpy(prob, dp, m * BLOCK_VOLUME * sizeof(float), cudaMemcpyDeviceToHost);
      cudaMemcpy(rgbw, dr, m * BLOCK_VOLUME * 4, cudaMemcpyDeviceToHost);
      cudaFree(dt);
      cudaFree(dp);
      cudaFree(dr);
    }
    cudaFree(dc);
    cudaFree(dn);
    return n;
  }
};

inline SE3<float> pose_of(const float* p) {  // {qw,qx,qy,qz,tx,ty,tz} of cam_T_world
  return SE3<float>(Eigen::Quaternionf(p[0], p[1], p[2], p[3]), Eigen::Vector3f(p[4], p[5], p[6]));
}

}  // namespace

extern "C" {

void* ref_create(float voxel_size, float truncation) { return new RefGrid(voxel_size, truncation); }
void ref_destroy(void* g) { delete static_cast<RefGrid*>(g); }

// TSDFGrid::Integrate (utils/tsdf/voxel_tsdf.cu:416-452) on caller-owned host images.
void ref_integrate(void* g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H, int W,
                   float max_depth, const float* K, const float* pose) {
  const cv::Mat m_rgb(H, W, CV_8UC3, const_cast<uint8_t*>(rgb));
  const cv::Mat m_depth(H, W, CV_32FC1, const_cast<float*>(depth));
  const cv::Mat m_ht(H, W, CV_32FC1, const_cast<float*>(ht));
  const cv::Mat m_lt(H, W, CV_32FC1, const_cast<float*>(lt));
  static_cast<RefGrid*>(g)->Integrate(m_rgb, m_depth, m_ht, m_lt, max_depth,
                                      CameraIntrinsics<float>(K[0], K[1], K[2], K[3]), pose_of(pose));
}

// TSDFGrid::RayCast (utils/tsdf/voxel_tsdf.cu:885-902); rgba / normal: optional host outputs (H*W*4 u8).
void ref_raycast(void* g, float max_depth, int H, int W, const float* K, const float* pose, uint8_t* rgba,
                 uint8_t* normal) {
  RefGrid* r = static_cast<RefGrid*>(g);
  r->RayCast(max_depth, CameraParams(CameraIntrinsics<float>(K[0], K[1], K[2], K[3]), H, W), pose_of(pose));
  r->ReadRaycast(H * W, rgba, normal);
}

int64_t ref_num_active_blocks(void* g) { return static_cast<RefGrid*>(g)->NumActive(); }
int64_t ref_num_live_entries(void* g) { return static_cast<RefGrid*>(g)->NumLiveEntries(); }

int64_t ref_export_blocks(void* g, int16_t* coords, float* tsdf, uint8_t* rgbw, float* prob, int64_t cap) {
  return static_cast<RefGrid*>(g)->Export(coords, tsdf, rgbw, prob, cap);
}

// GatherValid / GatherValidSemantic / GatherVoxels (utils/tsdf/voxel_tsdf.cu:476-559): rows of 4 / 5 / 4 f32.
int64_t ref_gather_valid(void* g, float* rows, int64_t cap_rows) {
  const std::vector<VoxelSpatialTSDF> v = static_cast<RefGrid*>(g)->GatherValid();
  cudaDeviceSynchronize();
  const int64_t n = (int64_t)v.size();
  if (rows) std::memcpy(rows, v.data(), sizeof(VoxelSpatialTSDF) * (n < cap_rows ? n : cap_rows));
  return n;
}
int64_t ref_gather_valid_semantic(void* g, float* rows, int64_t cap_rows) {
  const std::vector<VoxelSpatialTSDFSEGM> v = static_cast<RefGrid*>(g)->GatherValidSemantic();
  cudaDeviceSynchronize();
  const int64_t n = (int64_t)v.size();
  if (rows) std::memcpy(rows, v.data(), sizeof(VoxelSpatialTSDFSEGM) * (n < cap_rows ? n : cap_rows));
  return n;
}
int64_t ref_gather_voxels(void* g, const float* bounds, float* rows, int64_t cap_rows) {
  const BoundingCube<float> cube{bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5]};
  const std::vector<VoxelSpatialTSDF> v = static_cast<RefGrid*>(g)->GatherVoxels(cube);
  cudaDeviceSynchronize();
  const int64_t n = (int64_t)v.size();
  if (rows) std::memcpy(rows, v.data(), sizeof(VoxelSpatialTSDF) * (n < cap_rows ? n : cap_rows));
  return n;
}

// GatherValidMesh (utils/tsdf/voxel_tsdf.cu:736-845).  Call with NULL outputs to size, then again to fill
// (the reference recomputes; the mesh of a static grid is deterministic up to vertex order).
void ref_gather_valid_mesh(void* g, float* vertices, int64_t cap_v, int64_t* n_v, int32_t* triangles, int64_t cap_t,
                           int64_t* n_t, float* vertex_prob) {
  std::vector<Eigen::Vector3f> vb;
  std::vector<Eigen::Vector3i> ib;
  std::vector<float> pb;
  static_cast<RefGrid*>(g)->GatherValidMesh(&vb, &ib, &pb);
  *n_v = (int64_t)vb.size();
  *n_t = (int64_t)ib.size();
  const int64_t mv = *n_v < cap_v ? *n_v : cap_v, mt = *n_t < cap_t ? *n_t : cap_t;
  if (vertices) std::memcpy(vertices, vb.data(), sizeof(float) * 3 * mv);
  if (vertex_prob) std::memcpy(vertex_prob, pb.data(), sizeof(float) * mv);
  if (triangles) std::memcpy(triangles, ib.data(), sizeof(int32_t) * 3 * mt);
}

int ref_last_cuda_error() { return (int)cudaGetLastError(); }

}  // extern "C"

// rsf_kernels.h -- host-callable launch wrappers of the kernels in rsf_kernels.cu.
#pragma once
#include <algorithm>

#include "rsf_device.cuh"

namespace rsf {

void launch_init_pool(const GridView& g, cudaStream_t s);
void launch_begin_frame(const GridView& g, cudaStream_t s);
// empty the table, then re-insert every alive block of the dense block list (drops ghost keys); between frames only
void launch_rebuild_table(const GridView& g, size_t table_slots, int sm_count, cudaStream_t s);
void launch_allocate(const GridView& g, const Frame& f, const uint8_t* rgb, const float* depth, const float* ht,
                     const float* lt, FramePixel* frame, cudaStream_t s);
void launch_allocate_u16(const GridView& g, const Frame& f, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht,
                         const uint16_t* lt, float depth_alpha, float prob_alpha, FramePixel* frame, cudaStream_t s);
// per-tile maximum of the usable depth (16x16-pixel tiles) for the occlusion culling in select_visible
void launch_depth_tiles(const void* depth, bool u16, float depth_alpha, int W, int H, float max_depth, float* tiles, cudaStream_t s);
size_t depth_tiles_count(int W, int H);
// tiles == NULL: no culling (every alive block with a corner in view is selected)
void launch_select_visible(const GridView& g, const Frame& f, const float* tiles, int sm_count, cudaStream_t s);
int integrate_ctas_per_sm();   // resident CTAs per SM of integrate_kernel on the current device
void launch_integrate(const GridView& g, const Frame& f, const FramePixel* frame, int sm_count, int ctas_per_sm, cudaStream_t s);
// K-class semantic extension: per-class log-probability planes, updated with the weights from before integrate
void launch_class_update(const GridView& g, const Frame& f, const FramePixel* frame, const float* class_prob, int K,
                         float* class_pool, int sm_count, cudaStream_t s);
// rows of gather_list[blk0 .. blk0 + nblk) written to rows[0 ..)
void launch_download_class_rows(const GridView& g, float vs, int blk0, int nblk, int K, const float* class_pool, float* rows,
                                int sm_count, cudaStream_t s);
size_t raycast_ranges_bytes(int W, int H);
// merge != NULL: hits go to the merge planes (depth bits << 32 | payload, system-scope atomicMin) instead of the images
void launch_raycast(const GridView& g, const Frame& f, float step_size, unsigned* ranges, uchar4* rgba, uchar4* normal,
                    float* depth, float* prob, int sm_count, cudaStream_t s, unsigned long long* merge = nullptr,
                    int merge_planes = 0);
// the same in two steps: the per-tile pre-pass once, then one launch per band of 4-pixel patch rows (band = 0 / 1 picks
// the work queue); the host path renders two bands so the read-back of the first overlaps the march of the second
void launch_raycast_prepass(const GridView& g, const Frame& f, float step_size, unsigned* ranges, int sm_count, cudaStream_t s);
void launch_raycast_band(const GridView& g, const Frame& f, float step_size, const unsigned* ranges, uchar4* rgba, uchar4* normal,
                         float* depth, float* prob, int sm_count, cudaStream_t s, int patch_row0, int patch_rows, int band,
                         unsigned long long* merge = nullptr, int merge_planes = 0);
void launch_merge_finalize(unsigned long long* merge, int merge_planes, int P, uchar4* rgba, uchar4* normal, float* depth,
                           float* prob, int sm_count, cudaStream_t s);
// sharded frame: arrival flags in the presenting rank's memory (system-scope release / acquire)
void launch_shard_signal(unsigned* flag, unsigned seq, cudaStream_t s);
void launch_shard_wait(const unsigned* flags, int world, unsigned seq, cudaStream_t s);
void launch_gather_select(const GridView& g, int mode, int owned_only, S3 lo, S3 hi, int sm_count, cudaStream_t s);
void launch_download_rows(const GridView& g, float vs, int blk0, int nblk, int ncol, float* rows, int sm_count, cudaStream_t s);
void launch_export_blocks(const GridView& g, int nblk, short* coords, float* tsdf, uint32_t* rgbw, float* prob,
                          int sm_count, cudaStream_t s);
void launch_allocate_coords(const GridView& g, const short* coords, int n, int sm_count, cudaStream_t s);
void launch_delete_coords(const GridView& g, const short* coords, int n, cudaStream_t s);
void launch_retrieve_voxels(const GridView& g, const short* pts, int n, float* tsdf, uint32_t* rgbw, float* prob,
                            uint8_t* found, cudaStream_t s);
void launch_assign_voxels(const GridView& g, const short* pts, int n, const float* tsdf, const uint32_t* rgbw,
                          const float* prob, uint8_t* written, cudaStream_t s);
// rsf_mesh.cu: marching cubes over gather_list[0 .. nblk): count + scan, then emit into exact-size arrays
void launch_mesh_count(const GridView& g, int nblk, int2* counts, int2* offsets, long long* totals, cudaStream_t s);
void launch_mesh_emit(const GridView& g, float vs, int nblk, const int2* offsets, float* vertices, float* vprob,
                      int* triangles, cudaStream_t s);
void launch_trilinear(const GridView& g, const float* pts, int n, float* out, cudaStream_t s);

// cv::resize(src, dst, Size(dW, dH)) (INTER_LINEAR default) for 8UC3 / 8UC1 / 16UC1 / 32FC1 images; -1 = unsupported format
int launch_resize(const void* src, int sH, int sW, void* dst, int dH, int dW, int channels, int elem_bytes, cudaStream_t s);

}  // namespace rsf

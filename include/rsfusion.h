/* rsfusion.h -- C ABI of the B200-native voxel-hashing TSDF + semantic fusion engine.
 *
 * This is the drop-in boundary for RA-SLAM's `TSDFGrid` (reference utils/tsdf/voxel_tsdf.cuh:39-145)
 * and the table/pool below it (`VoxelHashTable` utils/tsdf/voxel_hash.cuh:47-166, `VoxelMemPool`
 * utils/tsdf/voxel_mem.cuh:97-176).  Plain pointers and sizes only: no OpenCV / Eigen / GL / torch
 * types cross this line.  The header-only C++ adapter include/rsfusion/tsdf_grid.hpp puts the
 * reference's class and method names back on top of it.
 *
 * Conventions
 *   - every entry point returns an `int` status (RSF_OK == 0); no exceptions cross the ABI;
 *     rsf_last_error() gives a human-readable message for the calling thread's last failure.
 *   - images are row-major, tightly packed (`cv::Mat::isContinuous()`), H rows x W columns.
 *   - K  = {fx, fy, cx, cy}                       (CameraIntrinsics<float>, utils/cuda/camera.cuh:13-31)
 *   - pose = {qw, qx, qy, qz, tx, ty, tz} of cam_T_world (SE3<float>: Eigen quaternion + translation,
 *     utils/cuda/lie_group.cuh:8-45).  The quaternion is used as given (not re-normalised).
 *   - `*_device` variants take device pointers valid on the grid's device and only enqueue work on
 *     the grid's stream; the host variants are blocking, like the reference's methods
 *     (voxel_tsdf.cu:450, :901).
 *   - a grid is not re-entrant (the reference serialises calls with TSDFSystem::mtx_read_,
 *     modules/tsdf_module.cc:40-107); different grids may be used from different threads.
 *   - there is no CPU fallback: every call fails with RSF_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RSFUSION_H_
#define RSFUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSF_ABI_VERSION 2

#if defined(__GNUC__)
#define RSF_API __attribute__((visibility("default")))
#else
#define RSF_API
#endif

enum rsf_status {
  RSF_OK = 0,
  RSF_ERR_INVALID = 1,      /* bad argument (NULL pointer, size, dtype contract)            */
  RSF_ERR_CUDA = 2,         /* CUDA runtime error, message in rsf_last_error()             */
  RSF_ERR_POOL_FULL = 3,    /* voxel-block pool exhausted (reference: device assert,       */
                            /* utils/tsdf/voxel_mem.cu:38-39)                              */
  RSF_ERR_TABLE_FULL = 4,   /* hash table probe limit reached                              */
  RSF_ERR_CAPACITY = 5      /* caller-provided output buffer too small (size is returned)  */
};

typedef struct rsf_grid_s* rsf_grid_t;

/* rsf_options.flags: use rsf_options.stream as given even when it is NULL (the legacy default stream). */
#define RSF_FLAG_EXTERNAL_STREAM 1u

/* Construction parameters.  voxel_size/truncation are the two arguments of
 * TSDFGrid::TSDFGrid (voxel_tsdf.cu:376-377); the rest replace compile-time #defines of the
 * reference (NUM_BLOCK voxel_mem.cuh:11-13, NUM_ENTRY voxel_hash.cuh:13-25, MAX_IMG_SIZE
 * voxel_tsdf.cu:11-13) and add the spatial-shard description (new).  Zero selects the default. */
typedef struct rsf_options {
  float voxel_size;        /* [m]                                                           */
  float truncation;        /* [m]                                                           */
  int device;              /* CUDA ordinal; -1 = current device                             */
  int pool_blocks;         /* capacity in 8^3-voxel blocks; default 1<<18 (reference)       */
  int table_slots_log2;    /* open-addressed slots; default 22 (reference NUM_ENTRY)        */
  int max_pixels;          /* largest H*W accepted; default 1920*1080 (reference)           */
  int shard_rank;          /* spatial sharding: this grid integrates only blocks it owns    */
  int shard_world;         /*   (+halo); world <= 1 disables sharding                       */
  int shard_shift;         /*   owner = mix(block_pos >> shift) % world; default 4 (16^3-   */
                           /*   block super-bricks: ~1.4x replication with a 1-block halo)  */
  int shard_halo;          /*   redundant halo in blocks; default 1                         */
  void* stream;            /* optional external cudaStream_t; NULL = library-owned stream   */
  unsigned flags;          /* RSF_FLAG_* bits                                               */
  int num_classes;         /* K-class semantic extension: 0 / 2 = the reference's ht / lt only;   */
                           /*   3..64 adds a pool of K log-probability planes per block            */
                           /*   (pool_blocks * K * 2 KiB) and enables rsf_integrate_classes*       */
} rsf_options;

/* Per-frame counters of the last rsf_integrate* call (reference logs only block counts,
 * voxel_tsdf.cu:442,451,865). */
typedef struct rsf_frame_stats {
  int64_t active_blocks;   /* == TSDFGrid's hash_table_.NumActiveBlock() after the call     */
  int64_t new_blocks;      /* blocks allocated by this frame                                */
  int64_t visible_blocks;  /* blocks selected for integration                               */
  int64_t updated_voxels;  /* voxels that passed sdf > -truncation                          */
  int64_t carved_blocks;   /* blocks removed by space carving                               */
  int64_t high_water;      /* pool slots ever used                                          */
  int64_t culled_blocks;   /* of visible_blocks: proven un-updatable (occluded) and skipped */
  int64_t table_keys;      /* claimed hash-table slots = alive blocks + ghosts of carved    */
                           /* ones; the table is rebuilt when this passes half its size     */
  int64_t table_rebuilds;  /* rebuilds since creation                                       */
} rsf_frame_stats;

RSF_API void rsf_default_options(rsf_options* opts);
RSF_API int rsf_abi_version(void);
RSF_API const char* rsf_last_error(void);

/* TSDFGrid::TSDFGrid / ~TSDFGrid (voxel_tsdf.cu:376-414). */
RSF_API int rsf_create(const rsf_options* opts, rsf_grid_t* out);
RSF_API int rsf_destroy(rsf_grid_t g);

/* TSDFGrid::Integrate (voxel_tsdf.cu:416-452): allocate -> select visible -> integrate -> carve.
 * rgb: H*W*3 u8 (RGB order, CV_8UC3); depth: H*W f32 metres, z-depth, 0 = invalid (CV_32FC1);
 * ht/lt: H*W f32 high-/low-touch probabilities (CV_32FC1).  Host buffers; blocking. */
RSF_API int rsf_integrate(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt,
                  int H, int W, float max_depth, const float K[4], const float pose[7]);
/* Pipelined form: returns as soon as the copies and kernels are enqueued (H2D on a copy stream,
 * four staging slots, overlapping the kernels of the previous frames).  The host buffers must stay
 * valid and unchanged until a later rsf_synchronize / rsf_wait on this grid returns; pass
 * page-locked memory (rsf_host_alloc) for true asynchronous DMA.  Device-side errors (pool or table
 * exhaustion) surface at the next rsf_synchronize / blocking call.  This is the path
 * TSDFSystem's worker thread (modules/tsdf_module.cc:88-115) maps onto. */
RSF_API int rsf_integrate_async(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht,
                                const float* lt, int H, int W, float max_depth, const float K[4],
                                const float pose[7]);
/* Same with the frame already resident in device memory; asynchronous on the grid's stream. */
RSF_API int rsf_integrate_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht,
                         const float* lt, int H, int W, float max_depth, const float K[4],
                         const float pose[7]);

/* Integrate from the raw 16-bit images (new; GPU-side input preparation).  The reference's callers decode / receive
 * uint16 depth and uint16 ht / lt images and convert them on the CPU every frame before TSDFGrid::Integrate
 * (examples/tsdf/offline.cc:62-76: depth.convertTo(CV_32FC1, 1. / depth_scale), ht / lt .convertTo(CV_32FC1, 1. / 65535);
 * main/offline_eval.cc:74; modules/disinfect_slam.cc:37-40).  Here the conversion happens in the staging kernel with
 * convertTo's arithmetic, float(raw) * float(1. / factor), so results are identical to converting first and calling
 * rsf_integrate, 9 instead of 15 bytes per pixel cross PCIe and the host does no per-pixel work.  depth_factor is the
 * dataset's `depthmap_factor` (raw units per metre).  ht == lt == NULL means "no segmentation": ht = 0, lt = 1
 * (offline.cc:72-75).  Blocking / pipelined / device-resident forms as for rsf_integrate. */
RSF_API int rsf_integrate_u16(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht,
                              const uint16_t* lt, int H, int W, float depth_factor, float max_depth, const float K[4],
                              const float pose[7]);
RSF_API int rsf_integrate_u16_async(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht,
                                    const uint16_t* lt, int H, int W, float depth_factor, float max_depth,
                                    const float K[4], const float pose[7]);
RSF_API int rsf_integrate_u16_device(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht,
                                     const uint16_t* lt, int H, int W, float depth_factor, float max_depth,
                                     const float K[4], const float pose[7]);

/* cv::resize(src, dst, cv::Size(dst_W, dst_H)) with its default INTER_LINEAR, on the grid's device and stream (new;
 * GPU-side input preparation).  The reference's callers resize every frame on the CPU before TSDFGrid::Integrate:
 * disinfect_slam/disinfect_slam.cc:37-38 (rgb 8UC3 and the raw 16UC1 depth down to the TSDF size),
 * utils/offline_data_provider/scannet_sens_reader.cc:78 (1296x968 colour -> 640x480), segmentation/inference.cc:28 (32FC1
 * probability maps).  Formats: channels = 3 with elem_bytes = 1 (8UC3); channels = 1 with elem_bytes = 1, 2 or 4 (8UC1,
 * 16UC1, 32FC1).  Results equal OpenCV's portable code path bit for bit for the 8- and 16-bit types (fixed-point /
 * float weights as in imgproc/resize.cpp; exact 2x decimation is the 2x2 mean cv::resize switches to).  src and dst are
 * device pointers to continuous images; asynchronous on the grid's stream. */
RSF_API int rsf_resize_device(rsf_grid_t g, const void* src, int src_H, int src_W, void* dst, int dst_H, int dst_W,
                              int channels, int elem_bytes);

/* K-class semantic fusion (extension; rsf_options.num_classes = K > 2).  The reference fuses two planes, ht and lt
 * (voxel_tsdf.cu:241-248); its K-class form is
 *   p_k <- exp((w_old ln p_k + w_new ln q_k) / w) / sum_j exp((w_old ln p_j + w_new ln q_j) / w)
 * with the same w_old / w_new and the same update decision as the TSDF update.  class_prob: K planes of H*W f32
 * probabilities (q_k > 0).  Everything else (allocation, TSDF, weight, colour, carving) is exactly rsf_integrate; the binary
 * ht / lt channel is fed ht = lt and therefore keeps its prior.  _device: device pointers, asynchronous on the grid's
 * stream; the other form takes host pointers and blocks. */
RSF_API int rsf_integrate_classes_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* class_prob, int H,
                                         int W, float max_depth, const float K[4], const float pose[7]);
RSF_API int rsf_integrate_classes(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* class_prob, int H, int W,
                                  float max_depth, const float K[4], const float pose[7]);
/* K-class counterpart of GatherValidSemantic: rows of 4 + K f32 (x, y, z, tsdf, p_0 .. p_{K-1}), probabilities normalised
 * (uniform 1/K for never-updated voxels).  Size query with rows == NULL. */
RSF_API int rsf_gather_valid_classes(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows);

/* TSDFGrid::RayCast (voxel_tsdf.cu:885-902).  rgba / normal are the reference's two RGBA8 images
 * (img_tsdf_rgba_, img_tsdf_normal_); depth (z of the refined hit in the camera frame, 0 = no hit)
 * and prob (fused probability at the hit voxel) are new outputs.  Any output may be NULL.
 * Host buffers; blocking. */
RSF_API int rsf_raycast(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                uint8_t* rgba, uint8_t* normal, float* depth, float* prob);
/* Pipelined form: kernels on the grid's stream, read-back on a separate D2H stream, double-buffered.
 * *ticket identifies the read-back; the host outputs are valid after rsf_wait(g, ticket) (or
 * rsf_synchronize).  At most four read-backs are in flight: a fifth call reuses the first ticket's
 * staging and implicitly orders after it (so wait for ticket k before issuing call k + 4, and keep the host
 * output buffers of a call untouched until its ticket has been waited for). */
RSF_API int rsf_raycast_async(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                              uint8_t* rgba, uint8_t* normal, float* depth, float* prob, int* ticket);
RSF_API int rsf_wait(rsf_grid_t g, int ticket);
/* Device outputs; asynchronous on the grid's stream. */
RSF_API int rsf_raycast_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                       uint8_t* rgba, uint8_t* normal, float* depth, float* prob);

/* TSDFGrid::RayCast as the reference's callers use it (GLImage8UC4::LoadCuda takes a CUDA pointer, voxel_tsdf.cu:895-900):
 * render into the grid's OWN device images (the reference's img_tsdf_rgba_ / img_tsdf_normal_) and return their device
 * addresses and the cudaStream_t (as void*) the kernel was enqueued on.  The images are valid until the next raycast on
 * this grid; the call does not wait for the kernel. */
RSF_API int rsf_raycast_internal_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                                        const uint8_t** rgba_dev, const uint8_t** normal_dev, void** stream);

/* Spatially sharded grids (rsf_options.shard_world > 1): raycast with the per-pixel min-depth merge fused into the
 * kernel (new; the reference is single-GPU).  merge_planes points to n_planes * H * W 64-bit words preset to all-ones --
 * normally the PRESENTING rank's buffer, mapped into this process as NVLink peer memory (torch symmetric memory,
 * cudaIpc, ...).  Every hit this rank owns is one system-scope atomicMin per plane of (depth bits << 32 | payload):
 * plane 0 payload = RGBA8, plane 1 = the normal image, plane 2 = probability bits.  A hit whose refined position has
 * no positive finite depth (NaN from t_prev == t_curr in the interpolation, voxel_tsdf.cu:327) cannot be ordered and is
 * dropped.  Asynchronous on the grid's stream; no image is written locally. */
RSF_API int rsf_raycast_merge_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                                     uint64_t* merge_planes, int n_planes);
/* On the presenting rank, after all ranks' rsf_raycast_merge_device kernels have completed (cross-rank barrier):
 * unpack the planes into the reference's images (+ depth / prob, any may be NULL) and reset the planes to all-ones. */
RSF_API int rsf_merge_finalize_device(rsf_grid_t g, int H, int W, uint64_t* merge_planes, int n_planes, uint8_t* rgba,
                                      uint8_t* normal, float* depth, float* prob);

/* One frame of a spatially sharded scene in ONE call per rank (new): integrate the frame into this rank's shard, raycast
 * the virtual view against it with the min-depth merge fused in, raise this rank's arrival flag in the presenting rank's
 * memory and -- on the presenting rank -- wait for all flags, unpack the merged planes into the images and re-arm them.
 * Everything is enqueued on the grid's stream; nothing blocks the host and no host-side barrier is involved.
 *
 * rsf_shard_link describes the presenting rank's buffers as seen from THIS process (its own device memory on the
 * presenting rank, NVLink peer mappings elsewhere -- torch symmetric memory, cudaIpc, ...):
 *   planes[2]  two sets (frame parity) of n_planes * H * W 64-bit words preset to all-ones; a set is re-armed by the
 *              presenting rank when it unpacks it, and written again two frames later -- the caller must order frame
 *              k + 2's input on every rank after frame k's call on the presenting rank (a broadcast of the frame from
 *              the presenting rank on the same stream does that);
 *   flags      shard_world 32-bit words preset to 0; rank r stores the frame sequence number into flags[r] (system-scope
 *              release), the presenting rank acquires them.  All ranks must issue the same sequence of calls.
 * With shard_world == 1 no link is needed and the call is rsf_integrate_device + rsf_raycast_device.
 * Inputs are device pointers (the frame as rsf_integrate_device takes it); the outputs are device pointers used on the
 * presenting rank only (any may be NULL). */
typedef struct rsf_shard_link {
  uint64_t* planes[2];
  uint32_t* flags;
  int n_planes;            /* 1..3 planes per set: rgba | + normal | + prob                   */
  unsigned skip_mask;      /* bit k set: nobody wants plane k (1 = normal, 2 = prob); it is   */
                           /* neither written nor unpacked.  Must be the same on all ranks.  */
  int root;                /* presenting rank                                                 */
  int H, W;                /* image size the planes hold                                      */
  /* ---- optional frame transport (rsf_shard_feed_device / rsf_shard_step_device); all zero = the caller moves frames ---- */
  int frame_H, frame_W;    /* input frame size; a packed frame is [rgb u8 x3 | depth f32 | ht f32 | lt f32] = 15 H W bytes  */
  uint8_t* rx[2];          /* this rank's receive buffers (frame parity), device memory its tree parent can write    */
  uint32_t* frame_flag;    /* 2 words next to them, preset 0: frame_flag[j] = sequence number of the frame in rx[j]  */
  int n_kids;              /* children of this rank in the push tree rooted at `root` (0..2)                         */
  uint8_t* kid_rx[2][2];   /* [kid][parity]: the children's receive buffers, peer-mapped into this process           */
  uint32_t* kid_flag[2];   /* [kid]: the children's frame_flag words, peer-mapped                                    */
} rsf_shard_link;
RSF_API int rsf_shard_set_link(rsf_grid_t g, const rsf_shard_link* link);
RSF_API int rsf_shard_frame_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt,
                                   int H, int W, float max_depth, const float K[4], const float pose[7],
                                   float raycast_max_depth, int rc_H, int rc_W, const float rc_K[4],
                                   const float rc_pose[7], uint8_t* rgba, uint8_t* normal, float* depth_out,
                                   float* prob_out);
/* The same frame with the transport inside the library: two calls per frame per rank, nothing else on the host.
 * rsf_shard_feed_device starts moving the NEXT frame down a binary tree of copy-engine pushes rooted at the presenting /
 * ingest rank (packed_frame: device pointer on that rank, ignored elsewhere): each rank's copy engines write the frame
 * into its children's receive buffers over NVLink and then the frame's sequence number into their frame_flag word
 * (stream memory operation); a rank forwards as soon as its own copy has landed (stream wait on its flag).  No SM and no
 * collective kernel is involved, so the transfer of frame k + 1 overlaps the kernels of frame k.  The ingest rank gates
 * the push of frame k + 2 on the completion of its own step k, which includes the arrival flags of every rank for frame k,
 * so a receive buffer is never overwritten before its frame has been consumed.  The packed frame must stay valid on the
 * ingest rank until the step that consumes it has completed.
 * rsf_shard_step_device consumes the oldest fed frame: wait (on the device) for its arrival, then exactly
 * rsf_shard_frame_device on the buffer.  Call order on every rank: feed(0); then per frame k: step(k); feed(k + 1). */
RSF_API int rsf_shard_feed_device(rsf_grid_t g, const uint8_t* packed_frame);
RSF_API int rsf_shard_step_device(rsf_grid_t g, float max_depth, const float K[4], const float pose[7], float raycast_max_depth,
                                  const float rc_K[4], const float rc_pose[7], uint8_t* rgba, uint8_t* normal,
                                  float* depth_out, float* prob_out);

/* Make the grid's stream wait until the 32-bit word at `flag` (device memory of this GPU, e.g. written by a peer's copy
 * engine over NVLink) satisfies (int32)(*flag - value) >= 0.  Uses the driver's stream memory operation (no SM involved);
 * falls back to a one-thread polling kernel.  Used by the sharded scene to order a frame's kernels after the frame's
 * arrival without a host round trip. */
RSF_API int rsf_stream_wait_flag_device(rsf_grid_t g, const uint32_t* flag, uint32_t value);
/* Same on an arbitrary stream of device `device` (cudaStream_t passed as void*): the forwarding streams of the frame tree. */
RSF_API int rsf_stream_wait_flag_on(int device, void* stream, const uint32_t* flag, uint32_t value);

/* Block until everything enqueued on the grid's stream is done; surfaces device-side errors
 * (pool/table exhaustion) of the asynchronous variants. */
RSF_API int rsf_synchronize(rsf_grid_t g);

/* VoxelHashTable::NumActiveBlock (voxel_hash.cu:225). */
RSF_API int rsf_num_active_blocks(rsf_grid_t g, int64_t* n);
RSF_API int rsf_get_stats(rsf_grid_t g, rsf_frame_stats* stats);

/* TSDFGrid::GatherValid (voxel_tsdf.cu:476-502): rows of VoxelSpatialTSDF {x,y,z,tsdf} (16 B);
 * TSDFGrid::GatherValidSemantic (:504-530): rows of VoxelSpatialTSDFSEGM {x,y,z,tsdf,prob} (20 B);
 * TSDFGrid::GatherVoxels (:532-559): GatherValid restricted to blocks fully inside
 * bounds = {xmin,xmax,ymin,ymax,zmin,zmax} [m] (BoundingCube<float>, voxel_tsdf.cuh:20-34).
 * Size-query-then-fill: call with rows == NULL to get *n_rows; with rows != NULL, cap_rows must be
 * >= the number of rows or RSF_ERR_CAPACITY is returned (and *n_rows still set).  With sharding,
 * only rows of blocks this rank owns are produced (halo copies are dropped). */
RSF_API int rsf_gather_valid(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows);
RSF_API int rsf_gather_valid_semantic(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows);
RSF_API int rsf_gather_voxels(rsf_grid_t g, const float bounds[6], float* rows, int64_t cap_rows, int64_t* n_rows);

/* TSDFGrid::GatherValidMesh (voxel_tsdf.cu:736-845): marching cubes over all active blocks.
 * vertices: n_vertices x 3 f32 [m]; triangles: n_triangles x 3 i32 indices into vertices;
 * vertex_prob: n_vertices f32.  Same size-query-then-fill convention (pass NULL buffers first). */
RSF_API int rsf_gather_valid_mesh(rsf_grid_t g, float* vertices, int64_t cap_vertices, int64_t* n_vertices,
                          int32_t* triangles, int64_t cap_triangles, int64_t* n_triangles,
                          float* vertex_prob);

/* Raw block state for parity tests: coords n x 3 i16 (block coordinates), tsdf n x 512 f32,
 * rgbw n x 512 x 4 u8 (r,g,b,weight), prob n x 512 f32; voxel order x + 8y + 64z
 * (OffsetToIndex, voxel_mem.cuh:67-70).  Block order is unspecified.  Outputs may be NULL. */
RSF_API int rsf_export_blocks(rsf_grid_t g, int owned_only, int16_t* coords, float* tsdf, uint8_t* rgbw,
                      float* prob, int64_t cap_blocks, int64_t* n_blocks);

/* ---- table / pool level operations (the reference exercises these from its unit tests,
 * utils/tests/voxel_hash_test.cu:36-54, through 1-block kernels) ----------------------------- */

/* Hash (voxel_hash.cu:19-23); mask = NUM_BUCKET-1 = 2097151 in the reference. */
RSF_API uint32_t rsf_hash(int16_t x, int16_t y, int16_t z, uint32_t mask);
/* Owner rank of a block under spatial sharding (new). */
RSF_API uint32_t rsf_shard_owner(int16_t x, int16_t y, int16_t z, int shift, int world);
/* VoxelHashTable::Allocate for n block coordinates (n x 3 i16, host), one launch. */
RSF_API int rsf_allocate_blocks(rsf_grid_t g, const int16_t* block_coords, int n);
/* VoxelHashTable::Delete for n block coordinates. */
RSF_API int rsf_delete_blocks(rsf_grid_t g, const int16_t* block_coords, int n);
/* Table hygiene (new).  Carved blocks leave their key behind ("ghost"; VoxelHashTable::Delete, voxel_hash.cu:110-159,
 * unlinks entries instead), so probe chains grow on a long session.  The library rebuilds the table from the alive-block
 * list by itself, between two frames, when ghosts exceed 1/8 of the slots or the load 3/4 (rsf_frame_stats.table_keys /
 * .table_rebuilds report it; a frame that returned RSF_ERR_TABLE_FULL forces one, so that error is recoverable).  This
 * call forces a rebuild now.  Block contents and pool indices are untouched. */
RSF_API int rsf_rebuild_table(rsf_grid_t g);
/* Rebuild as soon as more than max_ghosts ghost keys are known to exist (capped at slots / 8; 0 = the default). */
RSF_API int rsf_set_table_policy(rsf_grid_t g, int64_t max_ghosts);
/* VoxelHashTable::Retrieve<VoxelTSDF|VoxelRGBW|VoxelSEGM> at n voxel coordinates (n x 3 i16).
 * found[i] = 1 if the voxel's block is allocated, else 0 and the reference's miss defaults are
 * returned (tsdf -10, rgbw 0, prob 0; voxel_types.cu:3,8,11).  Outputs may be NULL. */
RSF_API int rsf_retrieve_voxels(rsf_grid_t g, const int16_t* voxel_coords, int n, float* tsdf, uint8_t* rgbw,
                        float* prob, uint8_t* found);
/* *RetrieveMutable<...>(p) = value at n voxel coordinates; fields may be NULL (left untouched).
 * written[i] = 0 when the voxel's block is not allocated. */
RSF_API int rsf_assign_voxels(rsf_grid_t g, const int16_t* voxel_coords, int n, const float* tsdf,
                      const uint8_t* rgbw, const float* prob, uint8_t* written);
/* VoxelHashTable::RetrieveTSDF (voxel_hash.cu:161-188) at n continuous grid points (n x 3 f32). */
RSF_API int rsf_retrieve_tsdf_trilinear(rsf_grid_t g, const float* points, int n, float* tsdf);

/* ---- measurement hooks (no reference equivalent; the reference only logs wall-clock ms,
 * modules/tsdf_module.cc:108-112) --------------------------------------------------------------
 * rsf_profile_enable(g, 1) brackets every subsequent kernel launch with a CUDA event pair on the
 * grid's stream; rsf_profile_read synchronises and returns, per kernel, the summed duration [ms]
 * and launch count since the last read / enable. */
enum rsf_prof_kernel { RSF_PROF_ALLOCATE = 0, RSF_PROF_SELECT = 1, RSF_PROF_INTEGRATE = 2, RSF_PROF_RAYCAST = 3,
                       RSF_PROF_MERGE = 4 /* sharded frame, presenting rank: wait for all arrival flags + unpack */, RSF_PROF_NUM = 5 };
RSF_API int rsf_profile_enable(rsf_grid_t g, int on);
RSF_API int rsf_profile_read(rsf_grid_t g, double ms[RSF_PROF_NUM], int64_t launches[RSF_PROF_NUM]);
/* Running totals since creation: {frames, new_blocks, visible_blocks, updated_voxels, carved_blocks}. */
RSF_API int rsf_get_totals(rsf_grid_t g, int64_t totals[5]);

/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost) so callers can hand rsf_integrate /
 * rsf_raycast page-locked buffers and get true asynchronous DMA. */
RSF_API int rsf_host_alloc(void** ptr, size_t bytes);
RSF_API int rsf_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RSFUSION_H_ */

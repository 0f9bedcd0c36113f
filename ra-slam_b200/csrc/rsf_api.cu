// rsf_api.cu -- host side of the C ABI declared in include/rsfusion.h.
//
// Orchestration that replaces TSDFGrid's host methods (utils/tsdf/voxel_tsdf.cu:376-559,
// :736-902).  Per Integrate the reference issues ~10 launches, 4 pageable H2D copies on two
// streams, 4 host synchronisations and 2 synchronous D2H counter reads; here a
// frame is 4 H2D copies + 1 memset + 3 launches on one stream and, for the blocking host entry
// point, one 128-byte counter read-back and one synchronisation.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "rsf_kernels.h"
#include "rsfusion.h"

using namespace rsf;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define RSF_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(RSF_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

struct rsf_grid_s {
  rsf_options opt{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  GridView view{};
  size_t table_slots = 0;
  size_t max_pixels = 0;
  // Host entry points run a 3-stage pipeline: H2D on `s_h2d` -> kernels on `stream` -> D2H on
  // `s_d2h`, chained with events and kSlots-deep, so the copy of frame k+1 and the read-back
  // of frame k-1 overlap the kernels of frame k (the reference copies, computes and reads back
  // serially on one stream, voxel_tsdf.cu:433-451).
  static constexpr int kSlots = 4;   // frames in flight on the pipelined host path (3 hide H2D + compute + D2H latency)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaStream_t s_side = nullptr;                      // select_visible runs here, next to allocate
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  struct InSlot {
    uint8_t* rgb = nullptr;
    float *depth = nullptr, *ht = nullptr, *lt = nullptr;
    cudaEvent_t ready = nullptr;   // H2D done (recorded on s_h2d)
    cudaEvent_t free_ = nullptr;   // allocate kernel has consumed the slot (recorded on stream)
  } in[kSlots];
  struct OutSlot {
    uchar4 *rgba = nullptr, *normal = nullptr;
    float *depth = nullptr, *prob = nullptr;
    cudaEvent_t ready = nullptr;   // raycast done (recorded on stream)
    cudaEvent_t done = nullptr;    // D2H done (recorded on s_d2h)
  } out[kSlots];
  unsigned in_next = 0, out_next = 0;
  float4* d_frame = nullptr;       // staged frame records (written by allocate, read by integrate)
  float* d_tiles = nullptr;        // per-16x16-tile maximum usable depth (occlusion culling in select_visible)
  int int_ctas_per_sm = 8;         // resident integrate_kernel CTAs per SM on this grid's device
  unsigned* d_ranges = nullptr;   // per-tile depth ranges of the raycast pre-pass
  float* class_pool = nullptr;    // K-class extension: [pool_blocks][K][512] un-normalised log-probabilities
  // pinned mirror of the device counters
  Counters* h_ctr = nullptr;
  rsf_frame_stats last{};
  // table hygiene: the host's (possibly a few frames old) knowledge of the table load, see maybe_rebuild_table
  int64_t table_rebuilds = 0;
  Counters* h_ctr_snap = nullptr;   // pinned; filled by an asynchronous copy every kSnapEvery frames (no host sync)
  cudaEvent_t ev_alloc_done = nullptr, ev_snap = nullptr;
  cudaEvent_t ev_prepass_go = nullptr, ev_prepass = nullptr;   // fused frame: raycast pre-pass on the side stream, next to integrate
  bool snap_pending = false, snap_stale = false, force_rebuild = false;
  int64_t known_keys = 0, known_alive = 0;
  int64_t max_ghosts = 0;           // rebuild threshold (rsf_set_table_policy); 0 = table_slots / 8
  unsigned frame_no = 0;
  // sharded frame (rsf_shard_frame_device): the presenting rank's planes / flags as mapped into this process
  rsf_shard_link link{};
  bool linked = false;
  unsigned shard_seq = 0;
  // in-library frame transport (rsf_shard_feed_device / rsf_shard_step_device)
  cudaStream_t kid_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_step_done[2] = {nullptr, nullptr};   // recorded on `stream` after the step of that frame parity
  unsigned fed = 0, stepped = 0;
  const uint8_t* fed_src[2] = {nullptr, nullptr};     // ingest rank: the caller's packed frame of that parity
  // export staging (rsf_gather_*): a ring of device chunks and pinned host chunks, allocated on first use and kept
  static constexpr int kExpRing = 3;
  static constexpr size_t kExpChunk = (size_t)4 << 20;
  struct ExpSlot {
    uint8_t *d = nullptr, *h = nullptr;
    cudaEvent_t filled = nullptr;   // download kernel done (recorded on stream)
    cudaEvent_t copied = nullptr;   // D2H done (recorded on s_d2h)
  } exp_ring[kExpRing];
  // optional per-kernel CUDA-event timing (rsf_profile_*)
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[RSF_PROF_NUM];
  size_t prof_used[RSF_PROF_NUM] = {};
};

namespace {

int check_grid(rsf_grid_t g) {
  if (!g) return fail(RSF_ERR_INVALID, "null grid handle");
  return RSF_OK;
}

int make_frame(rsf_grid_t g, int H, int W, float max_depth, const float* K, const float* pose, Frame* f) {
  if (!K || !pose) return fail(RSF_ERR_INVALID, "K / pose must not be NULL");
  if (H <= 0 || W <= 0 || (size_t)H * (size_t)W > g->max_pixels)
    return fail(RSF_ERR_INVALID, "image %dx%d outside (0, max_pixels=%zu]", W, H, g->max_pixels);
  // the kernels divide by truncation / max_depth through a shared refined reciprocal: keep both well inside
  // the normal range (any physical value is)
  if (!(max_depth >= 1e-6f && max_depth <= 1e6f)) return fail(RSF_ERR_INVALID, "max_depth %g outside [1e-6, 1e6] m", (double)max_depth);
  f->K = Intr{K[0], K[1], K[2], K[3]};
  f->Ki = f->K.inverse();                       // CameraParams ctor, camera.cuh:61-67
  f->T_cw = Pose{Quat{pose[0], pose[1], pose[2], pose[3]}, V3{pose[4], pose[5], pose[6]}};
  f->T_wc = f->T_cw.inverse();                  // cam_T_world.Inverse() on the host, voxel_tsdf.cu:459,892
  f->W = W;
  f->H = H;
  f->vs = g->opt.voxel_size;
  f->tr = g->opt.truncation;
  f->max_depth = max_depth;
  return RSF_OK;
}

// Read the counters back (blocking).  report = true (the integrate / allocate / synchronize entry points): translate and
// clear the sticky device error bits.  report = false (exports, statistics): leave them for the next reporting call -- a
// gather that succeeded must not fail because an earlier frame overflowed the pool.
int sync_counters(rsf_grid_t g, bool report = true) {
  RSF_CUDA(cudaMemcpyAsync(g->h_ctr, g->view.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  const Counters& c = *g->h_ctr;
  g->last.active_blocks = (int64_t)g->view.pool_blocks - c.free_count;
  g->last.high_water = c.high_water;
  g->last.new_blocks = (int64_t)c.new_blocks;
  g->last.visible_blocks = (int64_t)c.vis_count + c.culled;   // culled blocks are visible in the reference's sense
  g->last.updated_voxels = (int64_t)c.updated_voxels;
  g->last.carved_blocks = (int64_t)c.carved_blocks;
  g->last.culled_blocks = c.culled;
  g->last.table_keys = c.table_keys;
  g->last.table_rebuilds = g->table_rebuilds;
  g->known_keys = c.table_keys;
  g->known_alive = g->last.active_blocks;
  if (c.error && (c.error & kErrTableFull)) g->force_rebuild = true;   // the next frame starts from a table without ghosts
  if (c.error && report) {
    const unsigned e = c.error;
    cudaMemsetAsync(&g->view.ctr->error, 0, sizeof(unsigned), g->stream);
    if (e & kErrPoolFull)
      return fail(RSF_ERR_POOL_FULL, "voxel-block pool exhausted (%d blocks); raise rsf_options.pool_blocks",
                  g->view.pool_blocks);
    if (e & kErrTableFull) return fail(RSF_ERR_TABLE_FULL, "hash table full (%zu slots)", g->table_slots);
  }
  return RSF_OK;
}

// RAII: brackets one kernel launch with a CUDA event pair when profiling is on
struct ProfScope {
  rsf_grid_t g;
  cudaEvent_t stop = nullptr;
  cudaStream_t st;
  ProfScope(rsf_grid_t grid, int which, cudaStream_t on = nullptr) : g(grid), st(on ? on : grid->stream) {
    if (!g->profiling) return;
    auto& v = g->prof_events[which];
    size_t& used = g->prof_used[which];
    if (used == v.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      v.emplace_back(a, b);
    }
    cudaEventRecord(v[used].first, st);
    stop = v[used].second;
    ++used;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

// Table hygiene (see rebuild_table_kernel): ghosts = claimed slots whose block has been carved.  The table is rebuilt
// between two frames when ghosts exceed 1/8 of the slots or the load passes 3/4 -- with slots >= 2 x pool blocks the load
// therefore never exceeds 1/2 + 1/8 and RSF_ERR_TABLE_FULL cannot become permanent (a frame that did hit it forces a
// rebuild before the next one).  The decision uses the newest counters the host has WITHOUT synchronising: the blocking
// entry points refresh them on every call (sync_counters), the asynchronous ones through a 128-byte snapshot copied
// on the side stream every kSnapEvery frames.
constexpr unsigned kSnapEvery = 16;
int rebuild_table_now(rsf_grid_t g) {
  launch_rebuild_table(g->view, g->table_slots, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  ++g->table_rebuilds;
  g->known_keys = g->known_alive;
  g->force_rebuild = false;
  if (g->snap_pending) g->snap_stale = true;   // taken before the rebuild: ignore it
  return RSF_OK;
}
int maybe_rebuild_table(rsf_grid_t g) {
  if (g->snap_pending && cudaEventQuery(g->ev_snap) == cudaSuccess) {
    g->snap_pending = false;
    if (!g->snap_stale) {
      g->known_keys = g->h_ctr_snap->table_keys;
      g->known_alive = (int64_t)g->view.pool_blocks - g->h_ctr_snap->free_count;
    }
    g->snap_stale = false;
  }
  const int64_t slots = (int64_t)g->table_slots, ghosts = g->known_keys - g->known_alive;
  const int64_t limit = g->max_ghosts > 0 ? std::min(g->max_ghosts, slots / 8) : slots / 8;
  if (g->force_rebuild || ghosts > limit || g->known_keys > slots / 4 * 3) return rebuild_table_now(g);
  return RSF_OK;
}
// after this frame's allocate kernel (the only kernel that claims slots) has been enqueued on g->stream
int snapshot_counters(rsf_grid_t g) {
  if (g->snap_pending || (g->frame_no++ % kSnapEvery) != 0) return RSF_OK;
  RSF_CUDA(cudaEventRecord(g->ev_alloc_done, g->stream));
  RSF_CUDA(cudaStreamWaitEvent(g->s_side, g->ev_alloc_done, 0));
  RSF_CUDA(cudaMemcpyAsync(g->h_ctr_snap, g->view.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, g->s_side));
  RSF_CUDA(cudaEventRecord(g->ev_snap, g->s_side));
  g->snap_pending = true;
  return RSF_OK;
}

// u16 = true: depth / ht / lt point to uint16 images, converted in the allocate kernel with the two scale factors
int enqueue_integrate(rsf_grid_t g, const Frame& f, const uint8_t* rgb, const void* depth, const void* ht,
                      const void* lt, cudaEvent_t inputs_consumed = nullptr, bool u16 = false, float depth_alpha = 1.f,
                      float prob_alpha = 1.f, const float* class_prob = nullptr, const Frame* prepass_for = nullptr) {
  if (int rc = maybe_rebuild_table(g)) return rc;
  launch_begin_frame(g->view, g->stream);
  // allocate (pixels -> new blocks, appended to the visible list as *fresh*) and select_visible (blocks that were
  // alive before this frame -> visible list) are independent: select skips every slot that is dead or fresh, and a
  // pool slot only ever goes dead -> fresh during allocate, so it runs on a side stream next to allocate instead of
  // after it.  Both append through the same atomic counter.
  RSF_CUDA(cudaEventRecord(g->ev_fork, g->stream));
  RSF_CUDA(cudaStreamWaitEvent(g->s_side, g->ev_fork, 0));
  {
    ProfScope p(g, RSF_PROF_SELECT, g->s_side);
    launch_depth_tiles(depth, u16, depth_alpha, f.W, f.H, f.max_depth, g->d_tiles, g->s_side);
    launch_select_visible(g->view, f, g->d_tiles, g->sm_count, g->s_side);
  }
  RSF_CUDA(cudaEventRecord(g->ev_join, g->s_side));
  {
    ProfScope p(g, RSF_PROF_ALLOCATE);
    if (u16)
      launch_allocate_u16(g->view, f, rgb, (const uint16_t*)depth, (const uint16_t*)ht, (const uint16_t*)lt, depth_alpha,
                          prob_alpha, g->d_frame, g->stream);
    else
      launch_allocate(g->view, f, rgb, (const float*)depth, (const float*)ht, (const float*)lt, g->d_frame, g->stream);
  }
  // the allocate kernel is the only reader of the raw images (it stages them into d_frame)
  if (inputs_consumed) RSF_CUDA(cudaEventRecord(inputs_consumed, g->stream));
  if (prepass_for) {
    // Fused frame (rsf_shard_frame_device): the raycast's pre-pass for the view `prepass_for` runs on the side stream
    // next to the integrate kernel instead of after it.  It needs the blocks allocate has just added; the blocks
    // integrate is about to carve may or may not still be in the list it walks -- either way the per-tile ranges only
    // get wider (a carved block is a miss for the march), so the rendered image is unchanged.
    RSF_CUDA(cudaEventRecord(g->ev_prepass_go, g->stream));
    RSF_CUDA(cudaStreamWaitEvent(g->s_side, g->ev_prepass_go, 0));
    launch_raycast_prepass(g->view, *prepass_for, g->opt.truncation / 2, g->d_ranges, g->sm_count, g->s_side);
    RSF_CUDA(cudaEventRecord(g->ev_prepass, g->s_side));
  }
  RSF_CUDA(cudaStreamWaitEvent(g->stream, g->ev_join, 0));
  if (int rc = snapshot_counters(g)) return rc;
  // K-class extension: the class planes are updated with the weights from BEFORE this frame's TSDF update
  if (class_prob)
    launch_class_update(g->view, f, g->d_frame, class_prob, g->opt.num_classes, g->class_pool, g->sm_count, g->stream);
  {
    ProfScope p(g, RSF_PROF_INTEGRATE);
    launch_integrate(g->view, f, g->d_frame, g->sm_count, g->int_ctas_per_sm, g->stream);
  }
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

// Export of `nblk` gathered blocks, `block_bytes` of rows each, to host memory `dst`, in chunks of kExpChunk through a
// persistent ring (no allocation per call): the download kernel of chunk i + 1 runs on the grid's stream while chunk i
// crosses PCIe on the read-back stream.  A pinned / registered `dst` receives the DMA directly; a pageable one (the
// reference's std::vector, a numpy array) is filled by the host from the pinned ring while later chunks are in flight --
// a cudaMemcpy to pageable memory would stage through the driver's own bounce buffer, serially.
// fill(blk0, n, d_rows): enqueue the kernel that writes the rows of gathered blocks [blk0, blk0 + n) to d_rows.
template <class Fill>
int export_chunked(rsf_grid_t g, int nblk, size_t block_bytes, uint8_t* dst, Fill fill) {
  if (block_bytes == 0 || block_bytes > rsf_grid_s::kExpChunk) return fail(RSF_ERR_INVALID, "export: %zu-byte blocks do not fit a chunk", block_bytes);
  for (auto& e : g->exp_ring) {
    if (!e.d) RSF_CUDA(cudaMalloc((void**)&e.d, rsf_grid_s::kExpChunk));
    if (!e.h) RSF_CUDA(cudaHostAlloc((void**)&e.h, rsf_grid_s::kExpChunk, cudaHostAllocDefault));
    if (!e.filled) RSF_CUDA(cudaEventCreateWithFlags(&e.filled, cudaEventDisableTiming));
    if (!e.copied) RSF_CUDA(cudaEventCreateWithFlags(&e.copied, cudaEventDisableTiming));
  }
  cudaPointerAttributes attr{};
  const bool direct = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();   // (older drivers report an error for plain malloc'ed memory)
  const int per_chunk = (int)(rsf_grid_s::kExpChunk / block_bytes);
  const int nchunks = (nblk + per_chunk - 1) / per_chunk;
  constexpr int R = rsf_grid_s::kExpRing;
  auto issue = [&](int i) -> int {
    auto& e = g->exp_ring[i % R];
    const int b0 = i * per_chunk, n = std::min(per_chunk, nblk - b0);
    if (i >= R) RSF_CUDA(cudaStreamWaitEvent(g->stream, e.copied, 0));        // the chunk's previous contents have left
    fill(b0, n, e.d);
    RSF_CUDA(cudaGetLastError());
    RSF_CUDA(cudaEventRecord(e.filled, g->stream));
    RSF_CUDA(cudaStreamWaitEvent(g->s_d2h, e.filled, 0));
    uint8_t* to = direct ? dst + (size_t)b0 * block_bytes : e.h;
    RSF_CUDA(cudaMemcpyAsync(to, e.d, (size_t)n * block_bytes, cudaMemcpyDeviceToHost, g->s_d2h));
    RSF_CUDA(cudaEventRecord(e.copied, g->s_d2h));
    return RSF_OK;
  };
  if (direct) {
    for (int i = 0; i < nchunks; ++i)
      if (int rc = issue(i)) return rc;
    RSF_CUDA(cudaStreamSynchronize(g->s_d2h));
    return RSF_OK;
  }
  for (int i = 0; i < std::min(nchunks, R); ++i)
    if (int rc = issue(i)) return rc;
  for (int i = 0; i < nchunks; ++i) {
    auto& e = g->exp_ring[i % R];
    const int b0 = i * per_chunk, n = std::min(per_chunk, nblk - b0);
    RSF_CUDA(cudaEventSynchronize(e.copied));
    std::memcpy(dst + (size_t)b0 * block_bytes, e.h, (size_t)n * block_bytes);
    if (i + R < nchunks)
      if (int rc = issue(i + R)) return rc;
  }
  return RSF_OK;
}

int gather_rows(rsf_grid_t g, int mode, const float* bounds, int ncol, float* rows, int64_t cap_rows, int64_t* n_rows) {
  if (int rc = check_grid(g)) return rc;
  if (!n_rows) return fail(RSF_ERR_INVALID, "n_rows must not be NULL");
  DeviceGuard dg(g->device);
  S3 lo{0, 0, 0}, hi{0, 0, 0};
  if (mode == 1) {
    if (!bounds) return fail(RSF_ERR_INVALID, "bounds must not be NULL");
    // BoundingCube<float>::Scale<short>(1./voxel_size), voxel_tsdf.cuh:28-33, voxel_tsdf.cu:534
    const float s = (float)(1. / g->opt.voxel_size);
    lo = S3{(short)(bounds[0] * s), (short)(bounds[2] * s), (short)(bounds[4] * s)};
    hi = S3{(short)(bounds[1] * s), (short)(bounds[3] * s), (short)(bounds[5] * s)};
  }
  RSF_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(g->view.ctr) + kCtrGatherOff, 0, sizeof(int), g->stream));
  launch_gather_select(g->view, mode, g->opt.shard_world > 1, lo, hi, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (int rc = sync_counters(g, false)) return rc;
  const int nblk = g->h_ctr->gather_count;
  *n_rows = (int64_t)nblk * kBlockVol;
  if (!rows) return RSF_OK;
  if (cap_rows < *n_rows) return fail(RSF_ERR_CAPACITY, "need %lld rows, capacity %lld", (long long)*n_rows, (long long)cap_rows);
  if (nblk == 0) return RSF_OK;
  return export_chunked(g, nblk, (size_t)kBlockVol * ncol * sizeof(float), reinterpret_cast<uint8_t*>(rows),
                        [&](int b0, int n, uint8_t* d) {
                          launch_download_rows(g->view, g->opt.voxel_size, b0, n, ncol, reinterpret_cast<float*>(d), g->sm_count, g->stream);
                        });
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  cudaStream_t s;
  explicit DevBuf(cudaStream_t st) : s(st) {}
  cudaError_t alloc(size_t n) { return cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), s); }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, s);
  }
};

}  // namespace

extern "C" {

void rsf_default_options(rsf_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->voxel_size = 0.01f;
  o->truncation = 0.06f;
  o->device = -1;
  o->pool_blocks = 1 << 18;          // NUM_BLOCK, voxel_mem.cuh:11-13
  o->table_slots_log2 = 22;          // NUM_ENTRY, voxel_hash.cuh:13-25
  o->max_pixels = 1920 * 1080;       // MAX_IMG_SIZE, voxel_tsdf.cu:11-13
  o->shard_rank = 0;
  o->shard_world = 1;
  o->shard_shift = 4;
  o->shard_halo = 1;
}

int rsf_abi_version(void) { return RSF_ABI_VERSION; }
const char* rsf_last_error(void) { return g_last_error.c_str(); }

int rsf_create(const rsf_options* opts, rsf_grid_t* out) {
  if (!opts || !out) return fail(RSF_ERR_INVALID, "opts / out must not be NULL");
  *out = nullptr;
  rsf_options o = *opts;
  rsf_options d;
  rsf_default_options(&d);
  if (o.pool_blocks <= 0) o.pool_blocks = d.pool_blocks;
  if (o.table_slots_log2 <= 0) o.table_slots_log2 = d.table_slots_log2;
  if (o.max_pixels <= 0) o.max_pixels = d.max_pixels;
  if (o.shard_world <= 0) o.shard_world = 1;
  if (o.shard_shift < 0) o.shard_shift = d.shard_shift;
  if (o.shard_halo < 0) o.shard_halo = d.shard_halo;
  if (!(o.voxel_size >= 1e-6f && o.voxel_size <= 1e6f) || !(o.truncation >= 1e-6f && o.truncation <= 1e6f))
    return fail(RSF_ERR_INVALID, "voxel_size and truncation must lie in [1e-6, 1e6] m");
  if (o.table_slots_log2 > 30 || ((size_t)1 << o.table_slots_log2) < 2 * (size_t)o.pool_blocks)
    return fail(RSF_ERR_INVALID, "table_slots_log2=%d must give >= 2x pool_blocks=%d slots", o.table_slots_log2, o.pool_blocks);
  if (o.shard_rank < 0 || o.shard_rank >= o.shard_world) return fail(RSF_ERR_INVALID, "shard_rank outside [0, shard_world)");
  // shard_wanted shifts by shard_shift and loops over (2 halo + 1)^3 neighbours: keep both small (int16 block coordinates)
  if (o.shard_shift > 12 || o.shard_halo > 4) return fail(RSF_ERR_INVALID, "shard_shift must be <= 12 and shard_halo <= 4");
  if (o.flags & ~(unsigned)RSF_FLAG_EXTERNAL_STREAM) return fail(RSF_ERR_INVALID, "unknown bits in flags");
  if (o.num_classes == 2) o.num_classes = 0;
  if (o.num_classes != 0 && (o.num_classes < 3 || o.num_classes > 64)) return fail(RSF_ERR_INVALID, "num_classes must be 0, 2 or 3..64");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RSF_ERR_CUDA, "no CUDA device (%s): rsfusion has no CPU fallback", cudaGetErrorString(e));
  int dev = o.device;
  if (dev < 0) RSF_CUDA(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(RSF_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev);
  cudaDeviceProp prop;
  RSF_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(RSF_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);

  DeviceGuard dg(dev);
  rsf_grid_s* g = new (std::nothrow) rsf_grid_s;
  if (!g) return fail(RSF_ERR_INVALID, "out of host memory");
  g->opt = o;
  g->device = dev;
  g->sm_count = prop.multiProcessorCount;
  g->table_slots = (size_t)1 << o.table_slots_log2;
  g->max_pixels = (size_t)o.max_pixels;
  if (o.stream || (o.flags & RSF_FLAG_EXTERNAL_STREAM)) {
    g->stream = (cudaStream_t)o.stream;
  } else {
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete g;
      return fail(RSF_ERR_CUDA, "cudaStreamCreate failed");
    }
    g->own_stream = true;
  }
  GridView& v = g->view;
  v.slot_mask = (uint32_t)(g->table_slots - 1);
  v.pool_blocks = o.pool_blocks;
  v.shard = ShardSpec{o.shard_rank, o.shard_world, o.shard_shift, o.shard_halo};
  const size_t P = g->max_pixels;
  bool ok = true;
  auto A = [&](void** p, size_t bytes) { ok = ok && cudaMalloc(p, bytes) == cudaSuccess; };
  A((void**)&v.table, g->table_slots * sizeof(TableEntry));
  A((void**)&v.pool, (size_t)o.pool_blocks * kSlabBytes);
  A((void**)&v.meta, (size_t)o.pool_blocks * sizeof(BlockMeta));
  A((void**)&v.heap, (size_t)o.pool_blocks * sizeof(int));
  A((void**)&v.vis_list, (size_t)o.pool_blocks * sizeof(VisEntry));
  A((void**)&v.gather_list, (size_t)o.pool_blocks * sizeof(int));
  A((void**)&v.ctr, sizeof(Counters));
  A((void**)&g->d_frame, P * sizeof(float4));
  A((void**)&g->d_tiles, (P / 16 + 4096) * sizeof(float));   // ceil(W/16) ceil(H/16) <= P/16 + 1 for any aspect ratio
  g->int_ctas_per_sm = integrate_ctas_per_sm();
  for (auto& sl : g->in) {
    A((void**)&sl.rgb, P * 3);
    A((void**)&sl.depth, P * sizeof(float));
    A((void**)&sl.ht, P * sizeof(float));
    A((void**)&sl.lt, P * sizeof(float));
    ok = ok && cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&sl.free_, cudaEventDisableTiming) == cudaSuccess;
  }
  for (auto& sl : g->out) {
    A((void**)&sl.rgba, P * sizeof(uchar4));
    A((void**)&sl.normal, P * sizeof(uchar4));
    A((void**)&sl.depth, P * sizeof(float));
    A((void**)&sl.prob, P * sizeof(float));
    ok = ok && cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaStreamCreateWithFlags(&g->s_h2d, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&g->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&g->s_side, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming) == cudaSuccess;
  A((void**)&g->d_ranges, (P / 16 + 4096) * 4 * sizeof(unsigned) + 64);   // tiles <= ceil(W/16) ceil(H/16) <= P/16 + ... for any aspect ratio
  if (o.num_classes > 2) A((void**)&g->class_pool, (size_t)o.pool_blocks * o.num_classes * kBlockVol * sizeof(float));
  ok = ok && cudaHostAlloc((void**)&g->h_ctr, sizeof(Counters), cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaHostAlloc((void**)&g->h_ctr_snap, sizeof(Counters), cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_alloc_done, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_prepass_go, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_prepass, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&g->ev_snap, cudaEventDisableTiming) == cudaSuccess;
  if (ok) {
    // key = kEmptyKey (all ones), val = kValFree (-1)
    ok = ok && cudaMemsetAsync(v.table, 0xff, g->table_slots * sizeof(TableEntry), g->stream) == cudaSuccess;
    launch_init_pool(v, g->stream);
    ok = ok && cudaStreamSynchronize(g->stream) == cudaSuccess;
  }
  if (!ok) {
    const cudaError_t le = cudaGetLastError();
    rsf_destroy(g);
    return fail(RSF_ERR_CUDA, "device allocation / initialisation failed: %s", cudaGetErrorString(le));
  }
  *out = g;
  return RSF_OK;
}

int rsf_destroy(rsf_grid_t g) {
  if (!g) return RSF_OK;
  DeviceGuard dg(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  GridView& v = g->view;
  if (g->s_h2d) cudaStreamSynchronize(g->s_h2d);
  if (g->s_d2h) cudaStreamSynchronize(g->s_d2h);
  if (g->s_side) cudaStreamSynchronize(g->s_side);
  void* ptrs[] = {v.table, v.pool, v.meta, v.heap, v.vis_list, v.gather_list, v.ctr, g->d_frame, g->d_tiles, g->d_ranges, g->class_pool};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& sl : g->in) {
    void* q[] = {sl.rgb, sl.depth, sl.ht, sl.lt};
    for (void* p : q)
      if (p) cudaFree(p);
    if (sl.ready) cudaEventDestroy(sl.ready);
    if (sl.free_) cudaEventDestroy(sl.free_);
  }
  for (auto& sl : g->out) {
    void* q[] = {sl.rgba, sl.normal, sl.depth, sl.prob};
    for (void* p : q)
      if (p) cudaFree(p);
    if (sl.ready) cudaEventDestroy(sl.ready);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  for (auto& st : g->kid_stream)
    if (st) cudaStreamDestroy(st);
  for (auto& e : g->ev_step_done)
    if (e) cudaEventDestroy(e);
  if (g->s_h2d) cudaStreamDestroy(g->s_h2d);
  if (g->s_d2h) cudaStreamDestroy(g->s_d2h);
  if (g->s_side) cudaStreamDestroy(g->s_side);
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->ev_join) cudaEventDestroy(g->ev_join);
  for (auto& e : g->exp_ring) {
    if (e.d) cudaFree(e.d);
    if (e.h) cudaFreeHost(e.h);
    if (e.filled) cudaEventDestroy(e.filled);
    if (e.copied) cudaEventDestroy(e.copied);
  }
  if (g->h_ctr) cudaFreeHost(g->h_ctr);
  if (g->h_ctr_snap) cudaFreeHost(g->h_ctr_snap);
  if (g->ev_alloc_done) cudaEventDestroy(g->ev_alloc_done);
  if (g->ev_prepass_go) cudaEventDestroy(g->ev_prepass_go);
  if (g->ev_prepass) cudaEventDestroy(g->ev_prepass);
  if (g->ev_snap) cudaEventDestroy(g->ev_snap);
  for (auto& v : g->prof_events)
    for (auto& e : v) {
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
  if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return RSF_OK;
}

int rsf_integrate_async(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H,
                        int W, float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!rgb || !depth || !ht || !lt) return fail(RSF_ERR_INVALID, "rgb / depth / ht / lt must not be NULL");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  const size_t P = (size_t)H * W;
  auto& sl = g->in[g->in_next++ % rsf_grid_s::kSlots];
  RSF_CUDA(cudaStreamWaitEvent(g->s_h2d, sl.free_, 0));       // previous user of this slot has consumed it
  RSF_CUDA(cudaMemcpyAsync(sl.depth, depth, P * sizeof(float), cudaMemcpyHostToDevice, g->s_h2d));
  RSF_CUDA(cudaMemcpyAsync(sl.rgb, rgb, P * 3, cudaMemcpyHostToDevice, g->s_h2d));
  RSF_CUDA(cudaMemcpyAsync(sl.ht, ht, P * sizeof(float), cudaMemcpyHostToDevice, g->s_h2d));
  RSF_CUDA(cudaMemcpyAsync(sl.lt, lt, P * sizeof(float), cudaMemcpyHostToDevice, g->s_h2d));
  RSF_CUDA(cudaEventRecord(sl.ready, g->s_h2d));
  RSF_CUDA(cudaStreamWaitEvent(g->stream, sl.ready, 0));
  if (int rc = enqueue_integrate(g, f, sl.rgb, sl.depth, sl.ht, sl.lt, sl.free_)) return rc;
  return RSF_OK;
}

int rsf_integrate(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H, int W,
                  float max_depth, const float K[4], const float pose[7]) {
  if (int rc = rsf_integrate_async(g, rgb, depth, ht, lt, H, W, max_depth, K, pose)) return rc;
  DeviceGuard dg(g->device);
  return sync_counters(g);
}

int rsf_integrate_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H,
                         int W, float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!rgb || !depth || !ht || !lt) return fail(RSF_ERR_INVALID, "rgb / depth / ht / lt must not be NULL");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  return enqueue_integrate(g, f, rgb, depth, ht, lt);
}

int rsf_integrate_classes_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* class_prob, int H, int W,
                                 float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!g->class_pool) return fail(RSF_ERR_INVALID, "grid was created without rsf_options.num_classes > 2");
  if (!rgb || !depth || !class_prob) return fail(RSF_ERR_INVALID, "rgb / depth / class_prob must not be NULL");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  // binary channel: no ht / lt images -> the staged observation is ln 1 - ln 1 = 0 exactly and the prior is kept
  // (feeding ht = lt = a class plane would give ln 0 - ln 0 = NaN wherever that plane is 0, e.g. one-hot maps)
  return enqueue_integrate(g, f, rgb, depth, nullptr, nullptr, nullptr, false, 1.f, 1.f, class_prob);
}

int rsf_integrate_classes(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* class_prob, int H, int W,
                          float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!g->class_pool) return fail(RSF_ERR_INVALID, "grid was created without rsf_options.num_classes > 2");
  if (!rgb || !depth || !class_prob) return fail(RSF_ERR_INVALID, "rgb / depth / class_prob must not be NULL");
  if (H <= 0 || W <= 0 || (size_t)H * (size_t)W > g->max_pixels) return fail(RSF_ERR_INVALID, "image size outside (0, max_pixels]");
  DeviceGuard dg(g->device);
  const size_t P = (size_t)H * W;
  DevBuf<uint8_t> d_rgb(g->stream);
  DevBuf<float> d_depth(g->stream), d_cls(g->stream);
  RSF_CUDA(d_rgb.alloc(P * 3));
  RSF_CUDA(d_depth.alloc(P));
  RSF_CUDA(d_cls.alloc(P * g->opt.num_classes));
  RSF_CUDA(cudaMemcpyAsync(d_rgb.p, rgb, P * 3, cudaMemcpyHostToDevice, g->stream));
  RSF_CUDA(cudaMemcpyAsync(d_depth.p, depth, P * sizeof(float), cudaMemcpyHostToDevice, g->stream));
  RSF_CUDA(cudaMemcpyAsync(d_cls.p, class_prob, P * g->opt.num_classes * sizeof(float), cudaMemcpyHostToDevice, g->stream));
  if (int rc = rsf_integrate_classes_device(g, d_rgb.p, d_depth.p, d_cls.p, H, W, max_depth, K, pose)) return rc;
  return sync_counters(g);
}

int rsf_gather_valid_classes(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows) {
  if (int rc = check_grid(g)) return rc;
  if (!g->class_pool) return fail(RSF_ERR_INVALID, "grid was created without rsf_options.num_classes > 2");
  if (!n_rows) return fail(RSF_ERR_INVALID, "n_rows must not be NULL");
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(g->view.ctr) + kCtrGatherOff, 0, sizeof(int), g->stream));
  launch_gather_select(g->view, 0, g->opt.shard_world > 1, S3{0, 0, 0}, S3{0, 0, 0}, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (int rc = sync_counters(g, false)) return rc;
  const int nblk = g->h_ctr->gather_count;
  *n_rows = (int64_t)nblk * kBlockVol;
  if (!rows) return RSF_OK;
  if (cap_rows < *n_rows) return fail(RSF_ERR_CAPACITY, "need %lld rows, capacity %lld", (long long)*n_rows, (long long)cap_rows);
  if (nblk == 0) return RSF_OK;
  const int nc = 4 + g->opt.num_classes;
  return export_chunked(g, nblk, (size_t)kBlockVol * nc * sizeof(float), reinterpret_cast<uint8_t*>(rows),
                        [&](int b0, int n, uint8_t* d) {
                          launch_download_class_rows(g->view, g->opt.voxel_size, b0, n, g->opt.num_classes, g->class_pool,
                                                     reinterpret_cast<float*>(d), g->sm_count, g->stream);
                        });
}

// cv::Mat::convertTo(CV_32FC1, alpha) multiplies in float: alpha = float(1. / factor)
static inline float u16_alpha(float factor) { return (float)(1.0 / (double)factor); }

int rsf_integrate_u16_async(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht, const uint16_t* lt,
                            int H, int W, float depth_factor, float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!rgb || !depth || ((ht == nullptr) != (lt == nullptr)))
    return fail(RSF_ERR_INVALID, "rgb / depth must not be NULL; ht and lt must be given together or both be NULL");
  if (!(depth_factor > 0.f)) return fail(RSF_ERR_INVALID, "depth_factor must be > 0");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  const size_t P = (size_t)H * W;
  auto& sl = g->in[g->in_next++ % rsf_grid_s::kSlots];
  RSF_CUDA(cudaStreamWaitEvent(g->s_h2d, sl.free_, 0));
  RSF_CUDA(cudaMemcpyAsync(sl.depth, depth, P * sizeof(uint16_t), cudaMemcpyHostToDevice, g->s_h2d));
  RSF_CUDA(cudaMemcpyAsync(sl.rgb, rgb, P * 3, cudaMemcpyHostToDevice, g->s_h2d));
  if (ht) {
    RSF_CUDA(cudaMemcpyAsync(sl.ht, ht, P * sizeof(uint16_t), cudaMemcpyHostToDevice, g->s_h2d));
    RSF_CUDA(cudaMemcpyAsync(sl.lt, lt, P * sizeof(uint16_t), cudaMemcpyHostToDevice, g->s_h2d));
  }
  RSF_CUDA(cudaEventRecord(sl.ready, g->s_h2d));
  RSF_CUDA(cudaStreamWaitEvent(g->stream, sl.ready, 0));
  return enqueue_integrate(g, f, sl.rgb, sl.depth, ht ? sl.ht : nullptr, ht ? sl.lt : nullptr, sl.free_, true,
                           u16_alpha(depth_factor), u16_alpha(65535.f));
}

int rsf_integrate_u16(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht, const uint16_t* lt, int H,
                      int W, float depth_factor, float max_depth, const float K[4], const float pose[7]) {
  if (int rc = rsf_integrate_u16_async(g, rgb, depth, ht, lt, H, W, depth_factor, max_depth, K, pose)) return rc;
  DeviceGuard dg(g->device);
  return sync_counters(g);
}

int rsf_integrate_u16_device(rsf_grid_t g, const uint8_t* rgb, const uint16_t* depth, const uint16_t* ht, const uint16_t* lt,
                             int H, int W, float depth_factor, float max_depth, const float K[4], const float pose[7]) {
  if (int rc = check_grid(g)) return rc;
  if (!rgb || !depth || ((ht == nullptr) != (lt == nullptr)))
    return fail(RSF_ERR_INVALID, "rgb / depth must not be NULL; ht and lt must be given together or both be NULL");
  if (!(depth_factor > 0.f)) return fail(RSF_ERR_INVALID, "depth_factor must be > 0");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  return enqueue_integrate(g, f, rgb, depth, ht, lt, nullptr, true, u16_alpha(depth_factor), u16_alpha(65535.f));
}

int rsf_resize_device(rsf_grid_t g, const void* src, int src_H, int src_W, void* dst, int dst_H, int dst_W, int channels,
                      int elem_bytes) {
  if (int rc = check_grid(g)) return rc;
  if (!src || !dst) return fail(RSF_ERR_INVALID, "src / dst must not be NULL");
  if (src_H <= 0 || src_W <= 0 || dst_H <= 0 || dst_W <= 0 || src_H > 32768 || src_W > 32768 || dst_H > 32768 || dst_W > 32768)
    return fail(RSF_ERR_INVALID, "image sizes must lie in [1, 32768]");
  DeviceGuard dg(g->device);
  if (launch_resize(src, src_H, src_W, dst, dst_H, dst_W, channels, elem_bytes, g->stream) != 0)
    return fail(RSF_ERR_INVALID, "unsupported format: channels=%d elem_bytes=%d (8UC3, 8UC1, 16UC1, 32FC1)", channels, elem_bytes);
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

int rsf_raycast_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7], uint8_t* rgba,
                       uint8_t* normal, float* depth, float* prob) {
  if (int rc = check_grid(g)) return rc;
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  // step_size = truncation / 2, voxel_tsdf.cu:892
  {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast(g->view, f, g->opt.truncation / 2, g->d_ranges, reinterpret_cast<uchar4*>(rgba),
                   reinterpret_cast<uchar4*>(normal), depth, prob, g->sm_count, g->stream);
  }
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

int rsf_raycast_internal_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                                const uint8_t** rgba_dev, const uint8_t** normal_dev, void** stream) {
  if (int rc = check_grid(g)) return rc;
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  auto& sl = g->out[0];
  RSF_CUDA(cudaStreamWaitEvent(g->stream, sl.done, 0));       // a pipelined read-back of this slot has finished
  {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast(g->view, f, g->opt.truncation / 2, g->d_ranges, sl.rgba, sl.normal, nullptr, nullptr, g->sm_count, g->stream);
  }
  RSF_CUDA(cudaGetLastError());
  if (rgba_dev) *rgba_dev = reinterpret_cast<const uint8_t*>(sl.rgba);
  if (normal_dev) *normal_dev = reinterpret_cast<const uint8_t*>(sl.normal);
  if (stream) *stream = (void*)g->stream;
  return RSF_OK;
}

int rsf_raycast_merge_device(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7],
                             uint64_t* merge_planes, int n_planes) {
  if (int rc = check_grid(g)) return rc;
  if (!merge_planes || n_planes < 1 || n_planes > 3) return fail(RSF_ERR_INVALID, "merge_planes must be non-NULL, 1 <= n_planes <= 3");
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast(g->view, f, g->opt.truncation / 2, g->d_ranges, nullptr, nullptr, nullptr, nullptr, g->sm_count, g->stream,
                   reinterpret_cast<unsigned long long*>(merge_planes), n_planes);
  }
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

int rsf_merge_finalize_device(rsf_grid_t g, int H, int W, uint64_t* merge_planes, int n_planes, uint8_t* rgba, uint8_t* normal,
                              float* depth, float* prob) {
  if (int rc = check_grid(g)) return rc;
  if (!merge_planes || n_planes < 1 || n_planes > 3 || H <= 0 || W <= 0)
    return fail(RSF_ERR_INVALID, "merge_planes must be non-NULL, 1 <= n_planes <= 3, H, W > 0");
  DeviceGuard dg(g->device);
  launch_merge_finalize(reinterpret_cast<unsigned long long*>(merge_planes), n_planes, H * W, reinterpret_cast<uchar4*>(rgba),
                        reinterpret_cast<uchar4*>(normal), depth, prob, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

static int stream_wait_flag(cudaStream_t stream, const uint32_t* flag, uint32_t value) {
  typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static wait32_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (wait32_fn)p;
  }();
  if (fn && fn((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ) == CUDA_SUCCESS) return RSF_OK;
  launch_shard_wait(flag, 1, value, stream);   // polling kernel
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

// *flag = value, ordered after everything already on `stream`; a stream memory operation (no SM), else a 32-bit memset
static int stream_write_flag(cudaStream_t stream, uint32_t* flag, uint32_t value) {
  typedef CUresult (*write32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  typedef CUresult (*memset32_fn)(CUdeviceptr, unsigned int, size_t, CUstream);
  static write32_fn wfn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (write32_fn)p;
  }();
  static memset32_fn mfn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemsetD32Async", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (memset32_fn)p;
  }();
  if (wfn && wfn((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WRITE_VALUE_DEFAULT) == CUDA_SUCCESS) return RSF_OK;
  if (mfn && mfn((CUdeviceptr)(uintptr_t)flag, value, 1, (CUstream)stream) == CUDA_SUCCESS) return RSF_OK;
  return fail(RSF_ERR_CUDA, "neither cuStreamWriteValue32 nor cuMemsetD32Async could write the frame flag");
}

int rsf_stream_wait_flag_device(rsf_grid_t g, const uint32_t* flag, uint32_t value) {
  if (int rc = check_grid(g)) return rc;
  if (!flag) return fail(RSF_ERR_INVALID, "flag must not be NULL");
  DeviceGuard dg(g->device);
  return stream_wait_flag(g->stream, flag, value);
}

int rsf_stream_wait_flag_on(int device, void* stream, const uint32_t* flag, uint32_t value) {
  if (!flag) return fail(RSF_ERR_INVALID, "flag must not be NULL");
  DeviceGuard dg(device);
  return stream_wait_flag((cudaStream_t)stream, flag, value);
}

int rsf_shard_set_link(rsf_grid_t g, const rsf_shard_link* link) {
  if (int rc = check_grid(g)) return rc;
  if (!link) {
    g->linked = false;
    return RSF_OK;
  }
  if (!link->planes[0] || !link->planes[1] || !link->flags) return fail(RSF_ERR_INVALID, "planes / flags must not be NULL");
  if (link->n_planes < 1 || link->n_planes > 3) return fail(RSF_ERR_INVALID, "1 <= n_planes <= 3");
  if (link->root < 0 || link->root >= g->opt.shard_world) return fail(RSF_ERR_INVALID, "root outside [0, shard_world)");
  if (g->opt.shard_world > 32) return fail(RSF_ERR_INVALID, "rsf_shard_frame_device supports shard_world <= 32");
  if (link->H <= 0 || link->W <= 0) return fail(RSF_ERR_INVALID, "H, W must be > 0");
  if (link->skip_mask & ~6u) return fail(RSF_ERR_INVALID, "skip_mask may only name planes 1 and 2");
  if (link->rx[0]) {
    if (!link->rx[1] || !link->frame_flag || link->frame_H <= 0 || link->frame_W <= 0 || link->n_kids < 0 || link->n_kids > 2)
      return fail(RSF_ERR_INVALID, "frame transport: rx[2], frame_flag, frame_H / frame_W and 0 <= n_kids <= 2 are required");
    if ((size_t)link->frame_H * (size_t)link->frame_W > g->max_pixels) return fail(RSF_ERR_INVALID, "frame size exceeds max_pixels");
    if (((size_t)link->frame_H * link->frame_W * 3) % 4) return fail(RSF_ERR_INVALID, "packed frames need 3 H W divisible by 4");
    for (int c = 0; c < link->n_kids; ++c)
      if (!link->kid_rx[c][0] || !link->kid_rx[c][1] || !link->kid_flag[c]) return fail(RSF_ERR_INVALID, "kid_rx / kid_flag missing");
    DeviceGuard dg(g->device);
    for (int c = 0; c < 2; ++c) {
      if (!g->kid_stream[c]) RSF_CUDA(cudaStreamCreateWithFlags(&g->kid_stream[c], cudaStreamNonBlocking));
      if (!g->ev_step_done[c]) RSF_CUDA(cudaEventCreateWithFlags(&g->ev_step_done[c], cudaEventDisableTiming));
    }
  }
  g->link = *link;
  g->linked = true;
  g->shard_seq = 0;
  g->fed = g->stepped = 0;
  return RSF_OK;
}

int rsf_shard_feed_device(rsf_grid_t g, const uint8_t* packed_frame) {
  if (int rc = check_grid(g)) return rc;
  if (!g->linked || !g->link.rx[0]) return fail(RSF_ERR_INVALID, "rsf_shard_set_link with a frame transport has not been called");
  if (g->fed - g->stepped >= 2) return fail(RSF_ERR_INVALID, "two frames are already in flight: call rsf_shard_step_device first");
  const rsf_shard_link& L = g->link;
  const bool ingest = g->opt.shard_rank == L.root;
  if (ingest && !packed_frame) return fail(RSF_ERR_INVALID, "packed_frame must not be NULL on the ingest rank");
  DeviceGuard dg(g->device);
  const unsigned seq = g->fed + 1;
  const int j = (int)(g->fed & 1u);
  const size_t nbytes = (size_t)15 * L.frame_H * L.frame_W;
  const uint8_t* src = ingest ? packed_frame : L.rx[j];
  g->fed_src[j] = src;
  for (int c = 0; c < L.n_kids; ++c) {
    cudaStream_t st = g->kid_stream[c];
    if (ingest) {
      // frame seq - 2 is through on every rank (its arrival flags were awaited by my step): buffers j are free
      if (g->fed >= 2) RSF_CUDA(cudaStreamWaitEvent(st, g->ev_step_done[j], 0));
    } else {
      if (int rc = stream_wait_flag(st, L.frame_flag + j, seq)) return rc;      // my own copy has landed
    }
    RSF_CUDA(cudaMemcpyAsync(L.kid_rx[c][j], src, nbytes, cudaMemcpyDeviceToDevice, st));   // data first ...
    if (int rc = stream_write_flag(st, L.kid_flag[c] + j, seq)) return rc;                  // ... then its sequence number
  }
  ++g->fed;
  return RSF_OK;
}

int rsf_shard_step_device(rsf_grid_t g, float max_depth, const float K[4], const float pose[7], float raycast_max_depth,
                          const float rc_K[4], const float rc_pose[7], uint8_t* rgba, uint8_t* normal, float* depth_out,
                          float* prob_out) {
  if (int rc = check_grid(g)) return rc;
  if (!g->linked || !g->link.rx[0]) return fail(RSF_ERR_INVALID, "rsf_shard_set_link with a frame transport has not been called");
  if (g->stepped >= g->fed) return fail(RSF_ERR_INVALID, "rsf_shard_step_device without a fed frame");
  const rsf_shard_link& L = g->link;
  const int j = (int)(g->stepped & 1u);
  const unsigned seq = g->stepped + 1;
  const uint8_t* src = g->fed_src[j];
  {
    DeviceGuard dg(g->device);
    if (g->opt.shard_rank != L.root)
      if (int rc = stream_wait_flag(g->stream, L.frame_flag + j, seq)) return rc;           // the frame is in rx[j]
  }
  const size_t P = (size_t)L.frame_H * L.frame_W;
  const int rc = rsf_shard_frame_device(g, src, reinterpret_cast<const float*>(src + 3 * P), reinterpret_cast<const float*>(src + 7 * P),
                                        reinterpret_cast<const float*>(src + 11 * P), L.frame_H, L.frame_W, max_depth, K, pose,
                                        raycast_max_depth, L.H, L.W, rc_K, rc_pose, rgba, normal, depth_out, prob_out);
  if (rc) return rc;
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaEventRecord(g->ev_step_done[j], g->stream));
  ++g->stepped;
  return RSF_OK;
}

int rsf_shard_frame_device(rsf_grid_t g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H, int W,
                           float max_depth, const float K[4], const float pose[7], float raycast_max_depth, int rc_H, int rc_W,
                           const float rc_K[4], const float rc_pose[7], uint8_t* rgba, uint8_t* normal, float* depth_out,
                           float* prob_out) {
  if (int rc = check_grid(g)) return rc;
  if (!rgb || !depth || !ht || !lt) return fail(RSF_ERR_INVALID, "rgb / depth / ht / lt must not be NULL");
  Frame f, rf;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  if (int rc = make_frame(g, rc_H, rc_W, raycast_max_depth, rc_K, rc_pose, &rf)) return rc;
  const int world = g->opt.shard_world, rank = g->opt.shard_rank;
  if (world > 1) {
    if (!g->linked) return fail(RSF_ERR_INVALID, "rsf_shard_set_link has not been called");
    if (rc_H != g->link.H || rc_W != g->link.W) return fail(RSF_ERR_INVALID, "raycast size differs from the linked planes");
  }
  DeviceGuard dg(g->device);
  if (int rc = enqueue_integrate(g, f, rgb, depth, ht, lt, nullptr, false, 1.f, 1.f, nullptr, &rf)) return rc;
  RSF_CUDA(cudaStreamWaitEvent(g->stream, g->ev_prepass, 0));      // tile ranges of `rf` (side stream, overlapped with integrate)
  const int patch_rows = (rc_H + 3) / 4;
  if (world <= 1) {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast_band(g->view, rf, g->opt.truncation / 2, g->d_ranges, reinterpret_cast<uchar4*>(rgba), reinterpret_cast<uchar4*>(normal),
                        depth_out, prob_out, g->sm_count, g->stream, 0, patch_rows, 0);
    RSF_CUDA(cudaGetLastError());
    return RSF_OK;
  }
  const unsigned seq = ++g->shard_seq;
  unsigned long long* planes = reinterpret_cast<unsigned long long*>(g->link.planes[seq & 1u]);
  const int plane_spec = g->link.n_planes | (int)(g->link.skip_mask << 8);
  {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast_band(g->view, rf, g->opt.truncation / 2, g->d_ranges, nullptr, nullptr, nullptr, nullptr, g->sm_count, g->stream,
                        0, patch_rows, 0, planes, plane_spec);
  }
  launch_shard_signal(g->link.flags + rank, seq, g->stream);
  if (rank == g->link.root) {
    ProfScope p(g, RSF_PROF_MERGE);
    launch_shard_wait(g->link.flags, world, seq, g->stream);
    launch_merge_finalize(planes, plane_spec, rc_H * rc_W, reinterpret_cast<uchar4*>(rgba), reinterpret_cast<uchar4*>(normal),
                          depth_out, prob_out, g->sm_count, g->stream);
  }
  RSF_CUDA(cudaGetLastError());
  return RSF_OK;
}

int rsf_raycast_async(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7], uint8_t* rgba,
                      uint8_t* normal, float* depth, float* prob, int* ticket) {
  if (int rc = check_grid(g)) return rc;
  Frame f;
  if (int rc = make_frame(g, H, W, max_depth, K, pose, &f)) return rc;
  DeviceGuard dg(g->device);
  const int slot = (int)(g->out_next++ % rsf_grid_s::kSlots);
  auto& sl = g->out[slot];
  RSF_CUDA(cudaStreamWaitEvent(g->stream, sl.done, 0));       // read-back of the previous user has finished
  // (Rendering in two row bands so that the read-back of the first overlaps the march of the second was measured and
  // dropped: the kernel is bound by its longest patch, two launches have two tails, and the five extra API calls per frame
  // cost the pipelined path 15 % -- 5 950 instead of 6 950 frames/s -- without helping the blocking one.)
  const size_t P = (size_t)H * W;
  {
    ProfScope p(g, RSF_PROF_RAYCAST);
    launch_raycast(g->view, f, g->opt.truncation / 2, g->d_ranges, rgba ? sl.rgba : nullptr, normal ? sl.normal : nullptr,
                   depth ? sl.depth : nullptr, prob ? sl.prob : nullptr, g->sm_count, g->stream);   // step_size = truncation / 2, voxel_tsdf.cu:892
  }
  RSF_CUDA(cudaGetLastError());
  RSF_CUDA(cudaEventRecord(sl.ready, g->stream));
  RSF_CUDA(cudaStreamWaitEvent(g->s_d2h, sl.ready, 0));
  if (rgba) RSF_CUDA(cudaMemcpyAsync(rgba, sl.rgba, P * 4, cudaMemcpyDeviceToHost, g->s_d2h));
  if (normal) RSF_CUDA(cudaMemcpyAsync(normal, sl.normal, P * 4, cudaMemcpyDeviceToHost, g->s_d2h));
  if (depth) RSF_CUDA(cudaMemcpyAsync(depth, sl.depth, P * sizeof(float), cudaMemcpyDeviceToHost, g->s_d2h));
  if (prob) RSF_CUDA(cudaMemcpyAsync(prob, sl.prob, P * sizeof(float), cudaMemcpyDeviceToHost, g->s_d2h));
  RSF_CUDA(cudaEventRecord(sl.done, g->s_d2h));
  if (ticket) *ticket = slot;
  return RSF_OK;
}

int rsf_wait(rsf_grid_t g, int ticket) {
  if (int rc = check_grid(g)) return rc;
  if (ticket < 0 || ticket >= rsf_grid_s::kSlots) return fail(RSF_ERR_INVALID, "bad ticket %d", ticket);
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaEventSynchronize(g->out[ticket].done));
  return RSF_OK;
}

int rsf_raycast(rsf_grid_t g, float max_depth, int H, int W, const float K[4], const float pose[7], uint8_t* rgba,
                uint8_t* normal, float* depth, float* prob) {
  int ticket = 0;
  if (int rc = rsf_raycast_async(g, max_depth, H, W, K, pose, rgba, normal, depth, prob, &ticket)) return rc;
  return rsf_wait(g, ticket);
}

int rsf_synchronize(rsf_grid_t g) {
  if (int rc = check_grid(g)) return rc;
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaStreamSynchronize(g->s_h2d));
  if (int rc = sync_counters(g)) return rc;
  RSF_CUDA(cudaStreamSynchronize(g->s_d2h));
  return RSF_OK;
}

int rsf_num_active_blocks(rsf_grid_t g, int64_t* n) {
  if (int rc = check_grid(g)) return rc;
  if (!n) return fail(RSF_ERR_INVALID, "n must not be NULL");
  DeviceGuard dg(g->device);
  if (int rc = sync_counters(g, false)) return rc;
  *n = g->last.active_blocks;
  return RSF_OK;
}

int rsf_get_stats(rsf_grid_t g, rsf_frame_stats* s) {
  if (int rc = check_grid(g)) return rc;
  if (!s) return fail(RSF_ERR_INVALID, "stats must not be NULL");
  DeviceGuard dg(g->device);
  if (int rc = sync_counters(g, false)) return rc;
  *s = g->last;
  return RSF_OK;
}

int rsf_gather_valid(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows) {
  return gather_rows(g, 0, nullptr, 4, rows, cap_rows, n_rows);
}
int rsf_gather_valid_semantic(rsf_grid_t g, float* rows, int64_t cap_rows, int64_t* n_rows) {
  return gather_rows(g, 0, nullptr, 5, rows, cap_rows, n_rows);
}
int rsf_gather_voxels(rsf_grid_t g, const float bounds[6], float* rows, int64_t cap_rows, int64_t* n_rows) {
  return gather_rows(g, 1, bounds, 4, rows, cap_rows, n_rows);
}

int rsf_gather_valid_mesh(rsf_grid_t g, float* vertices, int64_t cap_vertices, int64_t* n_vertices, int32_t* triangles,
                          int64_t cap_triangles, int64_t* n_triangles, float* vertex_prob) {
  if (int rc = check_grid(g)) return rc;
  if (!n_vertices || !n_triangles) return fail(RSF_ERR_INVALID, "n_vertices / n_triangles must not be NULL");
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(g->view.ctr) + kCtrGatherOff, 0, sizeof(int), g->stream));
  launch_gather_select(g->view, 0, g->opt.shard_world > 1, S3{0, 0, 0}, S3{0, 0, 0}, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (int rc = sync_counters(g, false)) return rc;
  const int nblk = g->h_ctr->gather_count;
  DevBuf<int2> counts(g->stream), offsets(g->stream);
  DevBuf<long long> totals(g->stream);
  RSF_CUDA(counts.alloc(nblk));
  RSF_CUDA(offsets.alloc(nblk));
  RSF_CUDA(totals.alloc(2));
  launch_mesh_count(g->view, nblk, counts.p, offsets.p, totals.p, g->stream);
  RSF_CUDA(cudaGetLastError());
  long long h_tot[2] = {0, 0};
  RSF_CUDA(cudaMemcpyAsync(h_tot, totals.p, sizeof(h_tot), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  *n_vertices = h_tot[0];
  *n_triangles = h_tot[1];
  if (h_tot[0] > 0x7fffffffLL || h_tot[1] > 0x7fffffffLL) return fail(RSF_ERR_CAPACITY, "mesh exceeds 2^31 elements");
  if (!vertices && !triangles && !vertex_prob) return RSF_OK;        // size query
  if (!vertices || !triangles || !vertex_prob)
    return fail(RSF_ERR_INVALID, "vertices, triangles and vertex_prob must be given together");
  if (cap_vertices < h_tot[0] || cap_triangles < h_tot[1])
    return fail(RSF_ERR_CAPACITY, "need %lld vertices / %lld triangles, capacity %lld / %lld", h_tot[0], h_tot[1],
                (long long)cap_vertices, (long long)cap_triangles);
  if (h_tot[0] == 0 && h_tot[1] == 0) return RSF_OK;
  DevBuf<float> dv(g->stream), dp(g->stream);
  DevBuf<int> dt(g->stream);
  RSF_CUDA(dv.alloc((size_t)h_tot[0] * 3));
  RSF_CUDA(dp.alloc((size_t)h_tot[0]));
  RSF_CUDA(dt.alloc((size_t)h_tot[1] * 3));
  launch_mesh_emit(g->view, g->opt.voxel_size, nblk, offsets.p, dv.p, dp.p, dt.p, g->stream);
  RSF_CUDA(cudaGetLastError());
  RSF_CUDA(cudaMemcpyAsync(vertices, dv.p, (size_t)h_tot[0] * 3 * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaMemcpyAsync(vertex_prob, dp.p, (size_t)h_tot[0] * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaMemcpyAsync(triangles, dt.p, (size_t)h_tot[1] * 3 * sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  return RSF_OK;
}

int rsf_export_blocks(rsf_grid_t g, int owned_only, int16_t* coords, float* tsdf, uint8_t* rgbw, float* prob,
                      int64_t cap_blocks, int64_t* n_blocks) {
  if (int rc = check_grid(g)) return rc;
  if (!n_blocks) return fail(RSF_ERR_INVALID, "n_blocks must not be NULL");
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(g->view.ctr) + kCtrGatherOff, 0, sizeof(int), g->stream));
  launch_gather_select(g->view, 0, owned_only, S3{0, 0, 0}, S3{0, 0, 0}, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (int rc = sync_counters(g, false)) return rc;
  const int nblk = g->h_ctr->gather_count;
  *n_blocks = nblk;
  if (!coords && !tsdf && !rgbw && !prob) return RSF_OK;
  if (cap_blocks < nblk) return fail(RSF_ERR_CAPACITY, "need %d blocks, capacity %lld", nblk, (long long)cap_blocks);
  if (nblk == 0) return RSF_OK;
  DevBuf<short> dc(g->stream);
  DevBuf<float> dt(g->stream), dp(g->stream);
  DevBuf<uint32_t> dr(g->stream);
  const size_t nv = (size_t)nblk * kBlockVol;
  if (coords) RSF_CUDA(dc.alloc((size_t)nblk * 3));
  if (tsdf) RSF_CUDA(dt.alloc(nv));
  if (rgbw) RSF_CUDA(dr.alloc(nv));
  if (prob) RSF_CUDA(dp.alloc(nv));
  launch_export_blocks(g->view, nblk, dc.p, dt.p, dr.p, dp.p, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (coords) RSF_CUDA(cudaMemcpyAsync(coords, dc.p, (size_t)nblk * 3 * sizeof(short), cudaMemcpyDeviceToHost, g->stream));
  if (tsdf) RSF_CUDA(cudaMemcpyAsync(tsdf, dt.p, nv * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  if (rgbw) RSF_CUDA(cudaMemcpyAsync(rgbw, dr.p, nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
  if (prob) RSF_CUDA(cudaMemcpyAsync(prob, dp.p, nv * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  return RSF_OK;
}

uint32_t rsf_hash(int16_t x, int16_t y, int16_t z, uint32_t mask) { return hash_block(S3{x, y, z}, mask); }
uint32_t rsf_shard_owner(int16_t x, int16_t y, int16_t z, int shift, int world) {
  if (world <= 1) return 0;
  return shard_owner(S3{x, y, z}, shift, world);
}

int rsf_allocate_blocks(rsf_grid_t g, const int16_t* block_coords, int n) {
  if (int rc = check_grid(g)) return rc;
  if (n < 0 || (n > 0 && !block_coords)) return fail(RSF_ERR_INVALID, "bad block_coords / n");
  DeviceGuard dg(g->device);
  if (int rc = maybe_rebuild_table(g)) return rc;
  DevBuf<short> dc(g->stream);
  RSF_CUDA(dc.alloc((size_t)n * 3));
  RSF_CUDA(cudaMemcpyAsync(dc.p, block_coords, (size_t)n * 3 * sizeof(short), cudaMemcpyHostToDevice, g->stream));
  launch_allocate_coords(g->view, dc.p, n, g->sm_count, g->stream);
  RSF_CUDA(cudaGetLastError());
  return sync_counters(g);
}

int rsf_rebuild_table(rsf_grid_t g) {
  if (int rc = check_grid(g)) return rc;
  DeviceGuard dg(g->device);
  if (int rc = sync_counters(g)) return rc;      // known_alive must be exact: it becomes known_keys
  if (int rc = rebuild_table_now(g)) return rc;
  return sync_counters(g);
}

int rsf_set_table_policy(rsf_grid_t g, int64_t max_ghosts) {
  if (int rc = check_grid(g)) return rc;
  if (max_ghosts < 0) return fail(RSF_ERR_INVALID, "max_ghosts must be >= 0 (0 = default)");
  g->max_ghosts = max_ghosts;
  return RSF_OK;
}

int rsf_delete_blocks(rsf_grid_t g, const int16_t* block_coords, int n) {
  if (int rc = check_grid(g)) return rc;
  if (n < 0 || (n > 0 && !block_coords)) return fail(RSF_ERR_INVALID, "bad block_coords / n");
  DeviceGuard dg(g->device);
  DevBuf<short> dc(g->stream);
  RSF_CUDA(dc.alloc((size_t)n * 3));
  RSF_CUDA(cudaMemcpyAsync(dc.p, block_coords, (size_t)n * 3 * sizeof(short), cudaMemcpyHostToDevice, g->stream));
  launch_delete_coords(g->view, dc.p, n, g->stream);
  RSF_CUDA(cudaGetLastError());
  return sync_counters(g);
}

int rsf_retrieve_voxels(rsf_grid_t g, const int16_t* voxel_coords, int n, float* tsdf, uint8_t* rgbw, float* prob,
                        uint8_t* found) {
  if (int rc = check_grid(g)) return rc;
  if (n < 0 || (n > 0 && !voxel_coords)) return fail(RSF_ERR_INVALID, "bad voxel_coords / n");
  DeviceGuard dg(g->device);
  DevBuf<short> dc(g->stream);
  DevBuf<float> dt(g->stream), dp(g->stream);
  DevBuf<uint32_t> dr(g->stream);
  DevBuf<uint8_t> df(g->stream);
  RSF_CUDA(dc.alloc((size_t)n * 3));
  RSF_CUDA(dt.alloc(n));
  RSF_CUDA(dp.alloc(n));
  RSF_CUDA(dr.alloc(n));
  RSF_CUDA(df.alloc(n));
  RSF_CUDA(cudaMemcpyAsync(dc.p, voxel_coords, (size_t)n * 3 * sizeof(short), cudaMemcpyHostToDevice, g->stream));
  launch_retrieve_voxels(g->view, dc.p, n, dt.p, dr.p, dp.p, df.p, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (tsdf) RSF_CUDA(cudaMemcpyAsync(tsdf, dt.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  if (rgbw) RSF_CUDA(cudaMemcpyAsync(rgbw, dr.p, (size_t)n * 4, cudaMemcpyDeviceToHost, g->stream));
  if (prob) RSF_CUDA(cudaMemcpyAsync(prob, dp.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  if (found) RSF_CUDA(cudaMemcpyAsync(found, df.p, (size_t)n, cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  return RSF_OK;
}

int rsf_assign_voxels(rsf_grid_t g, const int16_t* voxel_coords, int n, const float* tsdf, const uint8_t* rgbw,
                      const float* prob, uint8_t* written) {
  if (int rc = check_grid(g)) return rc;
  if (n < 0 || (n > 0 && !voxel_coords)) return fail(RSF_ERR_INVALID, "bad voxel_coords / n");
  DeviceGuard dg(g->device);
  DevBuf<short> dc(g->stream);
  DevBuf<float> dt(g->stream), dp(g->stream);
  DevBuf<uint32_t> dr(g->stream);
  DevBuf<uint8_t> dw(g->stream);
  RSF_CUDA(dc.alloc((size_t)n * 3));
  RSF_CUDA(dw.alloc(n));
  RSF_CUDA(cudaMemcpyAsync(dc.p, voxel_coords, (size_t)n * 3 * sizeof(short), cudaMemcpyHostToDevice, g->stream));
  if (tsdf) {
    RSF_CUDA(dt.alloc(n));
    RSF_CUDA(cudaMemcpyAsync(dt.p, tsdf, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, g->stream));
  }
  if (rgbw) {
    RSF_CUDA(dr.alloc(n));
    RSF_CUDA(cudaMemcpyAsync(dr.p, rgbw, (size_t)n * 4, cudaMemcpyHostToDevice, g->stream));
  }
  if (prob) {
    RSF_CUDA(dp.alloc(n));
    RSF_CUDA(cudaMemcpyAsync(dp.p, prob, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, g->stream));
  }
  launch_assign_voxels(g->view, dc.p, n, dt.p, dr.p, dp.p, dw.p, g->stream);
  RSF_CUDA(cudaGetLastError());
  if (written) RSF_CUDA(cudaMemcpyAsync(written, dw.p, (size_t)n, cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  return RSF_OK;
}

int rsf_retrieve_tsdf_trilinear(rsf_grid_t g, const float* points, int n, float* tsdf) {
  if (int rc = check_grid(g)) return rc;
  if (n < 0 || (n > 0 && (!points || !tsdf))) return fail(RSF_ERR_INVALID, "bad points / tsdf / n");
  DeviceGuard dg(g->device);
  DevBuf<float> dpnt(g->stream), dout(g->stream);
  RSF_CUDA(dpnt.alloc((size_t)n * 3));
  RSF_CUDA(dout.alloc(n));
  RSF_CUDA(cudaMemcpyAsync(dpnt.p, points, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, g->stream));
  launch_trilinear(g->view, dpnt.p, n, dout.p, g->stream);
  RSF_CUDA(cudaGetLastError());
  RSF_CUDA(cudaMemcpyAsync(tsdf, dout.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  return RSF_OK;
}

int rsf_profile_enable(rsf_grid_t g, int on) {
  if (int rc = check_grid(g)) return rc;
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  g->profiling = on != 0;
  for (auto& u : g->prof_used) u = 0;
  return RSF_OK;
}

int rsf_profile_read(rsf_grid_t g, double ms[RSF_PROF_NUM], int64_t launches[RSF_PROF_NUM]) {
  if (int rc = check_grid(g)) return rc;
  if (!ms || !launches) return fail(RSF_ERR_INVALID, "ms / launches must not be NULL");
  DeviceGuard dg(g->device);
  RSF_CUDA(cudaStreamSynchronize(g->stream));
  for (int k = 0; k < RSF_PROF_NUM; ++k) {
    double sum = 0;
    for (size_t i = 0; i < g->prof_used[k]; ++i) {
      float t = 0.f;
      RSF_CUDA(cudaEventElapsedTime(&t, g->prof_events[k][i].first, g->prof_events[k][i].second));
      sum += t;
    }
    ms[k] = sum;
    launches[k] = (int64_t)g->prof_used[k];
    g->prof_used[k] = 0;
  }
  return RSF_OK;
}

int rsf_get_totals(rsf_grid_t g, int64_t totals[5]) {
  if (int rc = check_grid(g)) return rc;
  if (!totals) return fail(RSF_ERR_INVALID, "totals must not be NULL");
  DeviceGuard dg(g->device);
  if (int rc = sync_counters(g, false)) return rc;
  const Counters& c = *g->h_ctr;   // totals exclude the frame still held in the per-frame counters
  totals[0] = (int64_t)c.tot_frames;
  totals[1] = (int64_t)(c.tot_new_blocks + c.new_blocks);
  totals[2] = (int64_t)(c.tot_visible_blocks + (unsigned long long)c.vis_count);
  totals[3] = (int64_t)(c.tot_updated_voxels + c.updated_voxels);
  totals[4] = (int64_t)(c.tot_carved_blocks + c.carved_blocks);
  return RSF_OK;
}

int rsf_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(RSF_ERR_INVALID, "ptr must not be NULL");
  RSF_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return RSF_OK;
}
int rsf_host_free(void* ptr) {
  if (ptr) RSF_CUDA(cudaFreeHost(ptr));
  return RSF_OK;
}

}  // extern "C"

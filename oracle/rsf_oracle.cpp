// rsf_oracle.cpp -- CPU restatement of RA-SLAM's voxel-hashing TSDF + semantic fusion path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ra-slam_b200/, include/) may call,
// link or import this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker / the timed CPU baseline.
//
// Parity status: the hash function, miss defaults, fresh-block defaults and active-block counts
// are pinned by the reference's own unit tests (utils/tests/voxel_hash_test.cu:130-155,
// voxel_mem_test.cu:87; see tests/golden/).  Integrate / RayCast / Gather* have no reference
// tests, so they are pinned by OUTPUT OF THE REFERENCE ITSELF: its unmodified utils/tsdf/*.cu are
// compiled for sm_100a against stand-in Eigen / OpenCV / GL headers (oracle/Makefile target _ref,
// oracle/ref_shim/), run on a B200, and tests/golden/reference_small.npz holds what they produced;
// tests/test_reference_golden.py checks this file against it on CPU (identical block sets and
// labels, |dTSDF| <= 1e-4 apart from pixel-rounding-tie voxels).  tests/test_gpu_vs_reference.py
// does the same for the product against the live reference library at full resolution.
//
// Arithmetic contract: every float operation is rounded individually (build with
// -ffp-contract=off, no -ffast-math), in the order the reference source writes it.  Eigen 3.3.7
// (not vendored under /root/reference) semantics that matter are restated from its published
// headers: q*v = v + w*uv + q.vec x uv with uv = 2(q.vec x v) (Geometry/Quaternion.h
// _transformVector); q.inverse() = conj(q)/|q|^2; hnormalized() = (x/z, y/z);
// norm() = sqrt(x^2+(y^2+z^2)) (Redux.h splits the range in halves: c0 + (c1 + c2)); normalized() = v / norm (when norm^2 > 0).
//
// Allocation semantics: the reference's bucket try-lock insert is lossy and non-deterministic
// (utils/tsdf/voxel_hash.cu:71-76,93-105).  The oracle implements the *ideal* semantics
// (SURVEY.md D5/H1): every requested block is inserted in the frame that requests it, every
// carve succeeds.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../ra-slam_b200/csrc/rsf_mc_table.h"   // packed classic marching-cubes triangle table (data only)
#include <memory>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

constexpr int kBlockLen = 8;        // voxel_mem.cuh:17-22
constexpr int kBlockVol = 512;

struct V3 {
  float x, y, z;
};
struct Quat {
  float w, x, y, z;
};

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }   // Redux.h redux_novec_unroller: c0 + (c1 + c2)
inline float norm(V3 a) { return std::sqrt(a.x * a.x + (a.y * a.y + a.z * a.z)); }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3.7 QuaternionBase::_transformVector
inline V3 qrot(Quat q, V3 v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  const V3 c = cross(qv, uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
// Eigen 3.3.7 QuaternionBase::inverse
inline Quat qinv(Quat q) {
  const float n2 = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);   // coeffs x,y,z,w reduced in halves
  if (n2 > 0.f) return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return {0.f, 0.f, 0.f, 0.f};
}

// utils/cuda/lie_group.cuh:8-45
struct SE3 {
  Quat q;
  V3 t;
  V3 apply(V3 v) const { return qrot(q, v) + t; }      // lie_group.cuh:33-36
  SE3 inverse() const {                                // lie_group.cuh:25-27
    const Quat qi = qinv(q);
    return {qi, qrot(qi, neg(t))};
  }
};

// utils/cuda/camera.cuh:13-52
struct Intr {
  float fx, fy, cx, cy;
  Intr inverse() const {                               // camera.cuh:35-39
    const float fxi = 1.f / fx, fyi = 1.f / fy;
    return {fxi, fyi, -cx * fxi, -cy * fyi};
  }
  V3 mul(V3 v) const { return {fx * v.x + cx * v.z, fy * v.y + cy * v.z, v.z}; }  // camera.cuh:48-51
};

// float -> int as the GPU does it (cvt.rzi.s32.f32): NaN -> 0, saturating
inline int cvt_int(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.f) return 2147483647;
  if (v <= -2147483648.f) return (-2147483647 - 1);
  return (int)v;
}
inline int16_t rnd_s(float v) { return (int16_t)(int)std::round(v); }   // roundf + cast<short>

struct S3 {
  int16_t x, y, z;
};
inline S3 rnd3(V3 p) { return {rnd_s(p.x), rnd_s(p.y), rnd_s(p.z)}; }
inline S3 blk(S3 g) {                                  // voxel_mem.cuh:31-34 (arithmetic >>)
  return {(int16_t)(g.x >> 3), (int16_t)(g.y >> 3), (int16_t)(g.z >> 3)};
}
inline int off_idx(S3 g) {                             // voxel_mem.cuh:55-70
  return (g.x & 7) + (g.y & 7) * 8 + (g.z & 7) * 64;
}
inline uint64_t pack(S3 b) {
  return (uint64_t)(uint16_t)b.x | ((uint64_t)(uint16_t)b.y << 16) | ((uint64_t)(uint16_t)b.z << 32);
}
inline S3 unpack(uint64_t k) {
  return {(int16_t)(k & 0xffff), (int16_t)((k >> 16) & 0xffff), (int16_t)((k >> 32) & 0xffff)};
}

struct Block {
  S3 pos;
  float tsdf[kBlockVol];
  uint8_t rgbw[kBlockVol * 4];   // r,g,b,weight  (voxel_types.cuh:10-19)
  float prob[kBlockVol];
  std::vector<float> cls;        // K-class extension: normalised probabilities [k][voxel], empty when K <= 2
  void reset() {                                       // voxel_mem.cu:44-51 (rgb: undefined -> 0)
    for (int i = 0; i < kBlockVol; ++i) {
      tsdf[i] = -1.f;
      prob[i] = .5f;
      rgbw[4 * i + 0] = rgbw[4 * i + 1] = rgbw[4 * i + 2] = 0;
      rgbw[4 * i + 3] = 1;
    }
  }
};

struct Stats {
  int64_t new_blocks, visible_blocks, updated_voxels, carved_blocks;
};

// Spatial-shard ownership (new in this repo); mirrors
// ra-slam_b200/csrc/rsf_device.cuh rsf_shard_owner bit for bit.
inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
inline uint32_t shard_owner(S3 b, int shift, int world) {
  const int x = b.x >> shift, y = b.y >> shift, z = b.z >> shift;
  uint32_t h = mix32((uint32_t)x * 0x9e3779b1u ^ mix32((uint32_t)y * 0x85ebca77u ^ mix32((uint32_t)z * 0xc2b2ae3du)));
  return h % (uint32_t)world;
}

struct Grid {
  float vs, tr;
  int threads;
  std::unordered_map<uint64_t, std::unique_ptr<Block>> map;
  Stats stats{};
  // sharding
  int shard_rank = 0, shard_world = 1, shard_shift = 4, shard_halo = 1;
  // K-class semantic extension: per-class probability planes of the current frame
  int num_classes = 0;
  const float* cls_in = nullptr;     // [K][H*W]

  bool wanted(S3 b) const {
    if (shard_world <= 1) return true;
    // owned, or within `halo` blocks of an owned super-brick
    for (int dz = -shard_halo; dz <= shard_halo; ++dz)
      for (int dy = -shard_halo; dy <= shard_halo; ++dy)
        for (int dx = -shard_halo; dx <= shard_halo; ++dx) {
          S3 n{(int16_t)(b.x + dx), (int16_t)(b.y + dy), (int16_t)(b.z + dz)};
          if ((int)shard_owner(n, shard_shift, shard_world) == shard_rank) return true;
        }
    return false;
  }
  bool owned(S3 b) const {
    return shard_world <= 1 || (int)shard_owner(b, shard_shift, shard_world) == shard_rank;
  }

  const Block* find(S3 b) const {
    auto it = map.find(pack(b));
    return it == map.end() ? nullptr : it->second.get();
  }
};

template <class F>
void parallel_for(int n, int threads, F f) {
  if (threads <= 1 || n < 2) {
    for (int i = 0; i < n; ++i) f(i, 0);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  const int chunk = std::max(1, n / (threads * 8));
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t]() {
      for (;;) {
        const int s = next.fetch_add(chunk);
        if (s >= n) break;
        const int e = std::min(n, s + chunk);
        for (int i = s; i < e; ++i) f(i, t);
      }
    });
  for (auto& th : pool) th.join();
}

// voxel_tsdf.cu:64-73
inline bool voxel_visible(S3 g, const SE3& T_cw, const Intr& K, int W, int H, float vs) {
  const V3 pw{(float)g.x * vs, (float)g.y * vs, (float)g.z * vs};
  const V3 pc = T_cw.apply(pw);
  const V3 h = K.mul(pc);
  const float u = h.x / h.z, v = h.y / h.z;
  return (u >= 0 && u <= (float)(W - 1) && v >= 0 && v <= (float)(H - 1) && h.z >= 0);
}
// voxel_tsdf.cu:75-96
template <bool Full>
inline bool block_visible(S3 b, const SE3& T_cw, const Intr& K, int W, int H, float vs) {
  const int16_t x = (int16_t)(b.x << 3), y = (int16_t)(b.y << 3), z = (int16_t)(b.z << 3);
  bool vis = Full;
  for (int i = 0; i < 8; ++i) {
    const S3 c{(int16_t)(x + ((i >> 0) & 1) * (kBlockLen - 1)), (int16_t)(y + ((i >> 1) & 1) * (kBlockLen - 1)),
               (int16_t)(z + ((i >> 2) & 1) * (kBlockLen - 1))};
    if (Full)
      vis &= voxel_visible(c, T_cw, K, W, H, vs);
    else
      vis |= voxel_visible(c, T_cw, K, W, H, vs);
  }
  return vis;
}

// voxel_tsdf.cu:120-168 (block_allocate_kernel) with ideal insert
void allocate(Grid& g, const float* depth, int H, int W, float max_depth, const Intr& K, const SE3& T_cw) {
  const Intr Ki = K.inverse();
  const SE3 T_wc = T_cw.inverse();
  std::vector<std::vector<uint64_t>> fresh(std::max(1, g.threads));
  parallel_for(H, g.threads, [&](int y, int tid) {
    std::unordered_set<uint64_t> seen;
    auto& out = fresh[tid];
    for (int x = 0; x < W; ++x) {
      const float d = depth[y * W + x];
      const V3 pc = Ki.mul(V3{(float)x, (float)y, 1.f});
      const float r = norm(pc);
      if (d == 0 || d > max_depth) continue;
      const V3 pw = T_wc.apply(pc * d);
      const V3 dir_c = pc / r;
      const V3 dir_w = qrot(T_wc.q, dir_c);
      const V3 start_w = pw - dir_w * g.tr;
      const V3 dir_g = dir_w / g.vs;
      const V3 start_g = start_w / g.vs;
      const V3 ray_g = dir_g * (2.f * g.tr);
      const int n = (int)std::ceil(std::fmax(std::fmax(std::fabs(ray_g.x), std::fabs(ray_g.y)), std::fabs(ray_g.z)) /
                                   (float)kBlockLen);
      const V3 step = ray_g / std::fmax((float)n, 1.f);
      V3 p = start_g;
      for (int i = 0; i <= n; ++i, p = p + step) {
        const S3 b = blk(rnd3(p));
        const uint64_t key = pack(b);
        if (!seen.insert(key).second) continue;
        if (g.map.count(key)) continue;
        if (!g.wanted(b)) continue;
        if (block_visible<true>(b, T_cw, K, W, H, g.vs)) out.push_back(key);
      }
    }
  });
  int64_t n_new = 0;
  for (auto& v : fresh)
    for (uint64_t key : v) {
      auto& slot = g.map[key];
      if (!slot) {
        slot.reset(new Block);
        slot->pos = unpack(key);
        slot->reset();
        if (g.num_classes > 2) slot->cls.assign((size_t)g.num_classes * kBlockVol, 1.f / (float)g.num_classes);
        ++n_new;
      }
    }
  g.stats.new_blocks = n_new;
}

// voxel_tsdf.cu:170-251 (tsdf_integrate_kernel), one block; returns #updated voxels
int integrate_block(Block& B, const Grid& g, const uint8_t* rgb, const float* depth, const float* ht,
                    const float* lt, int H, int W, float max_depth, const Intr& K, const Intr& Ki, const SE3& T_cw) {
  int updated = 0;
  for (int tz = 0; tz < 8; ++tz)
    for (int ty = 0; ty < 8; ++ty)
      for (int tx = 0; tx < 8; ++tx) {
        const S3 ga{(int16_t)((B.pos.x << 3) + tx), (int16_t)((B.pos.y << 3) + ty), (int16_t)((B.pos.z << 3) + tz)};
        const V3 pw{(float)ga.x * g.vs, (float)ga.y * g.vs, (float)ga.z * g.vs};
        const V3 pc = T_cw.apply(pw);
        const V3 h = K.mul(pc);
        const float fu = h.x / h.z, fv = h.y / h.z;
        const int u = cvt_int(std::round(fu)), v = cvt_int(std::round(fv));
        if (!(u >= 0 && u < W && v >= 0 && v < H)) continue;
        const int pi = v * W + u;
        const float d = depth[pi];
        if (d == 0 || d > max_depth) continue;
        // img_depth_to_range written by the allocate kernel (voxel_tsdf.cu:134-140)
        const float d2r = norm(Ki.mul(V3{(float)u, (float)v, 1.f}));
        const float sdf = d2r * (d - h.z);
        if (!(sdf > -g.tr)) continue;
        const float tn = std::fmin(1.f, sdf / g.tr);
        const int i = tx + ty * 8 + tz * 64;
        const float wn = (1.f - d / max_depth) * 4.f;
        const float wo = (float)B.rgbw[4 * i + 3];
        const float w = wo + wn;
        for (int c = 0; c < 3; ++c) {
          const float o = (float)B.rgbw[4 * i + c], nw = (float)rgb[3 * pi + c];
          B.rgbw[4 * i + c] = (uint8_t)(int)std::round((o * wo + nw * wn) / w);
        }
        if (g.cls_in) {
          // K-class form of voxel_tsdf.cu:241-248:
          //   p_k <- exp((w_old ln p_k + w_new ln q_k) / w) / sum_j exp((w_old ln p_j + w_new ln q_j) / w)
          const int Kc = g.num_classes;
          const size_t P = (size_t)H * W;
          if (B.cls.empty()) B.cls.assign((size_t)Kc * kBlockVol, 1.f / (float)Kc);   // block made by a test hook
          float sum = 0.f;
          for (int k = 0; k < Kc; ++k) {
            float& pk = B.cls[(size_t)k * kBlockVol + i];
            pk = std::exp((wo * std::log(pk) + wn * std::log(g.cls_in[(size_t)k * P + pi])) / w);
            sum += pk;
          }
          for (int k = 0; k < Kc; ++k) B.cls[(size_t)k * kBlockVol + i] /= sum;
        }
        B.tsdf[i] = (B.tsdf[i] * wo + tn * wn) / w;
        B.rgbw[4 * i + 3] = (uint8_t)(int)std::fmin(std::round(w), 40.f);
        const float p = B.prob[i];
        const float pos = std::exp((wo * std::log(p) + wn * std::log(ht[pi])) / w);
        const float ng = std::exp((wo * std::log(1.f - p) + wn * std::log(lt[pi])) / w);
        B.prob[i] = pos / (pos + ng);
        ++updated;
      }
  return updated;
}

void integrate(Grid& g, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H, int W,
               float max_depth, const Intr& K, const SE3& T_cw) {
  allocate(g, depth, H, W, max_depth, K, T_cw);
  // voxel_tsdf.cu:98-109,465-474: visible = allocated blocks with ANY corner in view
  std::vector<Block*> all;
  all.reserve(g.map.size());
  for (auto& kv : g.map) all.push_back(kv.second.get());
  std::vector<uint8_t> vis(all.size());
  parallel_for((int)all.size(), g.threads,
               [&](int i, int) { vis[i] = block_visible<false>(all[i]->pos, T_cw, K, W, H, g.vs); });
  std::vector<Block*> visible;
  for (size_t i = 0; i < all.size(); ++i)
    if (vis[i]) visible.push_back(all[i]);
  g.stats.visible_blocks = (int64_t)visible.size();
  const Intr Ki = K.inverse();
  std::vector<int64_t> upd(std::max(1, g.threads), 0);
  std::vector<uint8_t> carve(visible.size());
  parallel_for((int)visible.size(), g.threads, [&](int i, int tid) {
    upd[tid] += integrate_block(*visible[i], g, rgb, depth, ht, lt, H, W, max_depth, K, Ki, T_cw);
    // voxel_tsdf.cu:253-276 (space_carving_kernel, threshold .9 at :880)
    float m = std::fabs(visible[i]->tsdf[0]);
    for (int j = 1; j < kBlockVol; ++j) m = std::fmin(m, std::fabs(visible[i]->tsdf[j]));
    carve[i] = (m >= 0.9f);
  });
  int64_t u = 0, c = 0;
  for (auto x : upd) u += x;
  for (size_t i = 0; i < visible.size(); ++i)
    if (carve[i]) {
      g.map.erase(pack(visible[i]->pos));
      ++c;
    }
  g.stats.updated_voxels = u;
  g.stats.carved_blocks = c;
}

// ---- lookups with the reference's miss defaults (voxel_types.cu:3,8,11; voxel_hash.cuh:104-143)
struct Probe {
  const Grid& g;
  uint64_t ckey = ~0ull;
  const Block* cblk = nullptr;
  explicit Probe(const Grid& gg) : g(gg) {}
  const Block* block_of(S3 p) {
    const uint64_t k = pack(blk(p));
    if (k != ckey) {
      ckey = k;
      cblk = g.find(blk(p));
    }
    return cblk;
  }
  float tsdf(S3 p) {
    const Block* b = block_of(p);
    return b ? b->tsdf[off_idx(p)] : -10.f;
  }
  int weight(S3 p) {
    const Block* b = block_of(p);
    return b ? b->rgbw[4 * off_idx(p) + 3] : 0;
  }
  float prob(S3 p) {
    const Block* b = block_of(p);
    return b ? b->prob[off_idx(p)] : 0.f;
  }
  void rgb(S3 p, uint8_t out[3]) {
    const Block* b = block_of(p);
    for (int c = 0; c < 3; ++c) out[c] = b ? b->rgbw[4 * off_idx(p) + c] : 0;
  }
  // voxel_hash.cu:161-188 (RetrieveTSDF) -- mirrored weights are a reference quirk, kept.
  float tri(V3 pt) {
    const V3 pl{std::floor(pt.x), std::floor(pt.y), std::floor(pt.z)};
    const V3 ph{pl.x + 1.f, pl.y + 1.f, pl.z + 1.f};
    const V3 a = ph - pt;
    float t[8];
    for (int i = 0; i < 8; ++i) {
      const S3 c{(int16_t)(int)(((i >> 2) & 1) ? pl.x : ph.x), (int16_t)(int)(((i >> 1) & 1) ? pl.y : ph.y),
                 (int16_t)(int)(((i >> 0) & 1) ? pl.z : ph.z)};
      t[i] = tsdf(c);
    }
    const float t00 = t[0] * a.z + t[1] * (1 - a.z);
    const float t01 = t[2] * a.z + t[3] * (1 - a.z);
    const float t10 = t[4] * a.z + t[5] * (1 - a.z);
    const float t11 = t[6] * a.z + t[7] * (1 - a.z);
    const float t0 = t00 * a.y + t01 * (1 - a.y);
    const float t1 = t10 * a.y + t11 * (1 - a.y);
    return t0 * a.x + t1 * (1 - a.x);
  }
};

// voxel_tsdf.cu:278-374 (ray_cast_kernel); step_size = truncation/2 (:892).
// depth / prob outputs are new: depth = z of the refined hit point in the camera
// frame, 0 when the ray finds no surface; prob = fused probability of the hit voxel.
void raycast(const Grid& g, float max_depth, int H, int W, const Intr& K, const SE3& T_cw, uint8_t* rgba,
             uint8_t* normal, float* depth_out, float* prob_out) {
  const Intr Ki = K.inverse();
  const SE3 T_wc = T_cw.inverse();
  const float step_size = g.tr / 2;
  const int max_step = (int)std::ceil(max_depth / step_size);
  parallel_for(H, g.threads, [&](int y, int) {
    for (int x = 0; x < W; ++x) {
      const int idx = y * W + x;
      const V3 pc = Ki.mul(V3{(float)x, (float)y, 1.f});
      const float n2 = pc.x * pc.x + (pc.y * pc.y + pc.z * pc.z);
      const V3 dir_c = n2 > 0.f ? pc / std::sqrt(n2) : pc;
      const V3 dir_w = qrot(T_wc.q, dir_c);
      const V3 coarse = (dir_w * step_size) / g.vs;
      V3 sv = coarse;
      V3 p = T_wc.t / g.vs;
      Probe pr(g);
      float tp = pr.tsdf(rnd3(p));
      p = p + sv;
      bool hit = false;
      for (int i = 1; i < max_step; ++i) {
        const S3 gi = rnd3(p);
        const float tc = pr.tsdf(gi);
        const int wc = pr.weight(gi);
        if (wc < 10) {
          p = p + sv;
          tp = tc;
          continue;
        }
        if (tp > 0 && tc <= 0 && tp - tc <= 2.0f) {
          // sharded grids: only the owner of the block holding the crossing sample answers (halo
          // copies are bit-identical but must stay silent so the min-depth merge takes one hit)
          if (!g.owned(blk(gi))) break;
          const V3 p1 = p - sv;
          const float ac = pr.tri(p);
          const float ap = pr.tri(p1);
          const V3 hp = p + sv * (ac / (ap - ac));
          const S3 f = rnd3(hp);
          uint8_t c3[3];
          pr.rgb(f, c3);
          const float pb = pr.prob(f);
          const V3 nr{pr.tsdf(S3{(int16_t)(f.x + 1), f.y, f.z}) - pr.tsdf(S3{(int16_t)(f.x - 1), f.y, f.z}),
                      pr.tsdf(S3{f.x, (int16_t)(f.y + 1), f.z}) - pr.tsdf(S3{f.x, (int16_t)(f.y - 1), f.z}),
                      pr.tsdf(S3{f.x, f.y, (int16_t)(f.z + 1)}) - pr.tsdf(S3{f.x, f.y, (int16_t)(f.z - 1)})};
          const float diff = std::fmax(dot(nr, neg(dir_w)) / norm(nr), 0.f);
          const float a = std::fmax(pb - 0.5f, 0.f) / .5f;
          if (rgba) {
            rgba[4 * idx + 0] = (uint8_t)(a * 255 + (1 - a) * (float)c3[0]);
            rgba[4 * idx + 1] = (uint8_t)((1 - a) * (float)c3[1]);
            rgba[4 * idx + 2] = (uint8_t)((1 - a) * (float)c3[2]);
            rgba[4 * idx + 3] = 255;
          }
          if (normal) {
            normal[4 * idx + 0] = (uint8_t)(a * 255 + (1 - a) * diff * 255);
            normal[4 * idx + 1] = (uint8_t)((1 - a) * diff * 255);
            normal[4 * idx + 2] = (uint8_t)((1 - a) * diff * 255);
            normal[4 * idx + 3] = 255;
          }
          if (depth_out) depth_out[idx] = T_cw.apply(hp * g.vs).z;
          if (prob_out) prob_out[idx] = pb;
          hit = true;
          break;
        }
        tp = tc;
        sv = (tc < 0.5f) ? coarse / 10.f : coarse;
        p = p + sv;
      }
      if (!hit) {
        if (rgba) std::memset(rgba + 4 * idx, 0, 4);
        if (normal) std::memset(normal + 4 * idx, 0, 4);
        if (depth_out) depth_out[idx] = 0.f;
        if (prob_out) prob_out[idx] = 0.f;
      }
    }
  });
}

std::vector<const Block*> sorted_blocks(const Grid& g, bool owned_only = false) {
  std::vector<const Block*> v;
  v.reserve(g.map.size());
  for (auto& kv : g.map)
    if (!owned_only || g.owned(kv.second->pos)) v.push_back(kv.second.get());
  std::sort(v.begin(), v.end(), [](const Block* a, const Block* b) { return pack(a->pos) < pack(b->pos); });
  return v;
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

struct ora_grid {
  Grid g;
};

ora_grid* ora_create(float voxel_size, float truncation, int threads) {
  auto* o = new ora_grid;
  o->g.vs = voxel_size;
  o->g.tr = truncation;
  o->g.threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  return o;
}
void ora_destroy(ora_grid* o) { delete o; }

void ora_set_shard(ora_grid* o, int rank, int world, int shift, int halo) {
  o->g.shard_rank = rank;
  o->g.shard_world = world;
  o->g.shard_shift = shift;
  o->g.shard_halo = halo;
}

// voxel_hash.cu:19-23 (Hash); mask = NUM_BUCKET-1 in the reference
uint32_t ora_hash(int16_t x, int16_t y, int16_t z, uint32_t mask) {
  return (((uint32_t)(int32_t)x * 73856093u) ^ ((uint32_t)(int32_t)y * 19349669u) ^ ((uint32_t)(int32_t)z * 83492791u)) & mask;
}
uint32_t ora_shard_owner(int16_t x, int16_t y, int16_t z, int shift, int world) {
  return shard_owner(S3{x, y, z}, shift, world);
}

static SE3 pose_of(const float* p) { return SE3{Quat{p[0], p[1], p[2], p[3]}, V3{p[4], p[5], p[6]}}; }

// pose = {qw,qx,qy,qz,tx,ty,tz} of cam_T_world; K = {fx,fy,cx,cy}
void ora_integrate(ora_grid* o, const uint8_t* rgb, const float* depth, const float* ht, const float* lt, int H, int W,
                   float max_depth, const float* K, const float* pose) {
  integrate(o->g, rgb, depth, ht, lt, H, W, max_depth, Intr{K[0], K[1], K[2], K[3]}, pose_of(pose));
}
// K-class extension: class_prob = [K][H*W] per-class probabilities of the frame.  The binary ht / lt channel sees
// ht = lt = plane 0, i.e. stays at its prior, exactly as the product's rsf_integrate_classes* does.
void ora_integrate_classes(ora_grid* o, const uint8_t* rgb, const float* depth, const float* class_prob, int num_classes,
                           int H, int W, float max_depth, const float* K, const float* pose) {
  o->g.num_classes = num_classes;
  o->g.cls_in = class_prob;
  integrate(o->g, rgb, depth, class_prob, class_prob, H, W, max_depth, Intr{K[0], K[1], K[2], K[3]}, pose_of(pose));
  o->g.cls_in = nullptr;
}
// rows (x, y, z, tsdf, p_0 .. p_{K-1}); block order = sorted key order; rows may be NULL to size
int64_t ora_gather_valid_classes(ora_grid* o, float* rows) {
  const Grid& g = o->g;
  const int Kc = g.num_classes, nc = 4 + Kc;
  const auto v = sorted_blocks(g);
  if (rows)
    for (size_t bi = 0; bi < v.size(); ++bi)
      for (int i = 0; i < kBlockVol; ++i) {
        const int ox = i & 7, oy = (i >> 3) & 7, oz = i >> 6;
        float* r = rows + (bi * kBlockVol + i) * nc;
        r[0] = (float)(int16_t)((v[bi]->pos.x << 3) + ox) * g.vs;
        r[1] = (float)(int16_t)((v[bi]->pos.y << 3) + oy) * g.vs;
        r[2] = (float)(int16_t)((v[bi]->pos.z << 3) + oz) * g.vs;
        r[3] = v[bi]->tsdf[i];
        for (int k = 0; k < Kc; ++k) r[4 + k] = v[bi]->cls.empty() ? 1.f / (float)Kc : v[bi]->cls[(size_t)k * kBlockVol + i];
      }
  return (int64_t)v.size() * kBlockVol;
}
void ora_raycast(ora_grid* o, float max_depth, int H, int W, const float* K, const float* pose, uint8_t* rgba,
                 uint8_t* normal, float* depth, float* prob) {
  raycast(o->g, max_depth, H, W, Intr{K[0], K[1], K[2], K[3]}, pose_of(pose), rgba, normal, depth, prob);
}
void ora_stats(ora_grid* o, int64_t* out4) {
  out4[0] = o->g.stats.new_blocks;
  out4[1] = o->g.stats.visible_blocks;
  out4[2] = o->g.stats.updated_voxels;
  out4[3] = o->g.stats.carved_blocks;
}
int64_t ora_num_blocks(ora_grid* o) { return (int64_t)o->g.map.size(); }

// direct table access, for the re-expression of utils/tests/voxel_hash_test.cu
void ora_allocate_block(ora_grid* o, int16_t x, int16_t y, int16_t z) {
  auto& slot = o->g.map[pack(S3{x, y, z})];
  if (!slot) {
    slot.reset(new Block);
    slot->pos = S3{x, y, z};
    slot->reset();
  }
}
void ora_delete_block(ora_grid* o, int16_t x, int16_t y, int16_t z) { o->g.map.erase(pack(S3{x, y, z})); }
// returns 1 when the voxel's block exists; out = {tsdf, prob, r, g, b, weight} (miss defaults otherwise)
int ora_retrieve(ora_grid* o, int16_t x, int16_t y, int16_t z, float* out6) {
  Probe pr(o->g);
  const S3 p{x, y, z};
  uint8_t c[3];
  pr.rgb(p, c);
  out6[0] = pr.tsdf(p);
  out6[1] = pr.prob(p);
  out6[2] = c[0];
  out6[3] = c[1];
  out6[4] = c[2];
  out6[5] = (float)pr.weight(p);
  return pr.block_of(p) != nullptr;
}
int ora_write_voxel(ora_grid* o, int16_t x, int16_t y, int16_t z, float tsdf, float prob, const uint8_t* rgbw) {
  const S3 p{x, y, z};
  auto it = o->g.map.find(pack(blk(p)));
  if (it == o->g.map.end()) return 0;
  const int i = off_idx(p);
  it->second->tsdf[i] = tsdf;
  it->second->prob[i] = prob;
  std::memcpy(it->second->rgbw + 4 * i, rgbw, 4);
  return 1;
}
float ora_trilinear(ora_grid* o, float x, float y, float z) {
  Probe pr(o->g);
  return pr.tri(V3{x, y, z});
}

// blocks sorted by packed key (x | y<<16 | z<<32, unsigned 16-bit fields); any output may be NULL
int64_t ora_export_blocks(ora_grid* o, int owned_only, int16_t* coords, float* tsdf, uint8_t* rgbw, float* prob) {
  const auto v = sorted_blocks(o->g, owned_only != 0);
  for (size_t i = 0; i < v.size(); ++i) {
    if (coords) {
      coords[3 * i + 0] = v[i]->pos.x;
      coords[3 * i + 1] = v[i]->pos.y;
      coords[3 * i + 2] = v[i]->pos.z;
    }
    if (tsdf) std::memcpy(tsdf + i * kBlockVol, v[i]->tsdf, sizeof(float) * kBlockVol);
    if (rgbw) std::memcpy(rgbw + i * kBlockVol * 4, v[i]->rgbw, kBlockVol * 4);
    if (prob) std::memcpy(prob + i * kBlockVol, v[i]->prob, sizeof(float) * kBlockVol);
  }
  return (int64_t)v.size();
}

// voxel_tsdf.cu:35-62,476-530: rows (x,y,z,tsdf[,prob]); block order here = sorted key order
static int64_t gather_rows(const Grid& g, const std::vector<const Block*>& v, int ncol, float* rows) {
  if (rows)
    for (size_t bi = 0; bi < v.size(); ++bi)
      for (int i = 0; i < kBlockVol; ++i) {
        const int ox = i & 7, oy = (i >> 3) & 7, oz = i >> 6;
        float* r = rows + (bi * kBlockVol + i) * ncol;
        r[0] = (float)(int16_t)((v[bi]->pos.x << 3) + ox) * g.vs;
        r[1] = (float)(int16_t)((v[bi]->pos.y << 3) + oy) * g.vs;
        r[2] = (float)(int16_t)((v[bi]->pos.z << 3) + oz) * g.vs;
        r[3] = v[bi]->tsdf[i];
        if (ncol == 5) r[4] = v[bi]->prob[i];
      }
  return (int64_t)v.size() * kBlockVol;
}
int64_t ora_gather_valid(ora_grid* o, float* rows) { return gather_rows(o->g, sorted_blocks(o->g), 4, rows); }
int64_t ora_gather_valid_semantic(ora_grid* o, float* rows) {
  return gather_rows(o->g, sorted_blocks(o->g), 5, rows);
}
// voxel_tsdf.cu:15-26,532-559 and voxel_tsdf.cuh:28-33: bounds*(1/vs) cast to short, block fully inside
int64_t ora_gather_voxels(ora_grid* o, const float* bounds6, float* rows) {
  const float s = (float)(1. / o->g.vs);
  int16_t b[6];
  for (int i = 0; i < 6; ++i) b[i] = (int16_t)(bounds6[i] * s);
  std::vector<const Block*> sel;
  for (const Block* B : sorted_blocks(o->g)) {
    const int16_t x = (int16_t)(B->pos.x << 3), y = (int16_t)(B->pos.y << 3), z = (int16_t)(B->pos.z << 3);
    if (x >= b[0] && y >= b[2] && z >= b[4] && x + 7 <= b[1] && y + 7 <= b[3] && z + 7 <= b[5]) sel.push_back(B);
  }
  return gather_rows(o->g, sel, 4, rows);
}

// ---- TSDFGrid::GatherValidMesh, voxel_tsdf.cu:561-845 (marching_cube_kernel + the two compactions) -------------
// Tables: corner offsets / edges / edge->lattice vertex in the reference's numbering (mcube_table.cuh:14-31); the
// 256-case triangle table is the classic public one, shared with the product as packed data (rsf_mc_table.h).
static const int kMcCorner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}};
static const int kMcDimCorner[3] = {1, 4, 3};                                   // offset_dim_indics
static const int kMcEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int kMcEdgeVertex[12][2] = {{0, 0}, {1, 2}, {3, 0}, {0, 2}, {4, 0}, {5, 2}, {7, 0}, {4, 2}, {0, 1}, {1, 1}, {2, 1}, {3, 1}};
static const uint64_t kMcTri[256] = {RSF_MC_TRI_TABLE};
static inline int mc_tri(int cubeindex, int k) {      // tri_table[cubeindex][k], -1 past the end
  if (k >= 15) return -1;
  const int e = (int)((kMcTri[cubeindex] >> (4 * k)) & 15u);
  return e == 15 ? -1 : e;
}

// Outputs (any may be NULL to size): vertices nv x 3, vertex_prob nv, triangles nt x 3.  Order: blocks in sorted
// key order; inside a block, vertices by lattice index (cube_idx * 3 + dim), triangles by voxel then slot --
// what the reference's stable compaction produces for its own block order.
int64_t ora_gather_valid_mesh(ora_grid* o, float* vertices, float* vertex_prob, int32_t* triangles, int64_t* n_triangles) {
  const Grid& g = o->g;
  int64_t nv = 0, nt = 0;
  static thread_local std::vector<float> cube_t, cube_p;
  cube_t.assign(16 * 16 * 16, 0.f);
  cube_p.assign(16 * 16 * 16, 0.f);
  auto C = [](int z, int y, int x) { return (z * 16 + y) * 16 + x; };
  for (const Block* B : sorted_blocks(g)) {
    // :589-635 the 2x2x2 block neighbourhood; weight <= 10 or missing block -> tsdf -10, prob 0
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          const Block* N = g.find(S3{(int16_t)(B->pos.x + k), (int16_t)(B->pos.y + j), (int16_t)(B->pos.z + i)});
          for (int v = 0; v < kBlockVol; ++v) {
            const int x = v & 7, y = (v >> 3) & 7, z = v >> 6;
            float t = -10.f, p = 0.f;
            if (N && (int)N->rgbw[4 * v + 3] > 10) {
              t = N->tsdf[v];
              p = N->prob[v];
            }
            cube_t[C(i * 8 + z, j * 8 + y, k * 8 + x)] = t;
            cube_p[C(i * 8 + z, j * 8 + y, k * 8 + x)] = p;
          }
        }
    // :672-712 triangles, marking referenced lattice vertices
    std::vector<uint8_t> vmask(9 * 9 * 9 * 3, 0);
    std::vector<int> tri;   // lattice ids, 3 per valid triangle
    for (int v = 0; v < kBlockVol; ++v) {
      const int x = v & 7, y = (v >> 3) & 7, z = v >> 6;
      float lt[8];
      int cubeindex = 0;
      for (int i = 0; i < 8; ++i) {
        lt[i] = cube_t[C(z + kMcCorner[i][2], y + kMcCorner[i][1], x + kMcCorner[i][0])];
        cubeindex |= ((lt[i] < 0) << i);
      }
      for (int i = 0; i < 5; ++i) {
        if (mc_tri(cubeindex, i * 3) == -1) continue;
        bool valid = true;
        int ids[3];
        for (int j = 0; j < 3; ++j) {
          const int e = mc_tri(cubeindex, i * 3 + j);
          const int v1 = kMcEdge[e][0], v2 = kMcEdge[e][1];
          if (std::fabs(lt[v2] - lt[v1]) < 1e-3 || std::fabs(lt[v2] - lt[v1]) >= 2) {
            valid = false;
            break;
          }
          const int lc = kMcEdgeVertex[e][0], dim = kMcEdgeVertex[e][1];
          const int cell = (kMcCorner[lc][2] + z) * 81 + (kMcCorner[lc][1] + y) * 9 + (kMcCorner[lc][0] + x);
          ids[j] = cell * 3 + dim;
          vmask[ids[j]] = 1;
        }
        if (valid) tri.insert(tri.end(), ids, ids + 3);
      }
    }
    // :639-668 vertices + compaction (stable, by lattice index)
    std::vector<int> rank(vmask.size(), -1);
    for (size_t l = 0; l < vmask.size(); ++l) {
      if (!vmask[l]) continue;
      rank[l] = (int)nv;
      if (vertices) {
        const int cell = (int)l / 3, j = (int)l % 3;
        const int vx = cell % 9, vy = cell / 9 % 9, vz = cell / 81;
        const int* off = kMcCorner[kMcDimCorner[j]];
        const float t1 = cube_t[C(vz, vy, vx)], t2 = cube_t[C(vz + off[2], vy + off[1], vx + off[0])];
        const float p1 = cube_p[C(vz, vy, vx)], p2 = cube_p[C(vz + off[2], vy + off[1], vx + off[0])];
        const float s = (-t1) / (t2 - t1);
        const float v1[3] = {(float)(int16_t)((int16_t)(B->pos.x << 3) + vx), (float)(int16_t)((int16_t)(B->pos.y << 3) + vy),
                             (float)(int16_t)((int16_t)(B->pos.z << 3) + vz)};
        for (int c = 0; c < 3; ++c) vertices[3 * nv + c] = (v1[c] + s * (float)off[c]) * g.vs;
        if (vertex_prob) vertex_prob[nv] = (p1 + p2) / 2;
      }
      ++nv;
    }
    for (size_t k = 0; k < tri.size(); k += 3) {
      if (triangles)
        for (int j = 0; j < 3; ++j) triangles[3 * nt + j] = rank[tri[k + j]];
      ++nt;
    }
  }
  if (n_triangles) *n_triangles = nt;
  return nv;
}

}  // extern "C"

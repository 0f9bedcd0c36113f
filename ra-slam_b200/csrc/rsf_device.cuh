// rsf_device.cuh -- PODs, camera / rigid-body math and the lock-free block table (device side).
//
// Replaces utils/cuda/{camera,lie_group}.cuh (Eigen-based) and the device halves of
// utils/tsdf/voxel_hash.{cuh,cu} / voxel_mem.{cuh,cu} of the reference.
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, IEEE div/sqrt, so
// every float operation is rounded on its own, in the order the reference source writes it
// (and oracle/rsf_oracle.cpp restates it).  That makes block membership, voxel selection,
// TSDF, weight and colour bit-reproducible against the oracle; only expf/logf (semantic
// fusion) differ by ulps.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rsf {

constexpr int kBlockLen = 8;
constexpr int kBlockVol = 512;
// One voxel block = one contiguous 6 KiB slab: [tsdf f32 x512 | rgbw u8x4 x512 | logit f32 x512].
// (The reference keeps three separate 512 MiB SoA pools, voxel_mem.cu:19-21; a slab keeps one
// block's state in 3 consecutive DRAM pages and lets a CTA stream it with 128-bit accesses.)
//
// The third plane holds logit(p) = ln p - ln(1-p) instead of the reference's probability p
// (VoxelSEGM, voxel_types.cuh:36-43).  The reference's fusion (voxel_tsdf.cu:241-248)
//     pos = exp((wo ln p + wn ln ht)/w), neg = exp((wo ln(1-p) + wn ln lt)/w), p' = pos/(pos+neg)
// is exactly logit(p') = (wo logit(p) + wn (ln ht - ln lt))/w, a running average with no
// transcendental per voxel (ln ht - ln lt is computed once per pixel).  Every API output converts
// back with p = 1/(1+exp(-logit)); measured |dp| vs the reference recurrence < 1e-6.
constexpr int kSlabBytes = 3 * kBlockVol * 4;
constexpr int kSlabTsdfOff = 0;
constexpr int kSlabRgbwOff = kBlockVol * 4;
constexpr int kSlabLogitOff = 2 * kBlockVol * 4;

constexpr uint64_t kEmptyKey = ~0ull;
constexpr int kValFree = -1;      // key present (or slot empty) but no block attached
constexpr int kValPending = -2;   // a thread is attaching a block right now

constexpr uint32_t kMetaAlive = 1u;
constexpr uint32_t kMetaFresh = 2u;   // slab contents not yet materialised (see integrate kernel)
// Set on blocks whose voxels were written by the table-level hooks (rsf_allocate_blocks / rsf_assign_voxels) and have
// not been through the integrate kernel's carve test since: such a block may hold min|tsdf| >= 0.9 and must reach the
// integrate kernel (which carves it, voxel_tsdf.cu:253-276) even when no voxel of it can be updated.  See
// select_visible_kernel's occlusion culling.
constexpr uint32_t kMetaNoCull = 4u;

constexpr unsigned kErrPoolFull = 1u;
constexpr unsigned kErrTableFull = 2u;

struct V3 {
  float x, y, z;
};
struct S3 {
  short x, y, z;
};
struct Quat {
  float w, x, y, z;
};

__host__ __device__ inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__host__ __device__ inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
__host__ __device__ inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }   // Eigen redux order: c0 + (c1 + c2)
__host__ __device__ inline float norm(V3 a) { return sqrtf(a.x * a.x + (a.y * a.y + a.z * a.z)); }
__host__ __device__ inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3 QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 (q.vec x v)
__host__ __device__ inline V3 qrot(Quat q, V3 v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  const V3 c = cross(qv, uv);
  return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
__host__ __device__ inline Quat qinv(Quat q) {
  const float n2 = (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);   // Eigen redux order for 4 coefficients
  if (n2 > 0.f) return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return {0.f, 0.f, 0.f, 0.f};
}

// SE3<float> (utils/cuda/lie_group.cuh:8-45)
struct Pose {
  Quat q;
  V3 t;
  __host__ __device__ V3 apply(V3 v) const { return qrot(q, v) + t; }
  __host__ __device__ Pose inverse() const {
    const Quat qi = qinv(q);
    return {qi, qrot(qi, neg(t))};
  }
};

// CameraIntrinsics<float> (utils/cuda/camera.cuh:13-52); the same product formula is used for
// K and K^-1 = {1/fx, 1/fy, -cx/fx, -cy/fy}.
struct Intr {
  float fx, fy, cx, cy;
  __host__ __device__ Intr inverse() const {
    const float fxi = 1.f / fx, fyi = 1.f / fy;
    return {fxi, fyi, -cx * fxi, -cy * fyi};
  }
  __host__ __device__ V3 mul(V3 v) const { return {fx * v.x + cx * v.z, fy * v.y + cy * v.z, v.z}; }
};

// Everything a kernel needs to know about one frame / one virtual camera; passed by value.
struct Frame {
  Intr K, Ki;
  Pose T_cw, T_wc;
  int W, H;
  float vs, tr, max_depth;
};

// Device-resident counters (one 128-byte line).  Bytes [16, 48) are per-frame and are cleared with
// one memset before each integrate; the rest persists for the life of the grid.
struct Counters {
  int free_count;      // stack pointer into heap[]: heap[0 .. free_count) are free pool slots
  int high_water;      // 1 + largest pool index ever handed out
  unsigned error;      // sticky kErr* bits
  int table_keys;      // table slots whose key is claimed (alive blocks + ghosts of carved ones): the probe-chain load
  int vis_count;       // length of the visible list built by the select kernel
  int culled;          // visible blocks the select kernel proved un-updatable this frame (occluded) and left out
  unsigned long long new_blocks, updated_voxels, carved_blocks;
  int gather_count;    // length of the list built by the gather-select kernel
  int pad2;
  // running totals over all frames (folded in by begin_frame_kernel); used by the benchmark
  unsigned long long tot_frames, tot_new_blocks, tot_visible_blocks, tot_updated_voxels, tot_carved_blocks;
  unsigned ray_work;   // next 8x4 pixel patch of the persistent raycast kernel
  unsigned ray_work2;  // same for the second row band of a banded raycast (rsf_raycast to host buffers)
  unsigned pad3[6];
};
static_assert(sizeof(Counters) == 128, "Counters must stay one 128-byte line");
constexpr size_t kCtrFrameOff = 16, kCtrFrameBytes = 32, kCtrGatherOff = 48;

struct BlockMeta {     // 16 B per pool slot: the dense allocated-block list
  uint64_t key;        // packed block coordinate
  int slot;            // hash-table slot that points here (so carving needs no probe)
  uint32_t flags;      // kMetaAlive | kMetaFresh
};

struct ShardSpec {
  int rank, world, shift, halo;
};

// One hash-table slot: key and value side by side so a probe is a single 128-bit load.
struct __align__(16) TableEntry {
  uint64_t key;        // packed block coordinate or kEmptyKey
  int val;             // pool index, kValFree or kValPending
  int pad;
};

// One visible-list entry: everything the integrate kernel needs to find the slab (no meta[] hop).
struct __align__(16) VisEntry {
  uint64_t key;
  int slot;
  int idx_fresh;       // pool index | fresh << 31
};

// One pixel of the staged frame: 16 B, gathered with one 128-bit load per voxel.
//   x = depth [m], y = rgb (r | g<<8 | b<<16 bits), z = ln ht - ln lt, w = depth_to_range = |K^-1 (u,v,1)|
typedef float4 FramePixel;

// View of one grid's device memory; passed to kernels by value.
struct GridView {
  TableEntry* table;   // [table_slots]
  uint32_t slot_mask;  // table_slots - 1
  uint8_t* pool;       // [pool_blocks * kSlabBytes]
  BlockMeta* meta;     // [pool_blocks]
  int* heap;           // [pool_blocks] free-list stack
  int pool_blocks;
  Counters* ctr;
  VisEntry* vis_list;  // [pool_blocks] blocks selected for integration
  int* gather_list;    // [pool_blocks]
  ShardSpec shard;
};

// ---- integer helpers (utils/tsdf/voxel_mem.cuh:31-70) ------------------------------------------
__host__ __device__ inline short rnd_s(float v) { return (short)(int)roundf(v); }
__host__ __device__ inline S3 rnd3(V3 p) { return {rnd_s(p.x), rnd_s(p.y), rnd_s(p.z)}; }
__host__ __device__ inline S3 point_to_block(S3 g) {
  return {(short)(g.x >> 3), (short)(g.y >> 3), (short)(g.z >> 3)};
}
__host__ __device__ inline int point_to_index(S3 g) { return (g.x & 7) + (g.y & 7) * 8 + (g.z & 7) * 64; }
__host__ __device__ inline uint64_t pack_key(S3 b) {
  return (uint64_t)(uint16_t)b.x | ((uint64_t)(uint16_t)b.y << 16) | ((uint64_t)(uint16_t)b.z << 32);
}
__host__ __device__ inline S3 unpack_key(uint64_t k) {
  return {(short)(k & 0xffff), (short)((k >> 16) & 0xffff), (short)((k >> 32) & 0xffff)};
}
// Hash (utils/tsdf/voxel_hash.cu:19-23): short -> uint sign-extends.
__host__ __device__ inline uint32_t hash_block(S3 b, uint32_t mask) {
  return (((uint32_t)(int)b.x * 73856093u) ^ ((uint32_t)(int)b.y * 19349669u) ^ ((uint32_t)(int)b.z * 83492791u)) & mask;
}

// ---- spatial-shard ownership (new) ----------------------------------------------
__host__ __device__ inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
__host__ __device__ inline uint32_t shard_owner_brick(int x, int y, int z, int world) {
  const uint32_t h =
      mix32((uint32_t)x * 0x9e3779b1u ^ mix32((uint32_t)y * 0x85ebca77u ^ mix32((uint32_t)z * 0xc2b2ae3du)));
  return h % (uint32_t)world;
}
__host__ __device__ inline uint32_t shard_owner(S3 b, int shift, int world) {
  return shard_owner_brick(b.x >> shift, b.y >> shift, b.z >> shift, world);
}
__host__ __device__ inline bool shard_owned(const ShardSpec& s, S3 b) {
  return s.world <= 1 || (int)shard_owner(b, s.shift, s.world) == s.rank;
}
// owned, or within `halo` blocks of a block this rank owns.  The (2 halo + 1)^3 neighbours fall into at most
// 2 x 2 x 2 distinct super-bricks (halo < brick size), a single one for the (1 - 2 halo / 2^shift)^3 of blocks in
// a brick's interior: test the bricks, not the 27 neighbours.
__host__ __device__ inline bool shard_wanted(const ShardSpec& s, S3 b) {
  if (s.world <= 1) return true;
  const int x0 = (short)(b.x - s.halo) >> s.shift, x1 = (short)(b.x + s.halo) >> s.shift;
  const int y0 = (short)(b.y - s.halo) >> s.shift, y1 = (short)(b.y + s.halo) >> s.shift;
  const int z0 = (short)(b.z - s.halo) >> s.shift, z1 = (short)(b.z + s.halo) >> s.shift;
  if ((1 << s.shift) <= 2 * s.halo || x1 < x0 || y1 < y0 || z1 < z0) {   // tiny bricks or int16 wrap-around: brute force
    for (int dz = -s.halo; dz <= s.halo; ++dz)
      for (int dy = -s.halo; dy <= s.halo; ++dy)
        for (int dx = -s.halo; dx <= s.halo; ++dx) {
          const S3 n{(short)(b.x + dx), (short)(b.y + dy), (short)(b.z + dz)};
          if ((int)shard_owner(n, s.shift, s.world) == s.rank) return true;
        }
    return false;
  }
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y)
      for (int x = x0; x <= x1; ++x)
        if ((int)shard_owner_brick(x, y, z, s.world) == s.rank) return true;
  return false;
}

#ifdef __CUDACC__
// ---- visibility (utils/tsdf/voxel_tsdf.cu:64-96) ----------------------------------------------
__device__ inline bool voxel_visible(S3 g, const Frame& f) {
  const V3 pw{(float)g.x * f.vs, (float)g.y * f.vs, (float)g.z * f.vs};
  const V3 pc = f.T_cw.apply(pw);
  const V3 h = f.K.mul(pc);
  const float u = h.x / h.z, v = h.y / h.z;
  return (u >= 0 && u <= (float)(f.W - 1) && v >= 0 && v <= (float)(f.H - 1) && h.z >= 0);
}
template <bool Full>
__device__ inline bool block_visible(S3 b, const Frame& f) {
  const short x = (short)(b.x << 3), y = (short)(b.y << 3), z = (short)(b.z << 3);
  bool vis = Full;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const S3 c{(short)(x + ((i >> 0) & 1) * (kBlockLen - 1)), (short)(y + ((i >> 1) & 1) * (kBlockLen - 1)),
               (short)(z + ((i >> 2) & 1) * (kBlockLen - 1))};
    if (Full)
      vis &= voxel_visible(c, f);
    else
      vis |= voxel_visible(c, f);
  }
  return vis;
}

// ---- lock-free open-addressed block table ------------------------------------------------------
// entry.key only ever goes kEmptyKey -> key (64-bit CAS) and is never cleared: a carved block
// leaves its key behind with val == kValFree ("ghost"), so re-allocating the same coordinate next
// frame reuses the slot and linear probing needs no tombstones.  val goes
// kValFree -> kValPending -> pool index (allocation) and pool index -> kValFree (carving, in a
// different kernel).  Lookups never run concurrently with inserts (separate launches).

__device__ __forceinline__ TableEntry load_entry_nc(const TableEntry* e) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(e));
  TableEntry r;
  r.key = (uint64_t)(uint32_t)v.x | ((uint64_t)(uint32_t)v.y << 32);
  r.val = v.z;
  r.pad = 0;
  return r;
}
__device__ __forceinline__ TableEntry load_entry_volatile(const TableEntry* e) {
  int4 v;
  asm volatile("ld.volatile.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(e));
  TableEntry r;
  r.key = (uint64_t)(uint32_t)v.x | ((uint64_t)(uint32_t)v.y << 32);
  r.val = v.z;
  r.pad = 0;
  return r;
}

// Read-only lookup: pool index of block `key`, or -1 (miss / ghost).
__device__ inline int table_lookup(const GridView& g, uint64_t key, uint32_t home) {
  uint32_t s = home;
  for (uint32_t probe = 0; probe <= g.slot_mask; ++probe) {
    const TableEntry e = load_entry_nc(&g.table[s]);
    if (e.key == key) return e.val >= 0 ? e.val : -1;
    if (e.key == kEmptyKey) return -1;
    s = (s + 1) & g.slot_mask;
  }
  return -1;
}
// Coherent probe (usable inside kernels that also insert), continuing from an already loaded
// entry `e` of slot `s`.  Returns the slot holding `key` (and its value in *val) or -1.
__device__ inline int table_find_slot_from(const GridView& g, uint64_t key, uint32_t s, TableEntry e, int* val) {
  for (uint32_t probe = 0; probe <= g.slot_mask; ++probe) {
    if (e.key == key) {
      *val = e.val;
      return (int)s;
    }
    if (e.key == kEmptyKey) return -1;
    s = (s + 1) & g.slot_mask;
    e = load_entry_volatile(&g.table[s]);
  }
  return -1;
}
// Find the slot of `key`, claiming an empty one if absent.  Returns -1 when the table is full.
__device__ inline int table_find_or_claim_slot(const GridView& g, uint64_t key, uint32_t home) {
  uint32_t s = home;
  for (uint32_t probe = 0; probe <= g.slot_mask; ++probe) {
    uint64_t k = *(volatile uint64_t*)&g.table[s].key;
    if (k == kEmptyKey) {
      k = atomicCAS((unsigned long long*)&g.table[s].key, (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (k == kEmptyKey) {
        atomicAdd(&g.ctr->table_keys, 1);
        return (int)s;
      }
    }
    if (k == key) return (int)s;
    s = (s + 1) & g.slot_mask;
  }
  return -1;
}

// p = 1 / (1 + exp(-logit)): logit -inf -> 0, +inf -> 1, 0 -> exactly 0.5, NaN -> NaN
__device__ __forceinline__ float logit_to_prob(float l) { return 1.f / (1.f + expf(-l)); }
__device__ __forceinline__ float prob_to_logit(float p) { return logf(p) - logf(1.f - p); }
#endif  // __CUDACC__

}  // namespace rsf

// rsf_mesh.cu -- marching cubes over the allocated blocks (TSDFGrid::GatherValidMesh, reference
// utils/tsdf/voxel_tsdf.cu:561-845, tables utils/tsdf/mcube_table.cuh).
//
// What the reference computes, per allocated block: the 16^3 voxel neighbourhood of the block and its 7
// positive neighbours (voxels with weight <= 10 or in unallocated blocks read tsdf -10 / prob 0); one cube per
// voxel with corners in the reference's vertex order; up to 5 triangles per cube from the classic triangle
// table, a triangle being dropped when one of its edges has |dtsdf| < 1e-3 or >= 2; vertices live on a 9^3
// lattice per block x 3 edge directions, position = (v1 + (-t1)/(t2-t1) e_dim) * voxel_size, probability =
// mean of the two end voxels; only vertices referenced by a triangle are emitted (including those referenced
// by a triangle before the edge that got it dropped, as in the reference).
//
// How it is computed here: the reference materialises worst-case arrays (2187 vertices + 2560 triangles per
// block, 56 KiB/block), runs two full prefix sums and four compaction launches.  This file does count -> scan
// -> emit: a 512-thread CTA per block builds the neighbourhood in shared memory, marks referenced lattice
// vertices in a 2187-bit shared bitmap and counts triangles; a scan turns per-block counts into output offsets;
// the emit pass recomputes the cubes (16 KiB of L2-resident reads) and writes compact, final arrays directly.
// Scratch is 16 B per block.
#include "rsf_kernels.h"
#include "rsf_mc_table.h"

namespace rsf {
namespace {

// Cube corner offsets (x, y, z) in the reference's vertex order (mcube_table.cuh:14-19) and the 12 edges as
// corner pairs (:25-27).  For an edge, the lattice vertex it lives on is its lower corner and a direction.
__constant__ int8_t kCorner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 1, 0}, {1, 1, 1}, {0, 1, 1}};
__constant__ int8_t kEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
// {lower corner, direction (0 = x, 1 = y, 2 = z)} per edge (mcube_table.cuh:29-31)
__constant__ int8_t kEdgeVertex[12][2] = {{0, 0}, {1, 2}, {3, 0}, {0, 2}, {4, 0}, {5, 2}, {7, 0}, {4, 2}, {0, 1}, {1, 1}, {2, 1}, {3, 1}};
// The classic 256-case triangle table (Lorensen & Cline / Bourke; the reference's tri_table, mcube_table.cuh:75),
// packed: nibble k of word c = k-th edge index of case c, 0xf = end of list.
__constant__ uint64_t kTriTable[256] = {RSF_MC_TRI_TABLE};

constexpr int kLat = 9;                        // lattice points per axis (BLOCK_VERT_LEN)
constexpr int kLatVol = kLat * kLat * kLat;    // 729
constexpr int kLatVerts = kLatVol * 3;         // 2187 candidate vertices per block
constexpr int kFlagWords = (kLatVerts + 31) / 32;   // 69

struct MeshShared {
  float tsdf[16][16][16];                      // [z][y][x]
  const uint8_t* slab[8];                      // [dz][dy][dx] -> index dz*4 + dy*2 + dx
  unsigned flags[kFlagWords];
  unsigned word_prefix[kFlagWords + 1];
  int warp_sum[16];
  int tri_total;
};

// Loads the 16^3 neighbourhood; `prob` (optional, emit pass) gets the fused probability of the same voxels.
__device__ __forceinline__ void load_cube(const GridView& g, S3 b, MeshShared& sh, float (*prob)[16][16]) {
  const int tid = threadIdx.x;
  if (tid < 8) {
    const S3 nb{(short)(b.x + (tid & 1)), (short)(b.y + ((tid >> 1) & 1)), (short)(b.z + (tid >> 2))};
    const int idx = table_lookup(g, pack_key(nb), hash_block(nb, g.slot_mask));
    sh.slab[tid] = idx >= 0 ? g.pool + (size_t)idx * kSlabBytes : nullptr;
  }
  for (int w = tid; w < kFlagWords; w += blockDim.x) sh.flags[w] = 0u;
  __syncthreads();
  const int x = tid & 7, y = (tid >> 3) & 7, z = tid >> 6;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const uint8_t* s = sh.slab[n];
    float t = -10.f, p = 0.f;
    if (s) {
      const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(s + kSlabRgbwOff) + tid);
      if ((int)(c >> 24) > 10) {               // voxel_tsdf.cu:606: more than 10 observations
        t = __ldg(reinterpret_cast<const float*>(s + kSlabTsdfOff) + tid);
        if (prob) p = logit_to_prob(__ldg(reinterpret_cast<const float*>(s + kSlabLogitOff) + tid));
      }
    }
    const int cz = (n >> 2) * 8 + z, cy = ((n >> 1) & 1) * 8 + y, cx = (n & 1) * 8 + x;
    sh.tsdf[cz][cy][cx] = t;
    if (prob) prob[cz][cy][cx] = p;
  }
  __syncthreads();
}

// Walks the triangles of this thread's cube.  Marks referenced lattice vertices in sh.flags, returns the number
// of valid triangles and their lattice-vertex ids (local_ids[3*k + j]).
__device__ __forceinline__ int cube_triangles(MeshShared& sh, int* local_ids) {
  const int tid = threadIdx.x;
  const int x = tid & 7, y = (tid >> 3) & 7, z = tid >> 6;
  float lt[8];
  int cubeindex = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    lt[i] = sh.tsdf[z + kCorner[i][2]][y + kCorner[i][1]][x + kCorner[i][0]];
    cubeindex |= (lt[i] < 0) << i;
  }
  const uint64_t row = kTriTable[cubeindex];
  int n = 0;
  for (int i = 0; i < 5; ++i) {
    if (((row >> (12 * i)) & 15u) == 15u) break;
    int ids[3];
    bool valid = true;
    for (int j = 0; j < 3; ++j) {
      const int e = (int)((row >> (12 * i + 4 * j)) & 15u);
      const float d = fabsf(lt[kEdge[e][1]] - lt[kEdge[e][0]]);
      // reference: d < 1e-3 (double literal) || d >= 2; for a float d, d < 1e-3 <=> d < 0.001f
      if (d < 0.001f || d >= 2.f) {
        valid = false;
        break;
      }
      const int lc = kEdgeVertex[e][0], dim = kEdgeVertex[e][1];
      const int cell = (kCorner[lc][2] + z) * (kLat * kLat) + (kCorner[lc][1] + y) * kLat + (kCorner[lc][0] + x);
      ids[j] = cell * 3 + dim;
      atomicOr(&sh.flags[ids[j] >> 5], 1u << (ids[j] & 31));   // marked even if a later edge drops the triangle
    }
    if (valid) {
      local_ids[3 * n] = ids[0];
      local_ids[3 * n + 1] = ids[1];
      local_ids[3 * n + 2] = ids[2];
      ++n;
    }
  }
  return n;
}

// block-wide exclusive scan of one int per thread (512 threads); total in sh.tri_total
__device__ __forceinline__ int block_exclusive_scan(MeshShared& sh, int v) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) sh.warp_sum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < 16 ? sh.warp_sum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < 16) sh.warp_sum[lane] = s;   // inclusive
    if (lane == 15) sh.tri_total = s;
  }
  __syncthreads();
  return inc - v + (w ? sh.warp_sum[w - 1] : 0);
}

__global__ void __launch_bounds__(512) mesh_count_kernel(GridView g, int nblk, int2* __restrict__ counts) {
  __shared__ MeshShared sh;
  const int idx = g.gather_list[blockIdx.x];
  load_cube(g, unpack_key(g.meta[idx].key), sh, nullptr);
  int ids[15];
  const int n = cube_triangles(sh, ids);
  block_exclusive_scan(sh, n);
  // vertices: popcount of the bitmap
  int c = 0;
  if (threadIdx.x < kFlagWords) c = __popc(sh.flags[threadIdx.x]);
  __syncthreads();                            // warp_sum is reused
  const int ntri = sh.tri_total;
  __syncthreads();
  block_exclusive_scan(sh, c);
  if (threadIdx.x == 0) counts[blockIdx.x] = make_int2(sh.tri_total, ntri);   // {vertices, triangles}
}

// Exclusive scan of the per-block {vertices, triangles} counts; one CTA, chunked.  totals[0] = sums.
__global__ void __launch_bounds__(1024) mesh_scan_kernel(const int2* __restrict__ counts, int nblk, int2* __restrict__ offsets,
                                                         long long* __restrict__ totals) {
  __shared__ long long warp_v[32], warp_t[32];
  __shared__ long long carry_v, carry_t;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_v = carry_t = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int2 c = i < nblk ? counts[i] : make_int2(0, 0);
    long long v = c.x, t = c.y;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long a = __shfl_up_sync(0xffffffffu, v, o), b = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) v += a, t += b;
    }
    if (lane == 31) warp_v[w] = v, warp_t[w] = t;
    __syncthreads();
    if (w == 0) {
      long long a = warp_v[lane], b = warp_t[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
        if (lane >= o) a += x, b += y;
      }
      warp_v[lane] = a, warp_t[lane] = b;
    }
    __syncthreads();
    const long long pv = carry_v + (w ? warp_v[w - 1] : 0) + v - c.x, pt = carry_t + (w ? warp_t[w - 1] : 0) + t - c.y;
    // offsets are int: a mesh beyond 2^31 vertices cannot be indexed by the reference's Vector3i either
    if (i < nblk) offsets[i] = make_int2((int)pv, (int)pt);
    __syncthreads();
    if (threadIdx.x == 1023) carry_v += warp_v[31], carry_t += warp_t[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[0] = carry_v, totals[1] = carry_t;
}

__global__ void __launch_bounds__(512)
mesh_emit_kernel(GridView g, float vs, int nblk, const int2* __restrict__ offsets, float* __restrict__ vertices,
                 float* __restrict__ vprob, int* __restrict__ triangles) {
  __shared__ MeshShared sh;
  __shared__ float prob[16][16][16];
  const int tid = threadIdx.x;
  const int idx = g.gather_list[blockIdx.x];
  const S3 b = unpack_key(g.meta[idx].key);
  load_cube(g, b, sh, prob);
  int ids[15];
  const int n = cube_triangles(sh, ids);
  const int trank = block_exclusive_scan(sh, n);
  // rank of every lattice vertex = prefix popcount of the bitmap
  if (tid < 32) {
    int run = 0;
    for (int w0 = 0; w0 < kFlagWords + 1; w0 += 32) {
      const int wi = w0 + tid;
      const int c = wi < kFlagWords ? __popc(sh.flags[wi]) : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (wi <= kFlagWords) sh.word_prefix[wi] = run + inc - c;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  const int2 base = offsets[blockIdx.x];
  // vertices (voxel_tsdf.cu:639-668)
  for (int l = tid; l < kLatVerts; l += blockDim.x) {
    if (!((sh.flags[l >> 5] >> (l & 31)) & 1u)) continue;
    const int rank = sh.word_prefix[l >> 5] + __popc(sh.flags[l >> 5] & ((1u << (l & 31)) - 1u));
    const int cell = l / 3, dim = l - cell * 3;
    const int vx = cell % kLat, vy = (cell / kLat) % kLat, vz = cell / (kLat * kLat);
    const int ox = dim == 0, oy = dim == 1, oz = dim == 2;
    const float t1 = sh.tsdf[vz][vy][vx], t2 = sh.tsdf[vz + oz][vy + oy][vx + ox];
    const float p1 = prob[vz][vy][vx], p2 = prob[vz + oz][vy + oy][vx + ox];
    const float s = (-t1) / (t2 - t1);
    // v1 = (BlockToPoint(block) + vertex_offset).cast<float>(): short arithmetic, then float
    const float fx = (float)(short)((short)(b.x << 3) + vx), fy = (float)(short)((short)(b.y << 3) + vy),
                fz = (float)(short)((short)(b.z << 3) + vz);
    float* o = vertices + 3 * (size_t)(base.x + rank);
    o[0] = (fx + s * (float)ox) * vs;
    o[1] = (fy + s * (float)oy) * vs;
    o[2] = (fz + s * (float)oz) * vs;
    vprob[base.x + rank] = (p1 + p2) / 2;
  }
  // triangles (voxel_tsdf.cu:672-712): lattice ids -> compact vertex ids
  for (int k = 0; k < n; ++k) {
    int* o = triangles + 3 * (size_t)(base.y + trank + k);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int l = ids[3 * k + j];
      o[j] = base.x + sh.word_prefix[l >> 5] + __popc(sh.flags[l >> 5] & ((1u << (l & 31)) - 1u));
    }
  }
}

}  // namespace

void launch_mesh_count(const GridView& g, int nblk, int2* counts, int2* offsets, long long* totals, cudaStream_t s) {
  if (nblk > 0) mesh_count_kernel<<<nblk, 512, 0, s>>>(g, nblk, counts);
  mesh_scan_kernel<<<1, 1024, 0, s>>>(counts, nblk, offsets, totals);
}
void launch_mesh_emit(const GridView& g, float vs, int nblk, const int2* offsets, float* vertices, float* vprob,
                      int* triangles, cudaStream_t s) {
  if (nblk > 0) mesh_emit_kernel<<<nblk, 512, 0, s>>>(g, vs, nblk, offsets, vertices, vprob, triangles);
}

}  // namespace rsf

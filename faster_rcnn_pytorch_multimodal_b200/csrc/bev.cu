// LiDAR BEV rasterisation on the device: the step right before the backbone
// for --net_type lidar.  Replaces the CPU path  roi_data_layer/minibatch.py:428-512  (filter_points,
// spconv.utils.VoxelGeneratorV2.generate, the per-voxel max-height / density / tanh-intensity /
// tanh-elongation scatters and the final transpose) for one frame of points.
//
// What has to be reproduced exactly is ORDER-DEPENDENT integer work:
//   * voxels exist in order of first appearance in the point list and only the first `max_voxels`
//     of them are kept (spconv 1.0 points_to_voxel_3d_np: a new voxel past the cap is skipped, the
//     scan continues);
//   * a voxel keeps its first `max_pts` points in input order;
//   * the three meta channels live per (x, y) column and the reference scatters them with a numpy
//     fancy-index assignment, so among the voxels of one column the LAST voxel in voxel order wins.
// All three are solved without a sort:
//   key     per point: voxel id (floor((p - lo) / size) in fp32, the reference's dtype) and a slot from
//           one atomicAdd on a dense per-voxel counter;
//   alloc   the point that drew slot 0 reserves a contiguous bucket for its voxel and appends the voxel
//           to the voxel list;
//   fill    every point drops its index into its bucket -> each voxel owns the (unordered) list of its
//           point indices, contiguous in memory;
//   first   one warp per voxel: smallest index = the voxel's first appearance, flagged in a per-point
//           bitmap; the number of set bits before it (one-CTA popcount scan over the bitmap words) is the
//           voxel's rank in the reference's order;
//   voxel   one warp per voxel: rank < max_voxels keeps it; if it has more than max_pts points the
//           ones with fewer than max_pts smaller indices are the first max_pts in input order
//           (bisection on the index value over the bucket, O(cnt log n)); max z, count and the
//           intensity / elongation sums are warp reductions; the height slice is written at once and
//           the voxel's rank goes into a per-column atomicMax;
//   meta    one thread per voxel: the voxel whose rank is the column maximum writes the meta channels.
// Traffic per frame at the reference's sizes (700 x 800 x 12 grid, 15 channels, ~180 k points):
// 3.6 MB of points in, 33.6 MB of map out (zero fill + scatter), 27 MB zero fill of the dense counter
// and ~2 random 4-byte atomics / loads per point on it.
#include "common.cuh"

namespace b2d {
namespace bev {

struct Grid {
  float x_lo, x_hi, y_lo, y_hi, z_lo, z_hi;   // filter_points ranges = grid origin (z shifted by -z_lo)
  float voxel_len, voxel_height;
  int nx, ny, nz;
  int max_pts, max_voxels, n_meta, elongation, n_feat;
};

struct Ws {
  int32_t *cnt, *base, *key, *slot, *bucket, *flag, *rank, *vkey, *vfirst, *vrank, *colmax, *counters;
  float* vval;   // [n][3]: density, tanh(mean intensity), tanh(mean elongation)
  size_t bytes, zero_bytes;   // [cnt .. counters) must be zeroed before each frame (colmax: any negative)
};

static Ws carve(void* base, int n, int nx, int ny, int nz) {
  Ws w{};
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  const size_t G = (size_t)nx * ny * nz;
  // zeroed region first: counter grid, first-appearance flags, cursors
  w.cnt = reinterpret_cast<int32_t*>(take(4 * G));
  w.flag = reinterpret_cast<int32_t*>(take(4 * ((size_t)n + 1)));
  w.counters = reinterpret_cast<int32_t*>(take(4 * 4));
  w.zero_bytes = off;
  w.colmax = reinterpret_cast<int32_t*>(take(4 * (size_t)nx * ny));
  w.base = reinterpret_cast<int32_t*>(take(4 * G));
  w.key = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.slot = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.bucket = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.rank = reinterpret_cast<int32_t*>(take(4 * ((size_t)n + 1)));
  w.vkey = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.vfirst = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.vrank = reinterpret_cast<int32_t*>(take(4 * (size_t)n));
  w.vval = reinterpret_cast<float*>(take(4 * 3 * (size_t)n));
  w.bytes = off;
  return w;
}

// voxel id of a point, or -1.  filter_points (minibatch.py:232-235) on the raw coordinates, then the
// voxeliser's own test on floor((p - lo) / size); z is shifted by -z_lo first (:454) and its grid
// starts at 0 (:442-443).  All arithmetic in fp32 without contraction, like the reference.
__device__ __forceinline__ int voxel_key(const float* __restrict__ p, const Grid& g) {
  const float x = p[0], y = p[1], z = p[2];
  if (!(x >= g.x_lo && y >= g.y_lo && z >= g.z_lo && x < g.x_hi && y < g.y_hi && z < g.z_hi)) return -1;
  const float zs = fsub(z, g.z_lo);
  const float cx = floorf(fdiv(fsub(x, g.x_lo), g.voxel_len));
  const float cy = floorf(fdiv(fsub(y, g.y_lo), g.voxel_len));
  const float cz = floorf(fdiv(fsub(zs, 0.0f), g.voxel_height));
  if (cx < 0.0f || cx >= (float)g.nx || cy < 0.0f || cy >= (float)g.ny || cz < 0.0f || cz >= (float)g.nz) return -1;
  return ((int)cx * g.ny + (int)cy) * g.nz + (int)cz;
}

__global__ void __launch_bounds__(256) key_kernel(int n, const float* __restrict__ pts, Grid g, Ws w) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = voxel_key(pts + (size_t)i * g.n_feat, g);
  w.key[i] = k;
  w.slot[i] = k >= 0 ? atomicAdd(&w.cnt[k], 1) : -1;
}

// warp-aggregated: one atomic per warp on each of the two cursors (100 k voxels on two addresses
// serialise otherwise: 79 us -> a few us)
__global__ void __launch_bounds__(256) alloc_kernel(int n, Ws w) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool owner = i < n && w.slot[i] == 0;
  const int k = owner ? w.key[i] : 0;
  const int c = owner ? w.cnt[k] : 0;
  const unsigned m = __ballot_sync(0xffffffffu, owner);
  if (!m) return;
  int inc = c;                                   // inclusive prefix of the bucket sizes over the warp
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  int b0 = 0, v0 = 0;
  if (lane == 31) {
    b0 = atomicAdd(&w.counters[0], inc);
    v0 = atomicAdd(&w.counters[1], __popc(m));
  }
  b0 = __shfl_sync(0xffffffffu, b0, 31);
  v0 = __shfl_sync(0xffffffffu, v0, 31);
  if (owner) {
    w.base[k] = b0 + inc - c;
    w.vkey[v0 + __popc(m & ((1u << lane) - 1u))] = k;
  }
}

__global__ void __launch_bounds__(256) fill_kernel(int n, Ws w) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = w.key[i];
  if (k >= 0) w.bucket[w.base[k] + w.slot[i]] = i;
}

__global__ void __launch_bounds__(256) first_kernel(Ws w) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int nvox = w.counters[1];
  for (int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < nvox; v += nw) {
    const int k = w.vkey[v], cnt = w.cnt[k];
    const int32_t* b = w.bucket + w.base[k];
    int m = 0x7fffffff;
    for (int j = lane; j < cnt; j += 32) m = min(m, b[j]);
    m = (int)__reduce_min_sync(0xffffffffu, (unsigned)m);
    if (lane == 0) {
      w.vfirst[v] = m;
      atomicOr(reinterpret_cast<unsigned*>(&w.flag[m >> 5]), 1u << (m & 31));
    }
  }
}

// flag is a bitmap over the points (bit i: point i is the first of its voxel); rank[wd] = number of set bits
// before word wd, so the rank of the voxel that starts at point i is rank[i >> 5] + popc(bits below i).
// One CTA: a few thousand words.
__global__ void __launch_bounds__(1024) scan_kernel(int n, Ws w) {
  pdl_trigger();
  pdl_wait();
  __shared__ int warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nwords = (n + 31) / 32;
  const int per = (nwords + 1023) / 1024;
  const int lo = min(nwords, tid * per), hi = min(nwords, lo + per);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += __popc((unsigned)w.flag[i]);
  int inc = s;
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = warp_sum[lane];
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, d);
      if (lane >= d) ws += t;
    }
    warp_sum[lane] = ws;
  }
  __syncthreads();
  int run = inc - s + (wid ? warp_sum[wid - 1] : 0);
  for (int i = lo; i < hi; ++i) {
    w.rank[i] = run;
    run += __popc((unsigned)w.flag[i]);
  }
}

__global__ void __launch_bounds__(256) voxel_kernel(int n_points, const float* __restrict__ pts, Grid g, Ws w,
                                                     float* __restrict__ out, int32_t* __restrict__ num_voxels_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int nvox = w.counters[1];
  const int C = g.nz + g.n_meta;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_voxels_out) *num_voxels_out = min(nvox, g.max_voxels);
  for (int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < nvox; v += nw) {
    const int f0 = w.vfirst[v];
    const int vr = w.rank[f0 >> 5] + __popc((unsigned)w.flag[f0 >> 5] & ((1u << (f0 & 31)) - 1u));
    if (vr >= g.max_voxels) {                  // created past the cap: the reference never saw this voxel
      if (lane == 0) w.vrank[v] = -1;
      continue;
    }
    const int k = w.vkey[v], cnt = w.cnt[k];
    const int32_t* b = w.bucket + w.base[k];
    // More points than the voxel keeps: the first max_pts in input order are the max_pts smallest indices.
    // Bisect on the index value (indices are distinct): the smallest `cut` with max_pts indices below it.
    // O(cnt log n) per voxel, so a degenerate cloud (zero padding: every point in one voxel) stays cheap.
    int cut = 0x7fffffff;
    if (cnt > g.max_pts) {
      int lo = 0, hi = n_points;               // count(idx < lo) < max_pts <= count(idx < hi)
      while (hi - lo > 1) {
        const int mid = lo + ((hi - lo) >> 1);
        int below = 0;
        for (int j = lane; j < cnt; j += 32) below += b[j] < mid;
        below = (int)__reduce_add_sync(0xffffffffu, (unsigned)below);
        if (below >= g.max_pts) hi = mid; else lo = mid;
      }
      cut = hi;
    }
    float zmax = -INFINITY, si = 0.0f, se = 0.0f;
    for (int j = lane; j < cnt; j += 32) {
      const int idx = b[j];
      const bool take = idx < cut;
      if (take) {
        const float* p = pts + (size_t)idx * g.n_feat;
        zmax = fmaxf(zmax, fsub(p[2], g.z_lo));
        si += p[3];
        if (g.elongation) se += p[4];
      }
    }
    for (int d = 16; d; d >>= 1) {
      zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, d));
      si += __shfl_xor_sync(0xffffffffu, si, d);
      se += __shfl_xor_sync(0xffffffffu, se, d);
    }
    if (lane == 0) {
      const int npv = min(cnt, g.max_pts);
      if (npv < g.max_pts) zmax = fmaxf(zmax, 0.0f);       // unfilled slots of the voxel buffer are zeros (np.amax, :468)
      const int cz = k % g.nz, cy = (k / g.nz) % g.ny, cx = k / (g.nz * g.ny);
      // :468 float32 - int32 * python float -> float64, cast on assignment into the float32 map
      out[((size_t)cy * g.nx + cx) * C + cz] = (float)((double)zmax - (double)cz * (double)g.voxel_height);
      w.vrank[v] = vr;
      w.vval[3 * (size_t)v] = (float)((double)npv / (double)g.max_pts);                  // :483
      w.vval[3 * (size_t)v + 1] = (float)tanh((double)si / (double)npv);                  // :492-496 (float32 sum / int32 -> float64)
      w.vval[3 * (size_t)v + 2] = g.elongation ? (float)tanh((double)se / (double)npv) : 0.0f;   // :501-509
      atomicMax(&w.colmax[cx * g.ny + cy], vr);
    }
  }
}

__global__ void __launch_bounds__(256) meta_kernel(Grid g, Ws w, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int nvox = w.counters[1];
  const int C = g.nz + g.n_meta;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += gridDim.x * blockDim.x) {
    const int vr = w.vrank[v];
    if (vr < 0) continue;
    const int k = w.vkey[v];
    const int cy = (k / g.nz) % g.ny, cx = k / (g.nz * g.ny);
    if (w.colmax[cx * g.ny + cy] != vr) continue;          // a later voxel of this column overwrites it (:489,498,509)
    float* o = out + ((size_t)cy * g.nx + cx) * C + g.nz;
    for (int m = 0; m < g.n_meta; ++m) o[m] = w.vval[3 * (size_t)v + m];
  }
}

}  // namespace bev
}  // namespace b2d

using namespace b2d;

extern "C" size_t b2d_bev_workspace_bytes(int max_points, int nx, int ny, int nz) {
  if (max_points <= 0 || nx <= 0 || ny <= 0 || nz <= 0) return 0;
  return bev::carve(nullptr, max_points, nx, ny, nz).bytes;
}

extern "C" int b2d_bev_rasterize(int num_points, int num_feat, const float* points, float x_lo, float x_hi, float y_lo,
                                 float y_hi, float z_lo, float z_hi, float voxel_len, float voxel_height, int nx, int ny,
                                 int nz, int max_pts_per_voxel, int max_voxels, int num_meta, int elongation,
                                 float* bev_map, int32_t* num_voxels, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  if (num_points < 0 || num_feat < 4 || (elongation && num_feat < 5) || nx <= 0 || ny <= 0 || nz <= 0 ||
      max_pts_per_voxel <= 0 || max_voxels <= 0 || num_meta < 0 || num_meta > 3 || !bev_map || (num_points && !points) ||
      !(voxel_len > 0.0f) || !(voxel_height > 0.0f))
    return B2D_ERR_INVALID_ARG;
  if ((double)nx * ny * nz >= 2147483647.0) return B2D_ERR_UNSUPPORTED;
  cudaStream_t st = as_stream(stream);
  const size_t map_bytes = sizeof(float) * (size_t)nx * ny * (nz + num_meta);
  B2D_CUDA(cudaMemsetAsync(bev_map, 0, map_bytes, st));
  if (num_points == 0) {
    if (num_voxels) B2D_CUDA(cudaMemsetAsync(num_voxels, 0, sizeof(int32_t), st));
    return B2D_OK;
  }
  bev::Ws w = bev::carve(workspace, num_points, nx, ny, nz);
  if (!workspace || workspace_bytes < w.bytes) return B2D_ERR_WORKSPACE;
  B2D_CUDA(cudaMemsetAsync(w.cnt, 0, w.zero_bytes, st));
  B2D_CUDA(cudaMemsetAsync(w.colmax, 0xff, sizeof(int32_t) * (size_t)nx * ny, st));
  const bev::Grid g{x_lo, x_hi, y_lo, y_hi, z_lo, z_hi, voxel_len, voxel_height, nx, ny, nz,
                    max_pts_per_voxel, max_voxels, num_meta, elongation, num_feat};
  const int blocks = ceil_div(num_points, 256);
  // seven short kernels, chained by programmatic dependent launches (each waits for its predecessor on the device:
  // the launch gaps were a sixth of the call)
  bev::key_kernel<<<blocks, 256, 0, st>>>(num_points, points, g, w);          // (ordered after the memsets above)
  B2D_LAUNCHED();
  B2D_CUDA(launch_pdl(bev::alloc_kernel, dim3(blocks), dim3(256), 0, st, true, num_points, w));
  B2D_LAUNCHED();
  B2D_CUDA(launch_pdl(bev::fill_kernel, dim3(blocks), dim3(256), 0, st, true, num_points, w));
  B2D_LAUNCHED();
  const int wgrid = min(ceil_div(num_points, 8), 8 * kNumSMs);
  B2D_CUDA(launch_pdl(bev::first_kernel, dim3(wgrid), dim3(256), 0, st, true, w));
  B2D_LAUNCHED();
  B2D_CUDA(launch_pdl(bev::scan_kernel, dim3(1), dim3(1024), 0, st, true, num_points, w));
  B2D_LAUNCHED();
  B2D_CUDA(launch_pdl(bev::voxel_kernel, dim3(wgrid), dim3(256), 0, st, true, num_points, points, g, w, bev_map, num_voxels));
  B2D_LAUNCHED();
  B2D_CUDA(launch_pdl(bev::meta_kernel, dim3(min(blocks, 4 * kNumSMs)), dim3(256), 0, st, true, g, w, bev_map));
  B2D_LAUNCHED();
  return B2D_OK;
}

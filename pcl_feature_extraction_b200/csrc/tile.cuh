// tile.cuh — "cell tile" helpers.  One warp owns one occupied grid cell: the points of the cell's
// 3x3x3 stencil (its candidate set, M points) are staged ONCE in shared memory and every query point
// of the cell is processed against that copy.  All neighbourhood gathers of the dense stages then hit
// shared memory instead of L1/L2, and a neighbour is named by its 16-bit ordinal inside the candidate
// set ("local index": stencil lane order, then position inside the cell), which is identical for every
// kernel that tiles the same grid.
#pragma once
#include "common.cuh"

namespace pfx {

struct TileTab {  // per-warp table in shared memory
  int start[27];   // first sorted position of stencil cell l
  int prefix[28];  // exclusive prefix of the cell sizes; prefix[27] = M
};

// fills the table for `cell`; returns M (all lanes)
__device__ __forceinline__ int tile_setup(const GridDev& g, int cell, int lane, TileTab* tab) {
  int c = (lane < 27) ? g.cell_nbr[cell * 27 + lane] : -1;
  int start = 0, cnt = 0;
  if (c >= 0) {
    start = g.cell_start[c];
    cnt = g.cell_start[c + 1] - start;
  }
  int inc = warp_incl_scan(cnt, lane);
  if (lane < 27) {
    tab->start[lane] = start;
    tab->prefix[lane] = inc - cnt;
  }
  int M = __shfl_sync(FULL, inc, 31);
  if (lane == 27) tab->prefix[27] = M;
  __syncwarp();
  return M;
}

// sorted position of local index t (t < M)
__device__ __forceinline__ int tile_global_index(const TileTab* tab, int t) {
  int lo = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    int m = lo + step;
    if (m < 27 && tab->prefix[m] <= t) lo = m;
  }
  return tab->start[lo] + (t - tab->prefix[lo]);
}

// local index of the first point of the cell itself (stencil lane 13)
__device__ __forceinline__ int tile_own_offset(const TileTab* tab) { return tab->prefix[13]; }

// stage M rows of ROW_F4 float4 each from a sorted global array into shared memory
template <int ROW_F4>
__device__ __forceinline__ void tile_stage(const float4* __restrict__ src, const TileTab* tab, int M, int lane,
                                           float4* __restrict__ dst) {
  const int total = M * ROW_F4;
  for (int e = lane; e < total; e += 32) {
    int t = e / ROW_F4, part = e - t * ROW_F4;
    int j = tile_global_index(tab, t);
    dst[e] = src[(size_t)j * ROW_F4 + part];
  }
  __syncwarp();
}

}  // namespace pfx

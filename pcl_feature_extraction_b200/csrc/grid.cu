// grid.cu — GPU voxel hash built with sort-and-scan (replaces the kd-tree / FLANN index that the
// reference builds at features.h:192-193, tools.h:29-30, keypoints.h:186-187,371-372,408-409).
//
// build (9 launches): bbox -> [occupied-cell probe at a trial edge, k-search grids] -> 30-bit Morton keys + the digit
// histograms of all four radix passes -> 4 x one-sweep LSD radix pass (8-bit digits, stable, chained scan with
// decoupled look-back: one read and one write of the pairs per pass) -> ONE kernel that finds the cell heads, scans
// them (decoupled look-back again), writes the cell table and the per-point cell ids, inserts the cells into the
// open-addressing hash and gathers the points into sorted order -> 27-neighbour adjacency table.  Nothing
// synchronises with the host: the grid description (GridParams) lives in device memory and every kernel that depends
// on a data-dependent count uses a grid-stride loop over a device-side bound.
#include "internal.h"

namespace pfx {

// ------------------------------------------------------------------------------------------ scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* smem /* >= 8 */, T* block_total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    T w = (lane < (SCAN_THREADS / 32)) ? smem[lane] : T(0);
    T winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      T t = __shfl_up_sync(FULL, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < 8) smem[lane] = winc - w;
    if (lane == 7) smem[8] = winc;
  }
  __syncthreads();
  T res = smem[wid] + inc - v;
  if (block_total) *block_total = smem[8];
  __syncthreads();
  return res;
}

template <typename T>
__global__ void scan_reduce_kernel(const int* __restrict__ in, int n, T* __restrict__ bsum) {
  __shared__ T sm[9];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  T s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  T tot;
  block_excl_scan<T>(s, sm, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

template <typename T>
__global__ void scan_bsum_kernel(T* bsum, int nb, T* total_out) {
  __shared__ T sm[9];
  __shared__ T carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    T v = (i < nb) ? bsum[i] : T(0);
    T tot;
    T ex = block_excl_scan<T>(v, sm, &tot);
    if (i < nb) bsum[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T>
__global__ void scan_apply_kernel(const int* __restrict__ in, T* __restrict__ out, int n,
                                  const T* __restrict__ bsum, int write_total_at_n) {
  __shared__ T sm[9];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  T s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  T ex = block_excl_scan<T>(s, sm, nullptr) + bsum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
    if (write_total_at_n && base + i == n - 1) out[n] = ex;
  }
}

template <typename T>
static int scan_exclusive_impl(Ctx* ctx, const int* in, T* out, int n, T* total_dev, DevBuf& bsum,
                               int write_total_at_n) {
  if (n <= 0) {
    if (write_total_at_n) PFX_CUDA(cudaMemsetAsync(out, 0, sizeof(T), ctx->stream));
    if (total_dev) PFX_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(T), ctx->stream));
    return 0;
  }
  int nb = div_up(n, SCAN_TILE);
  PFX_CUDA(bsum.ensure((size_t)(nb + 1) * sizeof(T)));
  PFX_LAUNCH(ctx, scan_reduce_kernel<T>, nb, SCAN_THREADS, 0, in, n, bsum.as<T>());
  PFX_LAUNCH(ctx, scan_bsum_kernel<T>, 1, SCAN_THREADS, 0, bsum.as<T>(), nb, total_dev);
  PFX_LAUNCH(ctx, scan_apply_kernel<T>, nb, SCAN_THREADS, 0, in, out, n, bsum.as<T>(),
             write_total_at_n);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

int scan_exclusive_i32(Ctx* ctx, const int* in, int* out, int n, int* total_dev, DevBuf& bsum) {
  return scan_exclusive_impl<int>(ctx, in, out, n, total_dev, bsum, 0);
}
// out has n + 1 entries (CSR offsets)
int scan_exclusive_i64(Ctx* ctx, const int* in, long long* out, int n, DevBuf& bsum) {
  return scan_exclusive_impl<long long>(ctx, in, out, n, nullptr, bsum, 1);
}

// ------------------------------------------------------------------------------------------ bbox
struct BuildAcc {  // all-zero initial state (one memset): the minima are kept as the complement of their ordered key
  uint32_t mn_inv[3], mx[3];
  int n_valid;
  int n_cells_probe;  // occupied cells counted at the trial edge (kNN density estimate)
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void bbox_kernel(const float4* __restrict__ pts, int n, BuildAcc* acc) {
  uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  int cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    if (finite3(p.x, p.y, p.z)) {
      uint32_t ox = f2ord(p.x), oy = f2ord(p.y), oz = f2ord(p.z);
      mn[0] = min(mn[0], ox); mx[0] = max(mx[0], ox);
      mn[1] = min(mn[1], oy); mx[1] = max(mx[1], oy);
      mn[2] = min(mn[2], oz); mx[2] = max(mx[2], oz);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = min(mn[a], __shfl_xor_sync(FULL, mn[a], o));
      mx[a] = max(mx[a], __shfl_xor_sync(FULL, mx[a], o));
    }
    cnt += __shfl_xor_sync(FULL, cnt, o);
  }
  __shared__ uint32_t s_mn[3], s_mx[3];
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      s_mn[a] = 0xffffffffu;
      s_mx[a] = 0u;
    }
    s_cnt = 0;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    for (int a = 0; a < 3; ++a) {
      atomicMin(&s_mn[a], mn[a]);
      atomicMax(&s_mx[a], mx[a]);
    }
    atomicAdd(&s_cnt, cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      atomicMax(&acc->mn_inv[a], ~s_mn[a]);
      atomicMax(&acc->mx[a], s_mx[a]);
    }
    atomicAdd(&acc->n_valid, s_cnt);
  }
}

// The cell edge and the grid dimensions.  edge_req > 0: radius grid.  edge_req <= 0: k-search grid, stage 0 picks a
// trial edge from the bbox (2-manifold guess), stage 1 rescales it (prev_edge) so that an occupied cell holds about
// `target_occ` points (occupancy measured by cell_probe_kernel).  A pure function of the accumulator: every kernel
// that needs the parameters before they are stored recomputes them.
__device__ __forceinline__ GridParams make_params(const BuildAcc* acc, float edge_req, int stage, float target_occ,
                                                   float prev_edge) {
  GridParams P;
  int nv = acc->n_valid;
  P.n_valid = nv;
  P.ncells = 0;
  if (nv == 0) {
    P.mnx = P.mny = P.mnz = P.mxx = P.mxy = P.mxz = 0.f;
  } else {
    P.mnx = ord2f(~acc->mn_inv[0]); P.mny = ord2f(~acc->mn_inv[1]); P.mnz = ord2f(~acc->mn_inv[2]);
    P.mxx = ord2f(acc->mx[0]); P.mxy = ord2f(acc->mx[1]); P.mxz = ord2f(acc->mx[2]);
  }
  float ex = P.mxx - P.mnx, ey = P.mxy - P.mny, ez = P.mxz - P.mnz;
  float edge = edge_req;
  if (!(edge_req > 0.f)) {
    if (stage == 0) {
      float a = fmaxf(ex, fmaxf(ey, ez)), c = fminf(ex, fminf(ey, ez));
      float b = ex + ey + ez - a - c;
      float area = fmaxf(a * b, 1e-20f);
      edge = sqrtf(area * target_occ / fmaxf((float)nv, 1.f));
    } else {
      float occ = (float)nv / fmaxf((float)acc->n_cells_probe, 1.f);
      edge = prev_edge * sqrtf(target_occ / fmaxf(occ, 1e-3f));
    }
    edge = fmaxf(edge, 1e-7f * fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f)));
    if (!(edge > 0.f) || !isfinite(edge)) edge = 1.f;
  }
  for (int it = 0; it < 64; ++it) {  // keep every axis within 10 Morton bits
    float m = fmaxf(ex, fmaxf(ey, ez)) / edge;
    if (m < 1020.f) break;
    edge *= 2.f;
  }
  P.edge = edge;
  P.inv_e = 1.0f / edge;
  P.ox = P.mnx; P.oy = P.mny; P.oz = P.mnz;
  P.nx = min(1024, (int)floorf(ex * P.inv_e) + 1);
  P.ny = min(1024, (int)floorf(ey * P.inv_e) + 1);
  P.nz = min(1024, (int)floorf(ez * P.inv_e) + 1);
  return P;
}
// the final parameters of a grid: radius grid (edge_req > 0) or the rescaled k-search grid
__device__ __forceinline__ GridParams final_params(const BuildAcc* acc, float edge_req, float target_occ) {
  if (edge_req > 0.f) return make_params(acc, edge_req, 0, 0.f, 0.f);
  const GridParams trial = make_params(acc, 0.f, 0, target_occ, 0.f);
  return make_params(acc, 0.f, 1, target_occ, trial.edge);
}

__global__ void empty_params_kernel(const BuildAcc* acc, GridParams* gp, float edge_req, float target_occ) {
  *gp = final_params(acc, edge_req, target_occ);  // an empty cloud: zero counts, unit dimensions
}

__device__ __forceinline__ uint32_t point_key(const GridParams& P, float x, float y, float z) {
  if (!finite3(x, y, z)) return KEY_INVALID;
  return morton3(cell_coord(x, P.ox, P.inv_e, P.nx), cell_coord(y, P.oy, P.inv_e, P.ny),
                 cell_coord(z, P.oz, P.inv_e, P.nz));
}

// count distinct occupied cells at the trial edge with a hash set (kNN density estimate)
__global__ void cell_probe_kernel(const float4* __restrict__ pts, int n, float target_occ,
                                  uint32_t* hset, uint32_t hmask, BuildAcc* acc) {
  __shared__ GridParams sP;
  if (threadIdx.x == 0) sP = make_params(acc, 0.f, 0, target_occ, 0.f);  // (n_cells_probe is not read at stage 0)
  __syncthreads();
  const GridParams P = sP;
  int local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    uint32_t key = point_key(P, p.x, p.y, p.z);
    if (key == KEY_INVALID) continue;
    uint32_t h = (hash_key(key) >> 7) & hmask;
    for (;;) {
      uint32_t prev = atomicCAS(&hset[h], KEY_INVALID, key);
      if (prev == KEY_INVALID) { ++local; break; }
      if (prev == key) break;
      h = (h + 1) & hmask;
    }
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&acc->n_cells_probe, local);
}

// digit histograms of the four radix passes, accumulated per block in shared memory
__device__ __forceinline__ void hist4_add(unsigned (*h)[256], uint32_t key) {
  atomicAdd(&h[0][key & 255u], 1u);
  atomicAdd(&h[1][(key >> 8) & 255u], 1u);
  atomicAdd(&h[2][(key >> 16) & 255u], 1u);
  atomicAdd(&h[3][key >> 24], 1u);
}
__device__ __forceinline__ void hist4_flush(unsigned (*h)[256], unsigned* __restrict__ ghist) {
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const unsigned c = (&h[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// Morton keys of the points under the grid's final parameters (computed here from the accumulator and stored by
// block 0), the identity permutation, and the digit histograms of the sort that follows
__global__ void __launch_bounds__(256)
keys_hist_kernel(const float4* __restrict__ pts, int n, const BuildAcc* acc, float edge_req, float target_occ,
                 GridParams* gp, uint32_t* __restrict__ keys, int* __restrict__ vals, unsigned* __restrict__ ghist) {
  __shared__ GridParams sP;
  __shared__ unsigned h[4][256];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&h[0][0])[i] = 0u;
  if (threadIdx.x == 0) {
    sP = final_params(acc, edge_req, target_occ);
    if (blockIdx.x == 0) *gp = sP;
  }
  __syncthreads();
  const GridParams P = sP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pts[i];
    const uint32_t key = point_key(P, p.x, p.y, p.z);
    keys[i] = key;
    vals[i] = i;
    hist4_add(h, key);
  }
  hist4_flush(h, ghist);
}

// the histograms alone, for keys made elsewhere (VoxelGrid ids, slab global ids)
__global__ void __launch_bounds__(256)
rs_hist4_kernel(const uint32_t* __restrict__ keys, int n, unsigned* __restrict__ ghist) {
  __shared__ unsigned h[4][256];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) (&h[0][0])[i] = 0u;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) hist4_add(h, keys[i]);
  hist4_flush(h, ghist);
}

// ------------------------------------------------------------------------------------ radix sort
// LSD, 8-bit digits, stable, ONE-SWEEP: the digit histograms of all four passes are taken up front (with the keys),
// so a pass reads its tile of pairs once, ranks them, learns where its digits start in the output from the tiles
// before it through a chained scan with decoupled look-back (a status word per (tile, digit): 2 flag bits + 30-bit
// count; tiles are handed out by an atomic ticket, so every predecessor of a running tile is itself running), and
// writes the pairs once.  Tile = 8 warps x 8 rounds x 32 lanes = 2048 pairs; warp w owns the contiguous chunk
// [w*256, w*256+256) of the tile so that (tile, warp, round, lane) is the input order.  Ranking inside a warp uses
// __match_any_sync (no shared-memory atomics on the hot path); the tile is staged in shared memory in sorted order so
// that the global writes of one digit are contiguous.
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
constexpr unsigned LB_AGG = 1u << 30, LB_INCL = 2u << 30, LB_FLAGS = 3u << 30, LB_VALUE = (1u << 30) - 1u;

__device__ __forceinline__ unsigned ld_status(const unsigned* p) {
  return *reinterpret_cast<const volatile unsigned*>(p);
}
__device__ __forceinline__ void st_status(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

// layout of the sort's control block (zeroed by one memset per sort): tickets, histograms, status words
struct SortCtl {
  unsigned ticket[4];
  unsigned pad[60];
  unsigned ghist[4][256];
};
static size_t sort_ctl_bytes(int ntiles) { return sizeof(SortCtl) + (size_t)4 * ntiles * 256 * sizeof(unsigned); }

__global__ void __launch_bounds__(RS_THREADS)
rs_onesweep_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ vals, int n, int pass, int morton_keys,
                   SortCtl* __restrict__ ctl, unsigned* __restrict__ status /* [ntiles][256] of this pass */,
                   uint32_t* __restrict__ okeys, int* __restrict__ ovals) {
  __shared__ int wcnt[8][256];
  __shared__ uint32_t skey[RS_TILE];
  __shared__ int sval[RS_TILE];
  __shared__ unsigned sdst[256];
  __shared__ unsigned scan_sm[9];
  __shared__ int s_tile;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctl->ticket[pass], 1u);
  for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
  // A pass whose digit is the same for every key moves nothing: the tile is copied across and that is all.  So does
  // the top pass of a voxel hash (morton_keys: 30-bit Morton keys, 0xffffffff for non-finite points) whose cells need
  // fewer than 24 key bits (any grid of up to 256 cells per axis): the keys of the finite points share the digit 0,
  // the invalid key is alone in digit 255, and the three passes before have already put the invalid keys last - as
  // long as no valid key ends in 0xffffff, which the histogram of the third digit tells (digit 255 there belongs to
  // the invalid keys alone).
  bool ident;
  {
    const unsigned gh = ctl->ghist[pass][threadIdx.x];
    ident = __syncthreads_or(gh == (unsigned)n) != 0;
    if (!ident && pass == 3 && morton_keys)
      ident = ctl->ghist[3][0] + ctl->ghist[3][255] == (unsigned)n && ctl->ghist[2][255] == ctl->ghist[3][255];
  }
  if (ident) {
    const int t0 = blockIdx.x * RS_TILE;
    for (int i = t0 + threadIdx.x; i < min(n, t0 + RS_TILE); i += RS_THREADS) {
      okeys[i] = keys[i];
      ovals[i] = vals[i];
    }
    return;
  }
  __syncthreads();
  const int tile = s_tile;
  const int shift = pass * 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int tbase = tile * RS_TILE;
  const int base = tbase + wid * (RS_ROUNDS * 32);
  uint32_t k[RS_ROUNDS];
  int v[RS_ROUNDS], rk[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    int i = base + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? keys[i] : 0u;
    v[r] = ok ? vals[i] : 0;
  }
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    bool ok = base + r * 32 + lane < n;
    int d = ok ? (int)((k[r] >> shift) & 255u) : 256;  // 256: padding lanes group together
    unsigned m = __match_any_sync(FULL, d);
    int old = ok ? wcnt[wid][d] : 0;
    __syncwarp();
    rk[r] = old + __popc(m & lt);
    if (ok && (m & lt) == 0) wcnt[wid][d] = old + __popc(m);
    __syncwarp();
  }
  __syncthreads();
  // thread d owns digit d: counts of the 8 warps -> exclusive prefix, tile total
  const int d = threadIdx.x;
  int run = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    int t = wcnt[w][d];
    wcnt[w][d] = run;
    run += t;
  }
  unsigned* st = status + (size_t)tile * 256 + d;
  st_status(st, (tile == 0 ? LB_INCL : LB_AGG) | (unsigned)run);
  // where digit d starts in the output (all tiles) and inside this tile
  const unsigned gbase = block_excl_scan<unsigned>(ctl->ghist[pass][d], scan_sm, nullptr);
  const unsigned lstart = block_excl_scan<unsigned>((unsigned)run, scan_sm, nullptr);
  unsigned excl = 0;
  if (tile > 0) {
    for (int t = tile - 1;; --t) {
      unsigned sv;
      do {
        sv = ld_status(status + (size_t)t * 256 + d);
      } while ((sv & LB_FLAGS) == 0u);
      excl += sv & LB_VALUE;
      if (sv & LB_INCL) break;
    }
    st_status(st, LB_INCL | (excl + (unsigned)run));
  }
  sdst[d] = gbase + excl - lstart;  // + position inside the sorted tile = position in the output
#pragma unroll
  for (int w = 0; w < 8; ++w) wcnt[w][d] += (int)lstart;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    if (base + r * 32 + lane < n) {
      const int pos = wcnt[wid][(k[r] >> shift) & 255u] + rk[r];
      skey[pos] = k[r];
      sval[pos] = v[r];
    }
  }
  __syncthreads();
  const int tile_n = min(RS_TILE, n - tbase);
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const int i = r * RS_THREADS + threadIdx.x;
    if (i < tile_n) {
      const uint32_t key = skey[i];
      const unsigned dst = sdst[(key >> shift) & 255u] + (unsigned)i;
      okeys[dst] = key;
      ovals[dst] = sval[i];
    }
  }
}

// sorts the pairs in g->keys / g->vals (result back in the same buffers).  have_hist: the control block was zeroed and
// the histograms were taken by the kernel that made the keys.
static int sort_ctl_prepare(Ctx* ctx, Grid* g, int n) {
  const int ntiles = div_up(n, RS_TILE);
  PFX_CUDA(g->ghist.ensure(sort_ctl_bytes(ntiles)));
  PFX_CUDA(cudaMemsetAsync(g->ghist.p, 0, sort_ctl_bytes(ntiles), ctx->stream));
  return 0;
}

static int radix_sort_pairs(Ctx* ctx, Grid* g, int n, bool have_hist = false) {
  const int ntiles = div_up(n, RS_TILE);
  uint32_t* k0 = g->keys.as<uint32_t>();
  uint32_t* k1 = g->keys2.as<uint32_t>();
  int* v0 = g->vals.as<int>();
  int* v1 = g->vals2.as<int>();
  if (!have_hist) {
    PFX_TRY(sort_ctl_prepare(ctx, g, n));
    PFX_LAUNCH(ctx, rs_hist4_kernel, std::min(ctx->sm_count * 4, div_up(n, 256)), 256, 0, k0, n,
               &g->ghist.as<SortCtl>()->ghist[0][0]);
  }
  SortCtl* ctl = g->ghist.as<SortCtl>();
  unsigned* status = reinterpret_cast<unsigned*>(ctl + 1);
  for (int pass = 0; pass < 4; ++pass) {
    PFX_LAUNCH(ctx, rs_onesweep_kernel, ntiles, RS_THREADS, 0, k0, v0, n, pass, have_hist ? 1 : 0, ctl,
               status + (size_t)pass * ntiles * 256, k1, v1);
    std::swap(k0, k1);
    std::swap(v0, v1);
  }
  PFX_CUDA(cudaGetLastError());
  return 0;  // 4 passes: result is back in keys / vals
}

// ascending sort of n (uint32 key, int value) pairs held in the scratch buffers of ctx->vg_scratch (keys / vals):
// used by the slab distribution (group.cu) to order the local points by global id
int sort_pairs_scratch(Ctx* ctx, int n, uint32_t** keys, int** vals) {
  Grid* g = &ctx->vg_scratch;
  const size_t nn = (size_t)std::max(n, 1);
  PFX_CUDA(g->keys.ensure(nn * sizeof(uint32_t)));
  PFX_CUDA(g->keys2.ensure(nn * sizeof(uint32_t)));
  PFX_CUDA(g->vals.ensure(nn * sizeof(int)));
  PFX_CUDA(g->vals2.ensure(nn * sizeof(int)));
  *keys = g->keys.as<uint32_t>();
  *vals = g->vals.as<int>();
  return 0;
}
int sort_pairs_scratch_run(Ctx* ctx, int n) { return n > 1 ? radix_sort_pairs(ctx, &ctx->vg_scratch, n) : 0; }

// ------------------------------------------------------------------------------- post-sort stages
__global__ void heads_kernel(const uint32_t* __restrict__ keys, int n, int* __restrict__ heads) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  heads[i] = (k != KEY_INVALID && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// Everything between the sort and the adjacency table in ONE pass over the sorted pairs: cell heads (a key that
// differs from its predecessor), their exclusive scan across tiles (single-value chained scan with decoupled look-back,
// tiles handed out by an atomic ticket), the cell table (cell_key / cell_start), the cell id of every point, the
// insertion of each new cell into the open-addressing hash, and the gather of the points into sorted order
// (w = original index) with the inverse permutation.  The tile that holds the last pair also stores the cell count
// and the end sentinel of cell_start.
constexpr int CF_THREADS = 256;
constexpr int CF_ITEMS = 8;
constexpr int CF_TILE = CF_THREADS * CF_ITEMS;

struct CellsCtl {
  unsigned ticket;
  unsigned pad[15];
};

__global__ void __launch_bounds__(CF_THREADS)
cells_fused_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ vals, const float4* __restrict__ pts,
                   int n, CellsCtl* __restrict__ ctl, unsigned* __restrict__ status /* [ntiles] */,
                   float4* __restrict__ sorted, int* __restrict__ inv_perm, int* __restrict__ pt_cell,
                   uint32_t* __restrict__ cell_key, int* __restrict__ cell_start, uint32_t* __restrict__ hkeys,
                   int* __restrict__ hvals, uint32_t hmask, GridParams* gp) {
  __shared__ int sm[9];
  __shared__ int s_tile, s_excl;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctl->ticket, 1u);
  __syncthreads();
  const int tile = s_tile;
  const int base = tile * CF_TILE + threadIdx.x * CF_ITEMS;
  uint32_t k[CF_ITEMS];
  uint32_t prev = KEY_INVALID;  // (the first pair of the cloud is a head whenever its key is valid)
  if (base > 0 && base < n) prev = keys[base - 1];
  int heads = 0;
  unsigned hmask_bits = 0;
#pragma unroll
  for (int i = 0; i < CF_ITEMS; ++i) {
    k[i] = (base + i < n) ? keys[base + i] : KEY_INVALID;
    const bool h = k[i] != KEY_INVALID && (base + i == 0 || k[i] != prev);
    prev = k[i];
    hmask_bits |= (h ? 1u : 0u) << i;
    heads += h ? 1 : 0;
  }
  int tile_total;
  const int thread_excl = block_excl_scan<int>(heads, sm, &tile_total);
  // exclusive prefix of this tile: warp 0 looks back over its predecessors, 32 at a time
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane == 0) st_status(status + tile, (tile == 0 ? LB_INCL : LB_AGG) | (unsigned)tile_total);
    unsigned excl = 0;
    for (int t0 = tile - 1; t0 >= 0; t0 -= 32) {
      const int t = t0 - lane;
      unsigned sv = LB_INCL;  // before tile 0: an inclusive prefix of zero
      if (t >= 0) {
        do {
          sv = ld_status(status + t);
        } while ((sv & LB_FLAGS) == 0u);
      }
      const unsigned incl = __ballot_sync(FULL, (sv & LB_INCL) != 0u);
      const int first = incl ? __ffs(incl) - 1 : 31;  // nearest predecessor that holds an inclusive prefix
      unsigned add = (lane <= first) ? (sv & LB_VALUE) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(FULL, add, o);
      excl += add;
      if (incl) break;
    }
    if (lane == 0) {
      if (tile > 0) st_status(status + tile, LB_INCL | (excl + (unsigned)tile_total));
      s_excl = (int)excl;
    }
  }
  __syncthreads();
  int ex = s_excl + thread_excl;  // cells that start before my first pair
#pragma unroll
  for (int i = 0; i < CF_ITEMS; ++i) {
    const int idx = base + i;
    if (idx >= n) break;
    const bool h = (hmask_bits >> i) & 1u;
    if (h) {
      cell_key[ex] = k[i];
      cell_start[ex] = idx;
      uint32_t slot = (hash_key(k[i]) >> 7) & hmask;
      for (;;) {
        const uint32_t was = atomicCAS(&hkeys[slot], KEY_INVALID, k[i]);
        if (was == KEY_INVALID) {
          hvals[slot] = ex;
          break;
        }
        slot = (slot + 1) & hmask;
      }
      ++ex;
    }
    pt_cell[idx] = (k[i] == KEY_INVALID) ? -1 : ex - 1;
    const int o = vals[idx];
    float4 p = pts[o];
    p.w = __int_as_float(o);
    sorted[idx] = p;
    inv_perm[o] = idx;
  }
  if (threadIdx.x == CF_THREADS - 1 && (tile + 1) * CF_TILE >= n) {  // the tile with the last pair: ex = all cells
    gp->ncells = ex;
    cell_start[ex] = gp->n_valid;
  }
}

__device__ __forceinline__ uint32_t compact1by2(uint32_t x) {
  x &= 0x09249249u;
  x = (x | (x >> 2)) & 0x030c30c3u;
  x = (x | (x >> 4)) & 0x0300f00fu;
  x = (x | (x >> 8)) & 0x030000ffu;
  x = (x | (x >> 16)) & 0x3ffu;
  return x;
}

__global__ void adjacency_kernel(GridDev g, int* __restrict__ cell_nbr) {
  const GridParams P = *g.gp;
  long long total = (long long)P.ncells * 27;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int c = (int)(t / 27), l = (int)(t % 27);
    uint32_t key = g.cell_key[c];
    int cx = (int)compact1by2(key), cy = (int)compact1by2(key >> 1), cz = (int)compact1by2(key >> 2);
    int x2 = cx + l % 3 - 1, y2 = cy + (l / 3) % 3 - 1, z2 = cz + l / 9 - 1;
    int r = -1;
    if (l == 13)
      r = c;
    else if (x2 >= 0 && x2 < P.nx && y2 >= 0 && y2 < P.ny && z2 >= 0 && z2 < P.nz)
      r = hash_lookup(g, morton3(x2, y2, z2));
    cell_nbr[t] = r;
  }
}

static uint32_t pow2_at_least(size_t v) {
  uint32_t p = 1024;
  while (p < v) p <<= 1;
  return p;
}

static int grid_build(Ctx* ctx, Grid* g, double radius, int knn_k) {
  const int n = (int)ctx->n;
  g->n = n;
  g->radius = radius;
  g->knn_k = knn_k;
  g->surf_version = ctx->surf_version;
  const size_t nn = (size_t)std::max(n, 1);
  PFX_CUDA(g->params.ensure(sizeof(GridParams)));
  PFX_CUDA(g->misc.ensure(sizeof(BuildAcc) + 64));
  PFX_CUDA(g->pts.ensure(nn * sizeof(float4)));
  PFX_CUDA(g->inv_perm.ensure(nn * sizeof(int)));
  PFX_CUDA(g->keys.ensure(nn * sizeof(uint32_t)));
  PFX_CUDA(g->keys2.ensure(nn * sizeof(uint32_t)));
  PFX_CUDA(g->vals.ensure(nn * sizeof(int)));
  PFX_CUDA(g->vals2.ensure(nn * sizeof(int)));
  PFX_CUDA(g->pt_cell.ensure(nn * sizeof(int)));
  PFX_CUDA(g->cell_key.ensure(nn * sizeof(uint32_t)));
  PFX_CUDA(g->cell_start.ensure((nn + 1) * sizeof(int)));
  PFX_CUDA(g->cell_nbr.ensure(nn * 27 * sizeof(int)));
  g->hmask = pow2_at_least(2 * nn) - 1;
  PFX_CUDA(g->hkeys.ensure(((size_t)g->hmask + 1) * sizeof(uint32_t)));
  PFX_CUDA(g->hvals.ensure(((size_t)g->hmask + 1) * sizeof(int)));

  BuildAcc* acc = g->misc.as<BuildAcc>();
  GridParams* gp = g->params.as<GridParams>();
  const float4* src = ctx->surf.as<float4>();
  const int T = 256;
  const int wide = ctx->sm_count * 4;
  const float edge_req = radius > 0 ? (float)(radius * (1.0 + 1e-3)) : 0.f;
  // points per occupied cell of a k-search grid: ~k/3 keeps the 3x3x3 stencil of a surface at 100-200 candidates
  // while the k-th neighbour still falls inside it (cell-tile kNN, knn_tile.cu)
  const float target = radius > 0 ? 0.f : std::max(2.0f, ctx->knn_occupancy * (float)knn_k);
  PFX_CUDA(cudaMemsetAsync(acc, 0, sizeof(BuildAcc) + 64, ctx->stream));
  if (n > 0) {
    PFX_LAUNCH(ctx, bbox_kernel, std::min(wide, div_up(n, T)), T, 0, src, n, acc);
    if (!(radius > 0)) {
      PFX_CUDA(cudaMemsetAsync(g->hkeys.p, 0xff, ((size_t)g->hmask + 1) * sizeof(uint32_t), ctx->stream));
      PFX_LAUNCH(ctx, cell_probe_kernel, std::min(wide, div_up(n, T)), T, 0, src, n, target, g->hkeys.as<uint32_t>(),
                 g->hmask, acc);
    }
    PFX_TRY(sort_ctl_prepare(ctx, g, n));
    PFX_LAUNCH(ctx, keys_hist_kernel, std::min(wide, div_up(n, T)), T, 0, src, n, acc, edge_req, target, gp,
               g->keys.as<uint32_t>(), g->vals.as<int>(), &g->ghist.as<SortCtl>()->ghist[0][0]);
    PFX_TRY(radix_sort_pairs(ctx, g, n, true));
    const int ntiles = div_up(n, CF_TILE);
    PFX_CUDA(g->bsum.ensure(sizeof(CellsCtl) + (size_t)ntiles * sizeof(unsigned)));
    PFX_CUDA(cudaMemsetAsync(g->bsum.p, 0, sizeof(CellsCtl) + (size_t)ntiles * sizeof(unsigned), ctx->stream));
    PFX_CUDA(cudaMemsetAsync(g->hkeys.p, 0xff, ((size_t)g->hmask + 1) * sizeof(uint32_t), ctx->stream));
    CellsCtl* cctl = g->bsum.as<CellsCtl>();
    PFX_LAUNCH(ctx, cells_fused_kernel, ntiles, CF_THREADS, 0, g->keys.as<uint32_t>(), g->vals.as<int>(), src, n, cctl,
               reinterpret_cast<unsigned*>(cctl + 1), g->pts.as<float4>(), g->inv_perm.as<int>(), g->pt_cell.as<int>(),
               g->cell_key.as<uint32_t>(), g->cell_start.as<int>(), g->hkeys.as<uint32_t>(), g->hvals.as<int>(), g->hmask,
               gp);
    PFX_LAUNCH(ctx, adjacency_kernel, wide * 2, T, 0, g->view(), g->cell_nbr.as<int>());
  } else {
    PFX_LAUNCH(ctx, empty_params_kernel, 1, 1, 0, acc, gp, edge_req, target);
    PFX_CUDA(cudaMemsetAsync(g->cell_start.p, 0, sizeof(int), ctx->stream));
  }
  // the parameters the device settled on (cell edge, dimensions) come back asynchronously: a later radius stage may
  // reuse this hash once its edge is known to cover the radius (grid_for_radius; never waited for)
  if (!g->host_params) {
    PFX_CUDA(cudaMallocHost(&g->host_params, sizeof(GridParams)));
    PFX_CUDA(cudaEventCreateWithFlags(&g->built, cudaEventDisableTiming));
  }
  PFX_CUDA(cudaMemcpyAsync(g->host_params, gp, sizeof(GridParams), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaEventRecord(g->built, ctx->stream));
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// Grids are cached per (surface version, radius | k): repeated compute() calls on the same cloud do
// not rebuild the index (the reference rebuilds a kd-tree per Feature object, features.h:192-193).
static int grid_consume_pending(Ctx* ctx, Grid* g) {
  if (g->pending) {
    PFX_CUDA(cudaStreamWaitEvent(ctx->stream, g->ready, 0));
    g->pending = false;
  }
  return 0;
}

int grid_wait_pending(Ctx* ctx) {
  for (Grid* g : ctx->grids) PFX_TRY(grid_consume_pending(ctx, g));
  return 0;
}

// Builds the grid of `radius` on the auxiliary stream, behind everything the main stream has enqueued so far
// (the surface upload), so that the build overlaps whatever the main stream does next.  The Grid's buffers are its
// own and the surface is only read, so the two streams share nothing else.
int grid_prepare_async(Ctx* ctx, double radius) {
  for (Grid* g : ctx->grids)
    if (g->surf_version == ctx->surf_version && g->radius == radius) return 0;  // cached or already in flight
  if (!ctx->aux_stream) {
    PFX_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    PFX_CUDA(cudaEventCreateWithFlags(&ctx->ev_surface, cudaEventDisableTiming));
  }
  PFX_CUDA(cudaEventRecord(ctx->ev_surface, ctx->stream));
  PFX_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_surface, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->aux_stream;
  Grid* g = nullptr;
  int rc = grid_get(ctx, radius, 0, &g);
  if (rc == 0) {
    if (!g->ready && cudaEventCreateWithFlags(&g->ready, cudaEventDisableTiming) != cudaSuccess) rc = PFX_E_STATE;
    if (rc == 0 && cudaEventRecord(g->ready, ctx->aux_stream) != cudaSuccess) rc = PFX_E_STATE;
    if (rc == 0) g->pending = true;
  }
  ctx->stream = main_stream;
  if (rc != 0 && ctx->err.empty()) ctx->err = "pfx_prepare_radius: could not enqueue the build";
  return rc;
}

// A hash for a radius search: the grid built for exactly this radius when it exists, else any grid of the current
// surface whose cell edge is KNOWN to cover the radius (3x3x3 stencils need edge >= radius (1 + 1e-3)) without
// being more than 1.6x too coarse, else a new one.  The edge of a radius grid is known on the host; the edge of a
// k-search grid is chosen on the device and read back asynchronously - it counts only when that read-back has
// already landed (cudaEventQuery: the host never waits here, so a pipeline that runs ahead of the device simply
// builds the radius grid, on the auxiliary stream when pfx_prepare_radius announced it).
int grid_for_radius(Ctx* ctx, double radius, Grid** out) {
  for (Grid* g : ctx->grids)
    if (g->surf_version == ctx->surf_version && g->radius == radius) return grid_get(ctx, radius, 0, out);
  const double need = radius * (1.0 + 1e-3);
  for (Grid* g : ctx->grids) {
    if (g->surf_version != ctx->surf_version || !g->host_params || g->n != (int)ctx->n || g->pending) continue;
    double edge = 0;
    if (g->radius > 0) edge = (double)(float)(g->radius * (1.0 + 1e-3));
    else if (cudaEventQuery(g->built) == cudaSuccess) edge = (double)static_cast<const GridParams*>(g->host_params)->edge;
    else (void)cudaGetLastError();  // cudaErrorNotReady is not an error
    if (edge >= need && edge <= 1.6 * need) {
      ctx->tick++;
      g->last_use = ctx->tick;
      ctx->last_grid = g;
      *out = g;
      return 0;
    }
  }
  return grid_get(ctx, radius, 0, out);
}

int grid_get(Ctx* ctx, double radius, int knn_k, Grid** out) {
  ctx->tick++;
  Grid* victim = nullptr;
  for (Grid* g : ctx->grids) {
    if (g->surf_version == ctx->surf_version &&
        ((radius > 0 && g->radius == radius) || (!(radius > 0) && g->knn_k == knn_k && !(g->radius > 0)))) {
      PFX_TRY(grid_consume_pending(ctx, g));
      g->last_use = ctx->tick;
      *out = g;
      ctx->last_grid = g;
      return 0;
    }
  }
  // a grid of an earlier surface is of no use any more: recycle its buffers before allocating another set (a new
  // cloud per step must not reach cudaMalloc in the steady state); among the stale ones a grid whose buffers already
  // hold this many points comes first (small clouds described in between leave small grids behind: growing them one
  // per step would put a cudaFree - a device-wide synchronisation - into each of the next steps), then the oldest
  const size_t need_bytes = std::max<size_t>(ctx->n, 1) * sizeof(float4);
  for (Grid* g : ctx->grids) {
    if (g->surf_version == ctx->surf_version) continue;
    const bool fits = g->pts.cap >= need_bytes, vfits = victim && victim->pts.cap >= need_bytes;
    if (!victim || (fits && !vfits) || (fits == vfits && g->last_use < victim->last_use)) victim = g;
  }
  if (victim) {
  } else if (ctx->grids.size() < 4) {
    victim = new Grid();
    ctx->grids.push_back(victim);
  } else {
    for (Grid* g : ctx->grids)  // prefer stale grids, then least recently used
      if (!victim || (g->surf_version != ctx->surf_version && victim->surf_version == ctx->surf_version) ||
          ((g->surf_version != ctx->surf_version) == (victim->surf_version != ctx->surf_version) &&
           g->last_use < victim->last_use))
        victim = g;
  }
  if (ctx->normals_sorted_for == victim) ctx->normals_sorted_for = nullptr;
  if (ctx->knn_grid == victim) ctx->knn_grid = nullptr;
  PFX_TRY(grid_consume_pending(ctx, victim));  // a build still in flight on the other stream owns these buffers
  victim->last_use = ctx->tick;
  int rc = grid_build(ctx, victim, radius, knn_k);
  if (rc != 0) {
    victim->surf_version = 0;
    return rc;
  }
  *out = victim;
  ctx->last_grid = victim;
  return 0;
}

// ------------------------------------------------------------------------------------- VoxelGrid
// pcl::VoxelGrid centroid filter (config C1 ingest): voxel id = (ix - min_bx) + (iy - min_by) * div_bx +
// (iz - min_bz) * div_bx * div_by with ix = floor(x * inv_leaf) in float; output = one centroid per occupied
// voxel in ascending id order.  Sort-and-scan like the voxel hash; the centroid of a voxel is summed by ONE
// thread in ascending point-index order, which is the order PCL's sorted index vector yields, so the float
// sums are bit-identical to the CPU.
__global__ void vg_keys_kernel(const float4* __restrict__ pts, int n, float inv, float minbx, float minby, float minbz,
                               int divx, int divxy, uint32_t* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  uint32_t key = KEY_INVALID;
  if (finite3(p.x, p.y, p.z)) {
    int a = (int)(floorf(__fmul_rn(p.x, inv)) - minbx), b = (int)(floorf(__fmul_rn(p.y, inv)) - minby),
        c = (int)(floorf(__fmul_rn(p.z, inv)) - minbz);
    key = (uint32_t)(a + b * divx + c * divxy);
  }
  keys[i] = key;
  vals[i] = i;
}

__global__ void vg_starts_kernel(const uint32_t* __restrict__ keys, const int* __restrict__ heads,
                                 const int* __restrict__ excl, int n, int* __restrict__ starts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (heads[i]) starts[excl[i]] = i;
}

__global__ void vg_centroid_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                   const int* __restrict__ vals, const int* __restrict__ starts,
                                   const int* __restrict__ nvox_dev, int n, int cap, float* __restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int nvox = *nvox_dev;
  if (v >= nvox || v >= cap) return;
  const int s = starts[v];
  const uint32_t key = keys[s];
  float cx = 0.f, cy = 0.f, cz = 0.f;
  int e = s;
  for (; e < n && keys[e] == key; ++e) {
    float4 p = pts[vals[e]];
    cx = __fadd_rn(cx, p.x);
    cy = __fadd_rn(cy, p.y);
    cz = __fadd_rn(cz, p.z);
  }
  const float cnt = (float)(e - s);
  out[3 * (size_t)v] = __fdiv_rn(cx, cnt);
  out[3 * (size_t)v + 1] = __fdiv_rn(cy, cnt);
  out[3 * (size_t)v + 2] = __fdiv_rn(cz, cnt);
}

// out_dev: cap x 3 floats (device).  *n_out = number of occupied voxels (may exceed cap: nothing beyond cap is
// written and the caller reports PFX_E_CAPACITY).
int voxel_grid_run(Ctx* ctx, float leaf, float* out_dev, size_t cap, size_t* n_out) {
  const int n = (int)ctx->n;
  *n_out = 0;
  if (n == 0) return 0;
  Grid* g = &ctx->vg_scratch;  // sort buffers only
  PFX_CUDA(g->misc.ensure(sizeof(BuildAcc) + 64));
  PFX_CUDA(g->keys.ensure((size_t)n * sizeof(uint32_t)));
  PFX_CUDA(g->keys2.ensure((size_t)n * sizeof(uint32_t)));
  PFX_CUDA(g->vals.ensure((size_t)n * sizeof(int)));
  PFX_CUDA(g->vals2.ensure((size_t)n * sizeof(int)));
  PFX_CUDA(g->pt_cell.ensure((size_t)n * sizeof(int)));
  PFX_CUDA(g->cell_start.ensure(((size_t)n + 1) * sizeof(int)));
  BuildAcc* acc = g->misc.as<BuildAcc>();
  const float4* src = ctx->surf.as<float4>();
  const int T = 256;
  PFX_CUDA(cudaMemsetAsync(acc, 0, sizeof(BuildAcc) + 64, ctx->stream));
  PFX_LAUNCH(ctx, bbox_kernel, std::min(ctx->sm_count * 4, div_up(n, T)), T, 0, src, n, acc);
  BuildAcc h;
  PFX_CUDA(cudaMemcpyAsync(&h, acc, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h.n_valid == 0) return 0;
  auto ord2f_host = [](uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &b, 4);
    return f;
  };
  const float inv = 1.0f / leaf;
  long long minb[3], divb[3];
  for (int a = 0; a < 3; ++a) {
    float mn = ord2f_host(~h.mn_inv[a]), mx = ord2f_host(h.mx[a]);
    minb[a] = (long long)floorf(mn * inv);
    divb[a] = (long long)floorf(mx * inv) - minb[a] + 1;
  }
  // PCL: "Leaf size is too small for the input dataset. Integer indices would overflow."
  if (divb[0] * divb[1] * divb[2] > 0x7fffffffll) return ctx->fail(PFX_E_PRECOND, "pfx_voxel_grid: leaf size too small for the dataset (voxel index overflows int32)");
  PFX_LAUNCH(ctx, vg_keys_kernel, div_up(n, T), T, 0, src, n, inv, (float)minb[0], (float)minb[1], (float)minb[2],
             (int)divb[0], (int)(divb[0] * divb[1]), g->keys.as<uint32_t>(), g->vals.as<int>());
  PFX_TRY(radix_sort_pairs(ctx, g, n));
  int* heads = g->vals2.as<int>();
  int* total = reinterpret_cast<int*>(g->misc.as<char>() + sizeof(BuildAcc));
  PFX_LAUNCH(ctx, heads_kernel, div_up(n, T), T, 0, g->keys.as<uint32_t>(), n, heads);
  PFX_TRY(scan_exclusive_i32(ctx, heads, g->pt_cell.as<int>(), n, total, g->bsum));
  PFX_LAUNCH(ctx, vg_starts_kernel, div_up(n, T), T, 0, g->keys.as<uint32_t>(), heads, g->pt_cell.as<int>(), n,
             g->cell_start.as<int>());
  int nvox = 0;
  PFX_CUDA(cudaMemcpyAsync(&nvox, total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_out = (size_t)nvox;
  if (nvox > 0 && cap > 0)
    PFX_LAUNCH(ctx, vg_centroid_kernel, div_up(std::min<long long>(nvox, (long long)cap), T), T, 0, src, g->keys.as<uint32_t>(),
               g->vals.as<int>(), g->cell_start.as<int>(), total, n, (int)std::min<size_t>(cap, 0x7fffffff), out_dev);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

void grid_free_all(Ctx* ctx) {
  ctx->vg_scratch.release();
  for (Grid* g : ctx->grids) {
    if (g->ready) cudaEventDestroy(g->ready);
    g->release();
    delete g;
  }
  ctx->grids.clear();
}

}  // namespace pfx

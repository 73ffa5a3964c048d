// match.cu — exact descriptor 1-NN (replaces Features<T>::getCorrespondences, reference
// features.h:253-273: KdTreeFLANN<FeatureT>::nearestKSearch(k = 1) for every source descriptor; in
// 33/36/352 dimensions that tree degenerates to a linear scan).
//
// match_exact_kernel: register-tiled all-pairs scan.  Every (a_i, b_j) distance is the sequential
// float sum over dimensions 0..D-1 of (a-b)^2 with separate multiply and add (FLANN L2_Simple, no
// FMA), so distances are bit-identical to the CPU path; the argmin is merged across tiles with a
// 64-bit atomicMin on (d2 bits << 32 | j), which is exactly "smallest distance, then lowest index".
// Rows containing a non-finite value are excluded (targets) / never matched (queries).
#include "internal.h"

namespace pfx {

constexpr int MT = 64;   // rows of A and of B per tile
constexpr int MDK = 32;  // dimensions staged per step
constexpr unsigned long long PACK_NONE = 0xffffffffffffffffull;

__global__ void row_finite_kernel(const float* __restrict__ m, int rows, int ld, int dim, unsigned char* __restrict__ ok) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  bool good = true;
  for (int d = lane; d < dim; d += 32) good = good && isfinite(m[(size_t)r * ld + d]);
  good = __all_sync(FULL, good);
  if (lane == 0) ok[r] = good ? 1 : 0;
}

__global__ void pack_init_kernel(unsigned long long* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = PACK_NONE;
}

// grid: (ceil(na / MT), splits); block 256 = 16 x 16; thread (ty, tx) owns a rows ty*4..ty*4+3 and
// b rows tx + 16*j (j = 0..3) of the tile (bank-conflict-free shared reads).
__global__ void __launch_bounds__(256)
match_exact_kernel(const float* __restrict__ A, int na, int lda, const unsigned char* __restrict__ aok,
                   const float* __restrict__ B, int nb, int ldb, const unsigned char* __restrict__ bok, int dim,
                   int tiles_per_split, unsigned long long* __restrict__ best) {
  __shared__ float As[MT][MDK + 1];
  __shared__ float Bs[MT][MDK + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int a0 = blockIdx.x * MT;
  const int nbt = (nb + MT - 1) / MT;
  const int t_begin = blockIdx.y * tiles_per_split, t_end = min(nbt, t_begin + tiles_per_split);
  unsigned long long mybest[4] = {PACK_NONE, PACK_NONE, PACK_NONE, PACK_NONE};
  for (int bt = t_begin; bt < t_end; ++bt) {
    const int b0 = bt * MT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d0 = 0; d0 < dim; d0 += MDK) {
      __syncthreads();
      for (int e = threadIdx.x; e < MT * MDK; e += 256) {
        int r = e / MDK, c = e % MDK;
        int d = d0 + c;
        As[r][c] = (a0 + r < na && d < dim) ? A[(size_t)(a0 + r) * lda + d] : 0.f;
        Bs[r][c] = (b0 + r < nb && d < dim) ? B[(size_t)(b0 + r) * ldb + d] : 0.f;
      }
      __syncthreads();
      const int dk = min(MDK, dim - d0);
      for (int c = 0; c < dk; ++c) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[ty * 4 + i][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[tx + 16 * j][c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float df = __fsub_rn(av[i], bv[j]);
            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(df, df));
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int bj = b0 + tx + 16 * j;
      if (bj < nb && bok[bj]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          unsigned long long key = ((unsigned long long)__float_as_uint(acc[i][j]) << 32) | (unsigned)bj;
          if (key < mybest[i]) mybest[i] = key;
        }
      }
    }
  }
  // reduce over the 16 tx lanes that share the same a rows (a half-warp), then one atomic per row
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unsigned long long k = mybest[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(FULL, k, o);
      k = other < k ? other : k;
    }
    int ai = a0 + ty * 4 + i;
    if (tx == 0 && ai < na && aok[ai] && k != PACK_NONE) atomicMin(&best[ai], k);
  }
}

// A FEW query rows against all targets (the rows the tensor-core matcher could not certify: a handful per call).  The
// tiled kernel above would spend a 64-row tile on them; here one THREAD owns one target row and keeps the distances
// to up to FEW_R query rows (staged in shared memory) in registers, walking its row once - the target matrix is read
// once per group of FEW_R query rows.  Same sequential float sum, same (d2, index) merge; a target row holding a
// non-finite value is skipped.
constexpr int FEW_R = 8, FEW_THREADS = 128, FEW_MAX_ROWS = 64, FEW_MAX_DIM = 1024;

__global__ void __launch_bounds__(FEW_THREADS)
match_few_kernel(const float* __restrict__ A, int na, int lda, const unsigned char* __restrict__ aok,
                 const float* __restrict__ B, int nb, int ldb, int dim, unsigned long long* __restrict__ best) {
  extern __shared__ float few_as[];  // [FEW_R][dim]
  __shared__ unsigned long long wbest[FEW_THREADS / 32][FEW_R];
  const int j = blockIdx.x * FEW_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int r0 = 0; r0 < na; r0 += FEW_R) {
    __syncthreads();
    for (int e = threadIdx.x; e < FEW_R * dim; e += FEW_THREADS) {
      const int i = e / dim, d = e - i * dim;
      few_as[e] = (r0 + i < na) ? A[(size_t)(r0 + i) * lda + d] : 0.f;
    }
    __syncthreads();
    float acc[FEW_R];
#pragma unroll
    for (int i = 0; i < FEW_R; ++i) acc[i] = 0.f;
    bool good = j < nb;
    if (good) {
      const float* brow = B + (size_t)j * ldb;
      int d = 0;
      if ((((size_t)brow) & 15) == 0) {  // 16-byte loads of the target row when it allows them
        const float4* b4 = reinterpret_cast<const float4*>(brow);
        for (; d + 4 <= dim; d += 4) {
          const float4 bq = b4[d >> 2];
          const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            good = good && isfinite(bv[e]);
#pragma unroll
            for (int i = 0; i < FEW_R; ++i) {
              const float df = __fsub_rn(few_as[i * dim + d + e], bv[e]);
              acc[i] = __fadd_rn(acc[i], __fmul_rn(df, df));
            }
          }
        }
      }
      for (; d < dim; ++d) {
        const float bv = brow[d];
        good = good && isfinite(bv);
#pragma unroll
        for (int i = 0; i < FEW_R; ++i) {
          const float df = __fsub_rn(few_as[i * dim + d], bv);
          acc[i] = __fadd_rn(acc[i], __fmul_rn(df, df));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < FEW_R; ++i) {
      unsigned long long key = good ? (((unsigned long long)__float_as_uint(acc[i]) << 32) | (unsigned)j) : PACK_NONE;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, key, o);
        key = other < key ? other : key;
      }
      if (lane == 0) wbest[wid][i] = key;
    }
    __syncthreads();
    if (threadIdx.x < FEW_R) {
      const int i = threadIdx.x;
      unsigned long long key = wbest[0][i];
#pragma unroll
      for (int w = 1; w < FEW_THREADS / 32; ++w) key = wbest[w][i] < key ? wbest[w][i] : key;
      const int ai = r0 + i;
      if (ai < na && aok[ai] && key != PACK_NONE) atomicMin(&best[ai], key);
    }
  }
}

__global__ void unpack_kernel(const unsigned long long* __restrict__ best, int n, int* __restrict__ idx,
                              float* __restrict__ d2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = best[i];
  idx[i] = (k == PACK_NONE) ? -1 : (int)(unsigned)(k & 0xffffffffull);
  if (d2) d2[i] = (k == PACK_NONE) ? CUDART_INF_F : __uint_as_float((unsigned)(k >> 32));
}

// a, b: device pointers; lda/ldb in floats; nn_idx/nn_d2: device, sized na
int match_nn_exact(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* nn_idx,
                   float* nn_d2) {
  if (na == 0) return 0;
  DevBuf& flags = ctx->match_flags;
  DevBuf& best = ctx->match_best;
  PFX_CUDA(flags.ensure((size_t)na + (size_t)std::max(nb, 1) + 16));
  PFX_CUDA(best.ensure((size_t)na * sizeof(unsigned long long)));
  unsigned char* aok = flags.as<unsigned char>();
  unsigned char* bok = aok + na;
  PFX_LAUNCH(ctx, row_finite_kernel, div_up(na, 8), 256, 0, a, na, lda, dim, aok);
  PFX_LAUNCH(ctx, pack_init_kernel, div_up(na, 256), 256, 0, best.as<unsigned long long>(), na);
  if (nb > 0 && na <= FEW_MAX_ROWS && dim <= FEW_MAX_DIM && nb >= 4096) {
    // a handful of query rows: one thread per target row (finite check of the targets inline)
    PFX_LAUNCH(ctx, match_few_kernel, div_up(nb, FEW_THREADS), FEW_THREADS, (size_t)FEW_R * dim * sizeof(float), a, na, lda,
               aok, b, nb, ldb, dim, best.as<unsigned long long>());
  } else if (nb > 0) {
    PFX_LAUNCH(ctx, row_finite_kernel, div_up(nb, 8), 256, 0, b, nb, ldb, dim, bok);
    const int nat = div_up(na, MT), nbt = div_up(nb, MT);
    // enough blocks for ~4 waves of the machine, never more splits than B tiles
    int splits = std::max(1, std::min(nbt, (ctx->sm_count * 4 + nat - 1) / nat));
    int tps = div_up(nbt, splits);
    splits = div_up(nbt, tps);
    dim3 grid(nat, splits);
    PFX_LAUNCH(ctx, match_exact_kernel, grid, 256, 0, a, na, lda, aok, b, nb, ldb, bok, dim, tps,
               best.as<unsigned long long>());
  }
  PFX_LAUNCH(ctx, unpack_kernel, div_up(na, 256), 256, 0, best.as<unsigned long long>(), na, nn_idx, nn_d2);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx

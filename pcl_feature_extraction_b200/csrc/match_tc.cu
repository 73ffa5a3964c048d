// match_tc.cu — descriptor 1-NN through the 5th-generation tensor cores (replaces the same reference
// code as match.cu: Features<T>::getCorrespondences, features.h:253-273, KdTreeFLANN<FeatureT> 1-NN in
// 33 / 36 / 352 dimensions, i.e. a linear scan).  The one dense contraction of the path.
//
//   d~2(i, j) = |a~_i|^2 + |b~_j|^2 - 2 a~_i . b~_j        (a~, b~ = operands rounded to bf16)
//
// Pipeline (all on the context's stream):
//  1. tc_prep_kernel      fp32 rows -> bf16 in the UMMA "core matrix" order (8 rows x 16 bytes, K-major,
//                         no swizzle), 128-row tiles, so that a tile (or a K-slab of it) is ONE contiguous
//                         block of global memory; per-row |x~|^2 and the rounding error |x - x~|.
//  2. tc_candidates_kernel warp-specialised: warp 0 streams B K-slabs with cp.async.bulk (TMA bulk copies,
//                         mbarrier complete_tx) through an N-stage ring, warp 1 issues tcgen05.mma
//                         (cta_group::1, kind::f16, M = 128, N = 128 per half tile, fp32 accumulators in
//                         TMEM, two 256-column accumulator buffers), warps 4-7 read the accumulators back
//                         with tcgen05.ld (one thread = one A row) and keep a running top-K of d~2 in
//                         registers while the next tile's MMAs run.
//  3. tc_rescore_kernel   exact fp32 distance (FLANN L2_Simple: sequential sum of (a-b)^2, no FMA) of the
//                         candidates, argmin with the (d2, index) rule, and a CERTIFICATE that no
//                         non-candidate can beat it: sqrt(d2*) < sqrt(kth d~2 - delta) - (|a-a~| + max|b-b~|).
//  4. rows that fail the certificate are redone by the exact all-pairs kernel (match.cu).
// Result: indices and distances are bit-identical to the exact path.
#include <cuda_bf16.h>

#include <cmath>

#include "internal.h"

namespace pfx {

#ifndef PFX_TC_UNSORTED_LIST
#define PFX_TC_UNSORTED_LIST 1
#endif
constexpr int TC_TILE = 128;              // rows per operand tile (= UMMA M, and N of one half tile)
constexpr int TC_K = 8;                   // candidates kept per (row, column split)
constexpr int TC_SLAB = TC_TILE * 16;     // bytes of one K-slab (8 bf16 of K for 128 rows)
constexpr int TC_STAGE = 2 * 4 * TC_SLAB; // a stage = 2 half tiles x 4 slabs (K = 32)
#ifndef PFX_TC_EPI_WARPS
#define PFX_TC_EPI_WARPS 8
#endif
constexpr int TC_EPI_WARPS = PFX_TC_EPI_WARPS;  // epilogue warps: (TMEM lane quarter) x (half or quarter of the 256 columns)
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = 128 + TC_EPI_THREADS;
constexpr int TC_LISTS = TC_EPI_WARPS / 4;      // top-K lists per (row, split): one per column group of a thread
constexpr int TC_COLS = 256 / TC_LISTS;         // columns of a pair per epilogue thread
constexpr int TC_KS = TC_K + 1;           // list stride in the candidate arrays: K entries + the bound slot
constexpr float TC_BIG = 1e30f;           // |x~|^2 of a row that must never match (non-finite / padding)

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one thread of the LEADER CTA issues an MMA of M = 256 (128 rows from each
// CTA's A tile) x N = 256 (128 columns from each CTA's B stage); both CTAs' TMEM receive their own 128 rows.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the address of one of my shared-memory objects in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // (default semantics, as CUTLASS' ClusterBarrier::arrive(cta_id): what the arrival publishes are TMEM reads fenced
  // with tcgen05.fence and shared memory written by the async proxy, not generic-proxy stores; release.cluster costs a
  // MEMBAR + ERRBAR per arrival - a quarter of the kernel's stall samples when the follower's epilogue warps and relay
  // lanes used it)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from the other CTA of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor: core matrices of 8 rows x 16 bytes stored as 128
// contiguous bytes; LBO = byte step between the two 16-byte K chunks of one MMA (one slab), SBO = byte
// step between 8-row groups; version 1 (Blackwell) in bits 46-47.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((TC_SLAB >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_TILE >> 3) << 17) |
                              ((uint32_t)(TC_TILE >> 4) << 24);
// kind::tf32: A = B = tf32 (format 2), K = 8 per instruction; a 16-byte K chunk then holds 4 elements, so the
// slab / stage / descriptor geometry in BYTES is the same as for bf16
constexpr uint32_t TC_IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_TILE >> 3) << 17) |
                                   ((uint32_t)(TC_TILE >> 4) << 24);

// CTA pair: M = 256, N = 256
constexpr uint32_t TC_IDESC2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t TC_IDESC2_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// ------------------------------------------------------------------------------------------ prep
// fp32 rows -> operand tiles + |x~|^2 + |x - x~| (+ the maxima of the latter two over the rows).  A block of 8 warps
// converts EIGHT rows (one row octet of a 128-row tile): lane = (row of the octet, quarter of a 16-byte K chunk), warp
// w takes the chunks w, w + 8, ...  The eight rows' 16-byte pieces of one chunk are 128 contiguous bytes of the tile
// (core-matrix order), so every warp store is one full line; loads are 32-byte row segments.  A row holding a
// non-finite value becomes zeros with an infinite norm (never a candidate).  Sums are combined in a fixed order.
// TF32 = false: bf16 elements (8 per 16-byte chunk); true: tf32 = fp32 with a 10-bit mantissa (4 per chunk).
template <bool TF32>
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ X, int n, int ld, int dim, int dpad, int npad, unsigned char* __restrict__ Xt,
               float* __restrict__ norm, float* __restrict__ err,
               unsigned* __restrict__ maxima /* [0] max err bits, [1] max finite norm bits */) {
  constexpr int EPC = TF32 ? 4 : 8;  // elements per 16-byte K chunk
  constexpr int EPT = EPC / 4;       // elements per thread and chunk
  __shared__ int s_bad[8];
  __shared__ float s_s2[8][8], s_e2[8][8];  // [warp][row]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ro = lane >> 2, part = lane & 3;
  const int r = blockIdx.x * 8 + ro;  // npad is a multiple of 128: r < npad
  const int nchunk = dpad / EPC;
  if (threadIdx.x < 8) s_bad[threadIdx.x] = 0;
  __syncthreads();
  const float* row = X + (size_t)r * ld;
  const bool in = r < n;
  const size_t tile_base = (size_t)(r >> 7) * (size_t)nchunk * TC_SLAB;  // bytes
  const int rr = r & 127;
  const size_t row_off = (size_t)(rr >> 3) * 128 + (size_t)(rr & 7) * 16;
  // ONE pass over the row: convert, store and look for non-finite values on the way; a row that holds one (rare) is
  // overwritten with zeros afterwards.  (A separate finiteness pass read every row twice.)
  bool bad = false;
  float s2 = 0.f, e2 = 0.f;
  for (int c = w; c < nchunk; c += 8) {
    float v[EPT], vr[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int k = c * EPC + part * EPT + e;
      v[e] = (in && k < dim) ? row[k] : 0.f;
      bad = bad || !isfinite(v[e]);
    }
    unsigned char* dst = Xt + tile_base + (size_t)c * TC_SLAB + row_off + (size_t)part * 4;
    if (TF32) {
      uint32_t bits;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bits) : "f"(v[0]));
      vr[0] = __uint_as_float(bits);
      *reinterpret_cast<float*>(dst) = vr[0];
    } else {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[0]), h1 = __float2bfloat16_rn(v[EPT - 1]);
      vr[0] = __bfloat162float(h0);
      vr[EPT - 1] = __bfloat162float(h1);
      __nv_bfloat162 pair;
      pair.x = h0;
      pair.y = h1;
      *reinterpret_cast<__nv_bfloat162*>(dst) = pair;
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      s2 = fmaf(vr[e], vr[e], s2);
      const float dv = v[e] - vr[e];
      e2 = fmaf(dv, dv, e2);
    }
  }
  if (bad) s_bad[ro] = 1;
  __syncthreads();
  if (s_bad[ro]) {  // the whole row becomes zeros with an infinite norm (never a candidate)
    for (int c = w; c < nchunk; c += 8)
      *reinterpret_cast<uint32_t*>(Xt + tile_base + (size_t)c * TC_SLAB + row_off + (size_t)part * 4) = 0u;
  }
  // the four lanes of a row, then the eight warps in a fixed order
  s2 += __shfl_xor_sync(FULL, s2, 1);
  s2 += __shfl_xor_sync(FULL, s2, 2);
  e2 += __shfl_xor_sync(FULL, e2, 1);
  e2 += __shfl_xor_sync(FULL, e2, 2);
  if (part == 0) {
    s_s2[w][ro] = s2;
    s_e2[w][ro] = e2;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int rw = blockIdx.x * 8 + threadIdx.x;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      a += s_s2[ww][threadIdx.x];
      b += s_e2[ww][threadIdx.x];
    }
    const bool f = rw < n && !s_bad[threadIdx.x];
    const float e = f ? sqrtf(b) * 1.001f + 1e-30f : 0.f;
    norm[rw] = f ? a : TC_BIG;
    err[rw] = e;
    if (f) {
      atomicMax(&maxima[0], __float_as_uint(e));
      atomicMax(&maxima[1], __float_as_uint(a));
    }
  }
}

// ------------------------------------------------------------------------------------------ GEMM + top-K
// (bits(v) & ~127) | col in one LOP3 (truth table 0xEA = (a & b) | c)
__device__ __forceinline__ float tc_embed(float v, int col) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(__float_as_uint(v)), "r"(0xFFFFFF80u), "r"((uint32_t)col));
  return __uint_as_float(r);
}

// Cold path of the epilogue: insert (d, j) into the calling thread's ascending top-K list, which lives in
// shared memory ([slot][thread], conflict-free) so that the hot loop carries only the K-th value in a
// register and stays small enough for the instruction cache.  Returns the new K-th value.
__device__ __noinline__ float tc_topk_insert(float* sd, int* sj, float d, int j) {
  int i = TC_K - 1;
  while (i > 0) {
    const float prev = sd[(i - 1) * TC_EPI_THREADS];
    if (!(d < prev)) break;
    sd[i * TC_EPI_THREADS] = prev;
    sj[i * TC_EPI_THREADS] = sj[(i - 1) * TC_EPI_THREADS];
    --i;
  }
  sd[i * TC_EPI_THREADS] = d;
  sj[i * TC_EPI_THREADS] = j;
  return sd[(TC_K - 1) * TC_EPI_THREADS];
}

struct TcArgs {
  const unsigned char* At;  // operand tiles (bf16 or tf32, see tc_prep_kernel)
  const unsigned char* Bt;
  const float* na;
  const float* nb;
  int n_btiles;         // 128-row tiles of B
  int nslab;            // 16-byte K chunks per row (dpad / 8 for bf16, dpad / 4 for tf32)
  int nsplit;           // column splits (units per A tile)
  int pairs_per_split;  // 256-column tile pairs per unit
  int nstage;
  int n_atiles;         // 128-row tiles of A (pair kernel: the second CTA of the last pair may have none)
  float* cand_d;        // [A rows padded][nsplit][TC_LISTS][TC_KS]; slot TC_K of a list = its bound
  int* cand_j;
};

// PAIR: two CTAs of a cluster (one SM each) work on two A tiles against the SAME column pairs.  Each loads ONE B tile of
// a pair (half the L2 -> shared-memory traffic per SM: with one CTA per unit every SM re-streams all of B and the
// kernel paces on that), the leader's MMA thread issues 256 x 256 x 16 instructions for both (cta_group::2), and each
// CTA's epilogue reads its own 128 rows x 256 columns from its own TMEM, exactly as in the single-CTA kernel.
// Cross-CTA signalling: the leader's commits arrive on the stage-free and accumulator-ready barriers of both CTAs
// (multicast); bulk copies can only signal a barrier of the CTA they land in, so the follower's idle MMA warp relays
// "my half of stage s has landed" to the leader's stage barrier, and the follower's epilogue warps arrive on the
// leader's accumulator-free barrier.
template <bool TF32, bool PAIR>
__device__ __forceinline__ void tc_candidates_body(const TcArgs& P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0u;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int split = unit % P.nsplit;
  const int a_tile_raw = PAIR ? 2 * (unit / P.nsplit) + (int)rank : unit / P.nsplit;
  const bool a_valid = a_tile_raw < P.n_atiles;
  const int a_tile = a_valid ? a_tile_raw : P.n_atiles - 1;  // (an A tile to keep the pair's MMAs fed; nothing is written)
  constexpr uint32_t STAGE_BYTES = PAIR ? TC_STAGE / 2 : TC_STAGE;
  const int nslab = P.nslab;
  const int nchunk = (nslab + 3) >> 2;
  const int npairs_all = (P.n_btiles + 1) >> 1;
  const int p0 = split * P.pairs_per_split, p1 = min(npairs_all, p0 + P.pairs_per_split);

  unsigned char* sA = smem;
  unsigned char* sB = sA + (size_t)nslab * TC_SLAB;
  float* nbs = reinterpret_cast<float*>(sB + (size_t)P.nstage * STAGE_BYTES);  // [2][256]
  float* list_d = nbs + 512;                                               // [TC_K][epilogue threads]
  int* list_j = reinterpret_cast<int*>(list_d + TC_K * TC_EPI_THREADS);    // [TC_K][epilogue threads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(list_j + TC_K * TC_EPI_THREADS);
  // bars: [0, nstage) full, [nstage, 2 nstage) empty, then afull, tfull[2], tempty[2], apeer (the follower's A tile)
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + P.nstage),
                 bar_afull = smem_u32(bars + 2 * P.nstage), bar_tfull = smem_u32(bars + 2 * P.nstage + 1),
                 bar_tempty = smem_u32(bars + 2 * P.nstage + 3), bar_apeer = smem_u32(bars + 2 * P.nstage + 5);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P.nstage + 6);

  if (warp == 1 && lane == 0) {
    // (pair, leader) a stage is full when my own bytes have landed AND the follower has reported its half; an
    // accumulator buffer is free when the epilogue warps of both CTAs are through with it
    const uint32_t both = (PAIR && leader) ? 2u : 1u;
    for (int s = 0; s < P.nstage; ++s) {
      mbar_init(bar_full + 8 * s, both);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_afull, 1);
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_tempty, both * TC_EPI_WARPS);
    mbar_init(bar_tempty + 8, both * TC_EPI_WARPS);
    mbar_init(bar_apeer, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PAIR) cluster_sync_all();  // the other CTA's barriers exist before anything arrives on them
  if (warp == 2) {
    if (PAIR) {  // (the same warp of BOTH CTAs issues the pair allocation)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: A once, then the B slabs of every tile pair through the ring
    if (lane == 0) {
      const unsigned char* srcA = P.At + (size_t)a_tile * nslab * TC_SLAB;
      const uint32_t a_bytes = (uint32_t)nslab * TC_SLAB;
      mbar_expect_tx(bar_afull, a_bytes);
      for (uint32_t off = 0; off < a_bytes; off += 16384u)
        bulk_g2s(smem_u32(sA + off), srcA + off, min(16384u, a_bytes - off), bar_afull);
      int s = 0;
      uint32_t ph = 0;
      const unsigned char* srcB = P.Bt;
      for (int p = p0; p < p1; ++p) {
        const int nh = min(2, P.n_btiles - 2 * p);
        for (int c = 0; c < nchunk; ++c) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const int ns = min(4, nslab - 4 * c);
          const uint32_t bytes = (uint32_t)ns * TC_SLAB;
          const uint32_t dst = smem_u32(sB) + (uint32_t)s * STAGE_BYTES;
          if (PAIR) {
            // my B tile of the pair: 2p + rank (the last pair of an odd tile count: the follower repeats tile 2p; its
            // 128 columns are computed and ignored)
            const int bt = min(2 * p + (int)rank, P.n_btiles - 1);
            mbar_expect_tx(bar_full + 8 * s, bytes);
            bulk_g2s(dst, srcB + ((size_t)bt * nslab + (size_t)4 * c) * TC_SLAB, bytes, bar_full + 8 * s);
          } else {
            mbar_expect_tx(bar_full + 8 * s, bytes * nh);
            const unsigned char* src = srcB + ((size_t)(2 * p) * nslab + (size_t)4 * c) * TC_SLAB;
            bulk_g2s(dst, src, bytes, bar_full + 8 * s);
            if (nh == 2) bulk_g2s(dst + 4 * TC_SLAB, src + (size_t)nslab * TC_SLAB, bytes, bar_full + 8 * s);
          }
          if (++s == P.nstage) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread).  The loop must stay far below the tensor pipe's 64 cycles per
    // MMA, so descriptors are one 32-bit add away from a precomputed base (only the 14-bit start-address field
    // changes) and the ring position is a counter, not a division.
    if (PAIR && !leader) {
      // follower: no MMAs to issue - relay "my A tile / my half of stage s has landed" to the leader's barriers
      // (one lane per ring stage: a remote arrive is a round trip through the cluster, and one thread relaying every
      // stage in turn could not keep up with the MMAs)
      if (lane == 0) {
        mbar_wait(bar_afull, 0);
        mbar_arrive_cluster(mapa_u32(bar_apeer, 0));
      }
      if (lane < P.nstage) {
        const uint32_t mine = bar_full + 8 * lane, lead = mapa_u32(mine, 0);
        const int uses = (p1 - p0) * nchunk;
        uint32_t ph = 0;
        for (int u = lane; u < uses; u += P.nstage) {
          mbar_wait(mine, ph);
          mbar_arrive_cluster(lead);
          ph ^= 1u;
        }
      }
    } else if (lane == 0) {
      mbar_wait(bar_afull, 0);
      if (PAIR) mbar_wait_cluster(bar_apeer, 0);
      const uint64_t desc_hi = (uint64_t)(((128u >> 4) & 0x3FFFu) | (1u << 14)) << 32;  // SBO = 128 B, version 1
      const uint32_t lbo = (uint32_t)((TC_SLAB >> 4) & 0x3FFF) << 16;
      const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3FFFu) | lbo;
      const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFFu) | lbo;
      constexpr uint32_t SLAB16 = TC_SLAB >> 4, STAGE16 = STAGE_BYTES >> 4;
      const int last_nm = (nslab - 4 * (nchunk - 1)) >> 1;  // MMA steps (two K chunks each) of the last chunk: 1 or 2
      auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
        if (PAIR) {
          if (TF32) umma2_tf32(d, ad, bd, TC_IDESC2_TF32, acc);
          else umma2_bf16(d, ad, bd, TC_IDESC2, acc);
        } else {
          if (TF32) umma_tf32(d, ad, bd, TC_IDESC_TF32, acc);
          else umma_bf16(d, ad, bd, TC_IDESC, acc);
        }
      };
      auto commit = [](uint32_t bar) {
        if (PAIR) umma2_commit(bar);
        else umma_commit(bar);
      };
      int s = 0;
      uint32_t ph = 0;
      for (int p = p0; p < p1; ++p) {
        const int lp = p - p0, buf = lp & 1;
        const bool two = (P.n_btiles - 2 * p) >= 2;
        if (PAIR) mbar_wait_cluster(bar_tempty + 8 * buf, (uint32_t)(((lp >> 1) & 1) ^ 1));
        else mbar_wait(bar_tempty + 8 * buf, (uint32_t)(((lp >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * 256), d1 = d0 + 128;
        for (int c = 0; c < nchunk; ++c) {
          if (PAIR) mbar_wait_cluster(bar_full + 8 * s, ph);
          else mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)c * (4 * SLAB16);
          const uint32_t b_lo = b_lo0 + (uint32_t)s * STAGE16;
          const uint32_t acc0 = c ? 1u : 0u;
          // (pair: one instruction covers the 256 columns - 128 from each CTA's stage)
          mma(d0, desc_hi | a_lo, desc_hi | b_lo, acc0);
          if (!PAIR && two) mma(d1, desc_hi | a_lo, desc_hi | (b_lo + 4 * SLAB16), acc0);
          if (c + 1 < nchunk || last_nm == 2) {
            mma(d0, desc_hi | (a_lo + 2 * SLAB16), desc_hi | (b_lo + 2 * SLAB16), 1u);
            if (!PAIR && two) mma(d1, desc_hi | (a_lo + 2 * SLAB16), desc_hi | (b_lo + 6 * SLAB16), 1u);
          }
          commit(bar_empty + 8 * s);  // frees the stage (in both CTAs) when these MMAs have read it
          if (++s == P.nstage) {
            s = 0;
            ph ^= 1u;
          }
        }
        commit(bar_tfull + 8 * buf);  // accumulators of this pair are complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = (A row, column half).  Per 256-column pair the three smallest
    // keys of the thread's 128 columns come out of a branch-free min/max network (the column number rides
    // in the 7 low mantissa bits of d~2, a 2^-16 relative perturbation that the certificate's slack covers);
    // the two smallest are offered to the thread's running top-K list, the third lower-bounds everything
    // that was never offered.
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    const int row = a_tile * TC_TILE + q * 32 + lane;
    const float na = P.na[row];
    const int et = threadIdx.x - 128;
    float* sd = list_d + et;
    int* sj = list_j + et;
#pragma unroll
    for (int i = 0; i < TC_K; ++i) {
      sd[i * TC_EPI_THREADS] = CUDART_INF_F;
      sj[i * TC_EPI_THREADS] = -1;
    }
    float worst = CUDART_INF_F;  // K-th smallest key offered so far
#if PFX_TC_UNSORTED_LIST
    // The list is kept UNSORTED with its largest entry tracked (value `worst`, slot `wpos`): an offer that beats the
    // largest entry overwrites it and the largest of the eight is found again - a fixed, branch-free sequence.  With
    // 32 rows per warp and two offers per column pair some lane inserts in every pair, so the sorted insertion (a
    // data-dependent shift loop of dependent shared-memory loads and stores, behind a call) ran twice per pair for
    // the whole warp and took a fifth of the epilogue's time.
    int wpos = 0;
    auto offer = [&](float d, int j) {
      if (d < worst) {
        sd[wpos * TC_EPI_THREADS] = d;
        sj[wpos * TC_EPI_THREADS] = j;
        float mx = sd[0];
        int mp = 0;
#pragma unroll
        for (int i = 1; i < TC_K; ++i) {
          const float v = sd[i * TC_EPI_THREADS];
          if (v > mx) {
            mx = v;
            mp = i;
          }
        }
        worst = mx;
        wpos = mp;
      }
    };
#else
    auto offer = [&](float d, int j) {
      if (d < worst) worst = tc_topk_insert(sd, sj, d, j);
    };
#endif
    float bound = CUDART_INF_F;  // lower bound of every key never offered
    // |b~|^2 of the pair's 256 columns: one value per epilogue thread, fetched one pair ahead so that the global
    // load's latency is hidden behind the previous pair's min/max network
    auto load_nb = [&](int p) -> float {
      const int nh = min(2, P.n_btiles - 2 * p);
      return (p < p1 && et < nh * 128) ? P.nb[(size_t)p * 256 + et] : TC_BIG;
    };
    float nb_next = load_nb(p0);
    for (int p = p0; p < p1; ++p) {
      const int lp = p - p0, buf = lp & 1;
      const int nh = min(2, P.n_btiles - 2 * p);
      float* nbuf = nbs + buf * 256;
      if (et < 256) nbuf[et] = nb_next;
      nb_next = load_nb(p + 1);
      asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
      mbar_wait(bar_tfull + 8 * buf, (uint32_t)((lp >> 1) & 1));
      tc_fence_after();
      if (((hsel * TC_COLS) >> 7) < nh) {  // my columns belong to a B tile that exists
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + hsel * TC_COLS);
        const float* nh_buf = nbuf + hsel * TC_COLS;
        float m1 = 3e38f, m2 = 3e38f, m3 = 3e38f;
        // sorted insertion of the 32 keys of one accumulator slab into (m1 <= m2 <= m3)
        auto consume = [&](const uint32_t (&v)[32], int c0) {
#ifdef TC_SKIP_NETWORK  // experiment: pace of the MMA / TMA side alone
          uint32_t acc = 0;
#pragma unroll
          for (int c = 0; c < 32; ++c) acc ^= v[c];
          if (acc == 0x12345678u) m1 = 0.f;
          return;
#endif
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 n4 = *reinterpret_cast<const float4*>(nh_buf + c0 + c);
            const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              // two columns per step: with their min / max in hand the sorted insertion of both into
              // (m1 <= m2 <= m3) takes 8 min/max instructions (two of them 3-input) instead of 10
              // keys are |b~|^2 - 2 a~.b~: the row constant |a~|^2 does not change their order and is added to the
              // three survivors only.  (bits & ~127) | column as ONE three-input logic instruction.
              const float dx = fmaf(-2.f, __uint_as_float(v[c + u]), nn[u]);
              const float dy = fmaf(-2.f, __uint_as_float(v[c + u + 1]), nn[u + 1]);
              const float x = tc_embed(dx, c0 + c + u);
              const float y = tc_embed(dy, c0 + c + u + 1);
              const float lo = fminf(x, y), hi = fmaxf(x, y);
              const float n3 = fminf(m3, fminf(fmaxf(m2, lo), fmaxf(m1, hi)));
              const float n2 = fminf(fmaxf(m1, lo), fminf(m2, hi));
              m1 = fminf(m1, lo);
              m2 = n2;
              m3 = n3;
            }
          }
        };
        // two register buffers: the tcgen05.ld of slab i + 1 is in flight while slab i goes through the network
        uint32_t va[32], vb[32];
        tmem_ld32(tbase, va);
        tmem_ld_wait();
#pragma unroll
        for (int sl = 0; sl < TC_COLS / 32; sl += 2) {
          tmem_ld32(tbase + 32u * (sl + 1), vb);
          consume(va, 32 * sl);
          tmem_ld_wait();
          if (sl + 2 < TC_COLS / 32) tmem_ld32(tbase + 32u * (sl + 2), va);
          consume(vb, 32 * (sl + 1));
          if (sl + 2 < TC_COLS / 32) tmem_ld_wait();
        }
        const int jbase = p * 256 + hsel * TC_COLS;
        const int j1 = (int)(__float_as_uint(m1) & 0x7Fu), j2 = (int)(__float_as_uint(m2) & 0x7Fu);
        m1 += na;
        m2 += na;
        m3 += na;
        bound = fminf(bound, m3);
        offer(m1, jbase + j1);
        offer(m2, jbase + j2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_cluster(mapa_u32(bar_tempty + 8 * buf, 0));
        else mbar_arrive(bar_tempty + 8 * buf);
      }
    }
    if (a_valid) {
      float* od = P.cand_d + (((size_t)row * P.nsplit + split) * TC_LISTS + hsel) * TC_KS;
      int* oj = P.cand_j + (((size_t)row * P.nsplit + split) * TC_LISTS + hsel) * TC_KS;
#if PFX_TC_UNSORTED_LIST
      // the rescore reads slot K-1 as "the K-th entry of a full list": the largest entry goes there (an unfilled list
      // keeps an infinite entry with index -1 in that slot, as the sorted list did)
      {
        const float dl = sd[(TC_K - 1) * TC_EPI_THREADS], dw = sd[wpos * TC_EPI_THREADS];
        const int jl = sj[(TC_K - 1) * TC_EPI_THREADS], jw = sj[wpos * TC_EPI_THREADS];
        sd[wpos * TC_EPI_THREADS] = dl;
        sj[wpos * TC_EPI_THREADS] = jl;
        sd[(TC_K - 1) * TC_EPI_THREADS] = dw;
        sj[(TC_K - 1) * TC_EPI_THREADS] = jw;
      }
#endif
#pragma unroll
      for (int i = 0; i < TC_K; ++i) {
        od[i] = sd[i * TC_EPI_THREADS];
        oj[i] = sj[i * TC_EPI_THREADS];
      }
      od[TC_K] = bound;
      oj[TC_K] = -1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it / write into it
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <bool TF32>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_candidates_kernel(const __grid_constant__ TcArgs P) {
  tc_candidates_body<TF32, false>(P);
}
template <bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
tc_candidates_pair_kernel(const __grid_constant__ TcArgs P) {
  tc_candidates_body<TF32, true>(P);
}

// ------------------------------------------------------------------------------------------ rescore
// A block of 8 warps takes 8 A rows.  (1) One warp per row reads its candidate lists, forms the bounds and decides which
// candidates can still win (usually one to three of 18).  (2) The surviving (row, candidate) pairs of the block are
// pooled in shared memory and ONE THREAD PER PAIR takes the exact distance (FLANN's sequential float sum, no FMA:
// 352 dependent additions - with a lane per candidate of its own row a warp ran that loop for two or three busy
// lanes); the best (d2, j) key per row is kept with a shared-memory atomicMin.  (3) Lane 0 of each row's warp checks
// the certificate.
#ifndef PFX_RS_UNROLL
#define PFX_RS_UNROLL 1  // (measured: 1 -> 0.170 ms, 4 / 8 / 16 -> 0.211 ms at 65536 rows x 352: the sum is a dependent chain either way, and the unrolled loop holds more registers)
#endif
constexpr int RS_UNROLL = PFX_RS_UNROLL;  // float4 pairs of the exact distance loop in flight per thread
#ifndef PFX_RS_ROWS
#define PFX_RS_ROWS 8
#endif
constexpr int RS_ROWS = PFX_RS_ROWS, RS_MAXPAIRS = 64 * RS_ROWS;

__device__ __forceinline__ float tc_exact_d2(const float* __restrict__ a, const float* __restrict__ b, int dim) {
  float acc = 0.f;
  int d = 0;
  // 16-byte loads when both rows allow it (the sum stays sequential over d: FLANN's order)
  if ((((size_t)a | (size_t)b) & 15) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll RS_UNROLL
    for (; d + 4 <= dim; d += 4) {
      const float4 av = a4[d >> 2], bv = b4[d >> 2];
      const float d0 = __fsub_rn(av.x, bv.x), d1 = __fsub_rn(av.y, bv.y), d2_ = __fsub_rn(av.z, bv.z), d3 = __fsub_rn(av.w, bv.w);
      acc = __fadd_rn(acc, __fmul_rn(d0, d0));
      acc = __fadd_rn(acc, __fmul_rn(d1, d1));
      acc = __fadd_rn(acc, __fmul_rn(d2_, d2_));
      acc = __fadd_rn(acc, __fmul_rn(d3, d3));
    }
  }
  for (; d < dim; ++d) {
    const float df = __fsub_rn(a[d], b[d]);
    acc = __fadd_rn(acc, __fmul_rn(df, df));
  }
  return acc;
}

__global__ void __launch_bounds__(RS_ROWS * 32)
tc_rescore_kernel(const float* __restrict__ A, int na, int lda, const float* __restrict__ B, int nb, int ldb, int dim,
                  const float* __restrict__ cand_d, const int* __restrict__ cand_j, int nlists /* nsplit * TC_LISTS */,
                  const float* __restrict__ norm_a, const float* __restrict__ err_a,
                  const unsigned* __restrict__ maxima_b, int* __restrict__ nn_idx, float* __restrict__ nn_d2,
                  int* __restrict__ redo_list, int* __restrict__ redo_count) {
  __shared__ unsigned long long s_best[RS_ROWS];
  __shared__ int s_pair_j[RS_MAXPAIRS];
  __shared__ unsigned char s_pair_row[RS_MAXPAIRS];
  __shared__ int s_npairs;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * RS_ROWS + wid;
  if (threadIdx.x == 0) s_npairs = 0;
  if (lane == 0) s_best[wid] = 0xffffffffffffffffull;
  __syncthreads();
  const bool row_ok = i < na;
  const float nai = row_ok ? norm_a[i] : TC_BIG;
  const bool finite_row = row_ok && nai < 0.5f * TC_BIG;  // a non-finite query row is never matched
  const float eb_max = __uint_as_float(maxima_b[0]), nb_max = __uint_as_float(maxima_b[1]);
  const float eps = (finite_row ? err_a[i] : 0.f) + eb_max;
  // slack of the approximate squared distances themselves: fp32 accumulation of `dim` exact bf16 products,
  // the cancellation in na + nb - 2 dot, and the column number stored in the 7 low mantissa bits of a key
  const float delta2 = 2.4e-7f * (float)dim * sqrtf(nai * nb_max) + 2e-6f * (nai + nb_max);
  const float REL = 2e-5f;
  const int ncand = nlists * TC_KS;
  float kth = CUDART_INF_F;
  if (finite_row) {
    const float* cd = cand_d + (size_t)i * ncand;
    const int* cj = cand_j + (size_t)i * ncand;
    // outside bound: per list, the bound slot (keys never offered) and, when the list is full, its last entry
    // (keys offered but rejected or evicted are >= it)
    float dmin = CUDART_INF_F;
    for (int c = lane; c < ncand; c += 32) {
      const int slot = c % TC_KS;
      const float d = cd[c];
      if (slot == TC_K) kth = fminf(kth, d);
      else if (cj[c] >= 0) {
        if (slot == TC_K - 1) kth = fminf(kth, d);
        dmin = fminf(dmin, d);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kth = fminf(kth, __shfl_xor_sync(FULL, kth, o));
      dmin = fminf(dmin, __shfl_xor_sync(FULL, dmin, o));
    }
    // a candidate can only win if its lower bound is below the upper bound of the approximate best
    const float win_hi = sqrtf(fmaxf(dmin + delta2 + REL * fabsf(dmin), 0.f)) + eps;
    for (int c = lane; c < ncand; c += 32) {
      if ((c % TC_KS) == TC_K) continue;
      const int j = cj[c];
      const float dc = cd[c];
      if (j >= 0 && j < nb && dc < 0.5f * TC_BIG) {
        const float lo = sqrtf(fmaxf(dc - delta2 - REL * fabsf(dc), 0.f)) - eps;
        if (lo <= win_hi) {
          const int slot = atomicAdd(&s_npairs, 1);
          if (slot < RS_MAXPAIRS) {
            s_pair_j[slot] = j;
            s_pair_row[slot] = (unsigned char)wid;
          } else {  // (more survivors than the pool holds: take this one here)
            const float acc = tc_exact_d2(A + (size_t)i * lda, B + (size_t)j * ldb, dim);
            atomicMin(&s_best[wid], ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j);
          }
        }
      }
    }
  }
  __syncthreads();
  const int npairs = min(s_npairs, RS_MAXPAIRS);
  for (int t = threadIdx.x; t < npairs; t += RS_ROWS * 32) {
    const int rw = s_pair_row[t], j = s_pair_j[t];
    const float acc = tc_exact_d2(A + (size_t)(blockIdx.x * RS_ROWS + rw) * lda, B + (size_t)j * ldb, dim);
    atomicMin(&s_best[rw], ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)j);
  }
  __syncthreads();
  if (lane == 0 && row_ok) {
    if (!finite_row) {
      nn_idx[i] = -1;
      if (nn_d2) nn_d2[i] = CUDART_INF_F;
      return;
    }
    const unsigned long long best = s_best[wid];
    bool certified = false;
    int j = -1;
    float d2 = CUDART_INF_F;
    const bool outside_any = kth < 0.5f * TC_BIG;  // some valid column is not in the lists
    if (best != 0xffffffffffffffffull) {
      j = (int)(unsigned)(best & 0xffffffffull);
      d2 = __uint_as_float((unsigned)(best >> 32));
      // every column outside the lists has true distance >= sqrt(kth - slack) - eps
      const float outside_lo =
          outside_any ? sqrtf(fmaxf(kth - delta2 - REL * fabsf(kth), 0.f)) - eps : CUDART_INF_F;
      certified = sqrtf(d2) * 1.000001f < outside_lo;
    } else {
      certified = !outside_any;  // no finite target at all
    }
    nn_idx[i] = j;
    if (nn_d2) nn_d2[i] = d2;
    if (!certified) redo_list[atomicAdd(redo_count, 1)] = i;
  }
}

__global__ void tc_gather_rows_kernel(const float* __restrict__ A, int lda, int dim, const int* __restrict__ rows,
                                      int nrows, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nrows) return;
  const float* src = A + (size_t)rows[r] * lda;
  for (int d = lane; d < dim; d += 32) out[(size_t)r * dim + d] = src[d];
}
__global__ void tc_scatter_kernel(const int* __restrict__ rows, int nrows, const int* __restrict__ idx,
                                  const float* __restrict__ d2, int* __restrict__ nn_idx, float* __restrict__ nn_d2) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  nn_idx[rows[r]] = idx[r];
  if (nn_d2) nn_d2[rows[r]] = d2[r];
}

// ------------------------------------------------------------------------------------------ host
// tf32 operands (10-bit mantissa: 8x tighter candidate distances than bf16, so fewer rows need the exact redo)
// whenever the resident A tile still leaves room for the B ring; bf16 for long descriptors (SHOT352).  The epilogue
// paces the kernel for short descriptors, so the tf32 MMAs (half the bf16 rate) cost nothing there.
static bool tc_use_tf32(int dim) { return ((dim + 7) & ~7) * 4 * TC_TILE <= 96 * 1024; }

constexpr size_t TC_FIXED_SMEM = 512 * sizeof(float) + 2 * TC_K * TC_EPI_THREADS * 4 + 32 * 8 + 64;
constexpr size_t TC_SMEM_BUDGET = 225 * 1024;

// does the resident A tile of a dim-wide descriptor leave room for at least two ring stages?  (bf16 rows up to
// 696 elements; SHOT1344 and USC1980 do not fit and are matched by the exact fp32 scan, which gives the same bits)
bool match_tc_fits(int dim) {
  const bool tf32 = tc_use_tf32(dim);
  const int dpad = tf32 ? ((dim + 7) & ~7) : ((dim + 15) & ~15);
  const size_t nslab = tf32 ? dpad / 4 : dpad / 8;
  return nslab * TC_SLAB + TC_FIXED_SMEM + 2 * (size_t)TC_STAGE <= TC_SMEM_BUDGET;
}

static int tc_prepare(Ctx* ctx, TcOperand& op, const float* x, int n, int ld, int dim) {
  const bool tf32 = tc_use_tf32(dim);
  op.n = n;
  op.tf32 = tf32;
  op.dpad = tf32 ? ((dim + 7) & ~7) : ((dim + 15) & ~15);
  op.npad = std::max(1, div_up(n, TC_TILE)) * TC_TILE;
  PFX_CUDA(op.tiles.ensure((size_t)op.npad * op.dpad * (tf32 ? 4 : 2)));
  PFX_CUDA(op.norm.ensure((size_t)(op.npad + 256) * sizeof(float)));
  PFX_CUDA(op.err.ensure((size_t)op.npad * sizeof(float)));
  PFX_CUDA(op.maxima.ensure(16));
  PFX_CUDA(cudaMemsetAsync(op.maxima.p, 0, 16, ctx->stream));
  if (tf32)
    PFX_LAUNCH(ctx, tc_prep_kernel<true>, div_up(op.npad, 8), 256, 0, x, n, ld, dim, op.dpad, op.npad,
               op.tiles.as<unsigned char>(), op.norm.as<float>(), op.err.as<float>(), op.maxima.as<unsigned>());
  else
    PFX_LAUNCH(ctx, tc_prep_kernel<false>, div_up(op.npad, 8), 256, 0, x, n, ld, dim, op.dpad, op.npad,
               op.tiles.as<unsigned char>(), op.norm.as<float>(), op.err.as<float>(), op.maxima.as<unsigned>());
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// 1-NN of every row of a among the rows of b.  prepared: bit 0 / bit 1 = operand slot 0 / 1 already holds
// a / b (reciprocal matching prepares each matrix once).  slot_a selects which slot plays the A role.
static int tc_run(Ctx* ctx, int slot_a, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim,
                  int* nn_idx, float* nn_d2) {
  TcOperand& A = ctx->tc_ops[slot_a];
  TcOperand& B = ctx->tc_ops[slot_a ^ 1];
  const int n_at = A.npad / TC_TILE, n_bt = B.npad / TC_TILE;
  const int npairs = (n_bt + 1) / 2;
  // Long units: the running top-K settles after a few hundred columns (insertions decay like K / n), so a
  // unit should own as many columns as possible.  B is split (a) to give idle SMs work when there are fewer A
  // tiles than SMs and (b) into 2-4 parts when that fills the last wave of CTAs better (512 A tiles on 148 SMs:
  // 3.46 waves as whole units, 6.92 as halves), as long as a unit keeps >= 32 column pairs.
  // CTA pairs: a unit is two A tiles on two SMs
  const bool pair = ctx->tc_pair != 0 && n_at >= 2;
  const int n_au = pair ? div_up(n_at, 2) : n_at, n_smu = pair ? ctx->sm_count / 2 : ctx->sm_count;
  int nsplit = std::max(1, std::min(npairs, n_smu / std::max(n_au, 1)));
  if (n_au >= n_smu) {
    double best_eff = 0.0;
    for (int cand = 1; cand <= 4; ++cand) {
      if (cand > 1 && npairs / cand < 32) break;
      const double waves = (double)n_au * cand / n_smu;
      const double eff = waves / std::ceil(waves);
      if (eff > best_eff + 0.02) {
        best_eff = eff;
        nsplit = cand;
      }
    }
  }
  int pps = div_up(npairs, nsplit);
  nsplit = div_up(npairs, pps);
  const int nslab = A.tf32 ? A.dpad / 4 : A.dpad / 8;
  const size_t a_bytes = (size_t)nslab * TC_SLAB;
  const size_t fixed = a_bytes + TC_FIXED_SMEM;
  // signed arithmetic: a resident A tile beyond the budget must give nstage < 2, not a wrapped size_t
  const long long room = (long long)TC_SMEM_BUDGET - (long long)fixed;
  const long long stage_bytes = pair ? TC_STAGE / 2 : TC_STAGE;  // a CTA of a pair holds one B tile of a stage
  const int nstage = room < 0 ? 0 : (int)std::min<long long>(pair ? 12 : 8, room / stage_bytes);
  if (nstage < 2) return ctx->fail(PFX_E_INVALID, "tensor-core matching: descriptor dimension too large for one A tile");
  const size_t smem = std::max<size_t>(fixed + (size_t)nstage * stage_bytes, 120 * 1024);
  const int nlists = nsplit * TC_LISTS;
  const int ncand = nlists * TC_KS;
  PFX_CUDA(ctx->tc_cand_d.ensure((size_t)A.npad * ncand * sizeof(float)));
  PFX_CUDA(ctx->tc_cand_j.ensure((size_t)A.npad * ncand * sizeof(int)));
  PFX_CUDA(ctx->tc_redo.ensure(((size_t)na + 16) * sizeof(int)));
  int* redo_count = ctx->tc_redo.as<int>();
  int* redo_list = redo_count + 16;
  PFX_CUDA(cudaMemsetAsync(redo_count, 0, 16 * sizeof(int), ctx->stream));
  if (!ctx->smem_attr_match_tc) {
    PFX_CUDA(cudaFuncSetAttribute(tc_candidates_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PFX_CUDA(cudaFuncSetAttribute(tc_candidates_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PFX_CUDA(cudaFuncSetAttribute(tc_candidates_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    PFX_CUDA(cudaFuncSetAttribute(tc_candidates_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ctx->smem_attr_match_tc = true;
  }
  TcArgs P;
  P.At = A.tiles.as<unsigned char>();
  P.Bt = B.tiles.as<unsigned char>();
  P.na = A.norm.as<float>();
  P.nb = B.norm.as<float>();
  P.n_btiles = n_bt;
  P.nslab = nslab;
  P.nsplit = nsplit;
  P.pairs_per_split = pps;
  P.nstage = nstage;
  P.n_atiles = n_at;
  P.cand_d = ctx->tc_cand_d.as<float>();
  P.cand_j = ctx->tc_cand_j.as<int>();
  if (pair) {
    if (A.tf32)
      PFX_LAUNCH(ctx, tc_candidates_pair_kernel<true>, 2 * n_au * nsplit, TC_THREADS, smem, P);
    else
      PFX_LAUNCH(ctx, tc_candidates_pair_kernel<false>, 2 * n_au * nsplit, TC_THREADS, smem, P);
  } else if (A.tf32)
    PFX_LAUNCH(ctx, tc_candidates_kernel<true>, n_at * nsplit, TC_THREADS, smem, P);
  else
    PFX_LAUNCH(ctx, tc_candidates_kernel<false>, n_at * nsplit, TC_THREADS, smem, P);
  PFX_LAUNCH(ctx, tc_rescore_kernel, div_up(na, RS_ROWS), RS_ROWS * 32, 0, a, na, lda, b, nb, ldb, dim, P.cand_d, P.cand_j, nlists,
             A.norm.as<float>(), A.err.as<float>(), B.maxima.as<unsigned>(), nn_idx, nn_d2, redo_list, redo_count);
  PFX_CUDA(cudaGetLastError());
  int redo = 0;
  PFX_CUDA(cudaMemcpyAsync(&redo, redo_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->match_rows += na;
  ctx->match_redo += redo;
  ctx->match_tc_calls++;
  if (redo > 0) {  // exact all-pairs scan of the rows without a certificate
    PFX_CUDA(ctx->tc_rows.ensure((size_t)redo * dim * sizeof(float)));
    PFX_CUDA(ctx->tc_res.ensure((size_t)redo * (sizeof(int) + sizeof(float))));
    float* ga = ctx->tc_rows.as<float>();
    int* ridx = ctx->tc_res.as<int>();
    float* rd2 = reinterpret_cast<float*>(ridx + redo);
    PFX_LAUNCH(ctx, tc_gather_rows_kernel, div_up(redo, 8), 256, 0, a, lda, dim, redo_list, redo, ga);
    PFX_TRY(match_nn_exact(ctx, ga, redo, dim, b, nb, ldb, dim, ridx, rd2));
    PFX_LAUNCH(ctx, tc_scatter_kernel, div_up(redo, 256), 256, 0, redo_list, redo, ridx, rd2, nn_idx, nn_d2);
    PFX_CUDA(cudaGetLastError());
  }
  return 0;
}

int match_nn_tc(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* nn_idx,
                float* nn_d2) {
  if (na == 0) return 0;
  PFX_TRY(tc_prepare(ctx, ctx->tc_ops[0], a, na, lda, dim));
  PFX_TRY(tc_prepare(ctx, ctx->tc_ops[1], b, nb, ldb, dim));
  return tc_run(ctx, 0, a, na, lda, b, nb, ldb, dim, nn_idx, nn_d2);
}

// both directions with one preparation of each matrix (Features<T>::findCorrespondences, features.h:232-237)
int match_pair_tc(Ctx* ctx, const float* a, int na, int lda, const float* b, int nb, int ldb, int dim, int* s2t,
                  float* sd2, int* t2s, float* td2) {
  PFX_TRY(tc_prepare(ctx, ctx->tc_ops[0], a, na, lda, dim));
  PFX_TRY(tc_prepare(ctx, ctx->tc_ops[1], b, nb, ldb, dim));
  PFX_TRY(tc_run(ctx, 0, a, na, lda, b, nb, ldb, dim, s2t, sd2));
  if (t2s) PFX_TRY(tc_run(ctx, 1, b, nb, ldb, a, na, lda, dim, t2s, td2));
  return 0;
}

void match_tc_release(Ctx* ctx) {
  for (TcOperand& o : ctx->tc_ops) {
    o.tiles.release();
    o.norm.release();
    o.err.release();
    o.maxima.release();
  }
  for (DevBuf* b : {&ctx->tc_cand_d, &ctx->tc_cand_j, &ctx->tc_redo, &ctx->tc_rows, &ctx->tc_res}) b->release();
}

}  // namespace pfx

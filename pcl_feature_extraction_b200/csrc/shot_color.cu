// shot_color.cu — SHOT1344 = SHOT shape + CIELab colour channel (replaces
// pcl::SHOTColorEstimation<PointXYZRGB, Normal, SHOT1344>::compute as instantiated at reference
// evaluation.cpp:786-805 and driven through features.h:181-195).
//
// lab_kernel: packed 0x00RRGGBB -> normalised CIELab (L / 100, a / 120, b / 120) once per point, through the two
// lookup tables PCL builds with powf (sRGB gamma, 256 entries; XYZ cube root, 4000 entries); the tables are
// computed on the host with the same libm calls and uploaded, the float arithmetic around them is written out
// without FMA, so the Lab triplets are bit-identical to the CPU's.
// shot_color_kernel: one warp per query, lane per neighbour, like shot_kernel; the 352 shape slots and the
// 32 x 31 colour slots live in one int32 fixed-point shared-memory histogram (order-independent, reproducible).
// Every neighbour casts SHOT352's quadrilinear votes twice - at its cosine step in the shape channel and at its
// colour-distance step in the colour channel - with the same spatial weights (interpolateDoubleChannel); one L2
// normalisation over the 1344 slots.  Frames are SHOT's (shot_lrf_compute) or the caller's.
#include "internal.h"

namespace pfx {

constexpr int SCW = 4;  // warps per block
constexpr int SC_LEN = 1344, SC_SHAPE = 352;

struct LabTab {
  float srgb[256];
  float sxyz[4000];
};

__device__ __forceinline__ int lab_lut_index(float v) {
  const int i = (int)__fmul_rn(v, 4000.f);
  return min(max(i, 0), 3999);  // upstream reads one element past the table for y = 1.0 (white): clamped
}

__global__ void lab_kernel(const unsigned char* __restrict__ rgb, size_t stride, int n, const LabTab* __restrict__ T,
                           float4* __restrict__ lab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t c = *reinterpret_cast<const uint32_t*>(rgb + (size_t)i * stride);
  const float fr = T->srgb[(c >> 16) & 0xff], fg = T->srgb[(c >> 8) & 0xff], fb = T->srgb[c & 0xff];
  const float x = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.412453f), __fmul_rn(fg, 0.357580f)), __fmul_rn(fb, 0.180423f));
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.212671f), __fmul_rn(fg, 0.715160f)), __fmul_rn(fb, 0.072169f));
  const float z = __fadd_rn(__fadd_rn(__fmul_rn(fr, 0.019334f), __fmul_rn(fg, 0.119193f)), __fmul_rn(fb, 0.950227f));
  const float vx = T->sxyz[lab_lut_index(__fdiv_rn(x, 0.95047f))];
  const float vy = T->sxyz[lab_lut_index(y)];
  const float vz = T->sxyz[lab_lut_index(__fdiv_rn(z, 1.08883f))];
  float L = __fsub_rn(__fmul_rn(116.0f, vy), 16.0f);
  if (L > 100.f) L = 100.0f;
  float A = __fmul_rn(500.0f, __fsub_rn(vx, vy));
  A = fminf(fmaxf(A, -120.f), 120.f);
  float B = __fmul_rn(200.0f, __fsub_rn(vy, vz));
  B = fminf(fmaxf(B, -120.f), 120.f);
  lab[i] = make_float4(__fdiv_rn(L, 100.0f), __fdiv_rn(A, 120.0f), __fdiv_rn(B, 120.0f), 0.f);
}

__device__ __forceinline__ void sc_add(int* h, int slot, double v, float scale) {
  atomicAdd(&h[slot], __double2int_rn(v * (double)scale));
}

template <bool DENSE>
__global__ void __launch_bounds__(SCW * 32)
shot_color_kernel(GridDev g, const float4* __restrict__ queries, int nq, const float4* __restrict__ nrm,
                  const float4* __restrict__ lab_orig, const float4* __restrict__ qlab, float r2, double R,
                  const float* __restrict__ rf9, float* __restrict__ out, size_t stride) {
  __shared__ int hist[SCW][SC_LEN];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int qi = blockIdx.x * SCW + wid;
  if (qi >= nq) return;
  int* h = hist[wid];
  const int n_valid = g.gp->n_valid;
  for (int c = lane; c < SC_LEN; c += 32) h[c] = 0;
  __syncwarp();
  const float4 q = DENSE ? g.pts[qi] : queries[qi];
  const size_t row = DENSE ? (size_t)__float_as_int(q.w) : (size_t)qi;
  float* o = out + row * stride;
  const float nanv = __int_as_float(0x7fc00000);
  float rf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) rf[i] = rf9[row * 9 + i];
  const bool ok = finite3(q.x, q.y, q.z) && (!DENSE || qi < n_valid) && isfinite(rf[0]) && isfinite(rf[3]) && isfinite(rf[6]);
  int n_nb = 0;
  CellBlock blk;
  blk.total = 0;
  if (ok) {
    blk = DENSE ? stencil_of_point(g, qi, lane) : stencil_of_pos(g, q.x, q.y, q.z, lane);
    for (int base = 0; base < blk.total; base += 32) {
      const int c = base + lane;
      bool valid = c < blk.total;
      const int j = block_candidate(blk, valid ? c : 0);
      if (valid) {
        const float4 p = g.pts[j];
        valid = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
      }
      n_nb += __popc(__ballot_sync(FULL, valid));
    }
  }
  if (!ok || n_nb == 0) {
    for (int c = lane; c < SC_LEN; c += 32) o[c] = nanv;
    if (lane < 9) o[SC_LEN + lane] = nanv;
    return;
  }
  if (lane < 9) o[SC_LEN + lane] = rf[lane];
  if (n_nb < 5) {  // computePointSHOT: too few neighbours -> NaN descriptor, frame kept
    for (int c = lane; c < SC_LEN; c += 32) o[c] = nanv;
    return;
  }
  const int bits = 32 - __clz(5 * n_nb + 8);  // every slot receives at most 4.5 per neighbour
  const float scale = exp2f((float)(30 - bits));
  const double r12 = R / 2, r14 = R / 4, r34 = 3 * R / 4;
  const float RAD45 = 0.78539816339744830962f, RAD90 = 1.57079632679489661923f, RAD135 = 2.35619449019234492885f,
              RAD_PI_7_8 = 2.7488935718910690836f;
  const float4 ref = DENSE ? lab_orig[row] : qlab[qi];
  for (int base = 0; base < blk.total; base += 32) {
    const int c = base + lane;
    const bool valid = c < blk.total;
    const int j = block_candidate(blk, valid ? c : 0);
    if (!valid) continue;
    const float4 p = g.pts[j];
    const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
    if (!(d2 < r2)) continue;
    const float4 nj = nrm[j];
    if (!finite3(nj.x, nj.y, nj.z)) continue;
    double cosd = (double)__fadd_rn(__fadd_rn(__fmul_rn(nj.x, rf[6]), __fmul_rn(nj.y, rf[7])), __fmul_rn(nj.z, rf[8]));
    cosd = fmin(1.0, fmax(-1.0, cosd));
    double bds = ((1.0 + cosd) * 10) / 2;
    const float4 lb = lab_orig[__float_as_int(p.w)];
    double cd = (fabs((double)__fsub_rn(ref.x, lb.x)) +
                 ((fabs((double)__fsub_rn(ref.y, lb.y)) + fabs((double)__fsub_rn(ref.z, lb.z))) / 2)) / 3;
    cd = fmin(1.0, fmax(0.0, cd));
    double bdc = cd * 30;
    const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
    const double dist = sqrt((double)d2);
    if (fabs(dist) < 1e-15) continue;
    double x = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[0]), __fmul_rn(dy, rf[1])), __fmul_rn(dz, rf[2]));
    double y = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[3]), __fmul_rn(dy, rf[4])), __fmul_rn(dz, rf[5]));
    double z = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[6]), __fmul_rn(dy, rf[7])), __fmul_rn(dz, rf[8]));
    if (fabs(y) < 1e-30) y = 0;
    if (fabs(x) < 1e-30) x = 0;
    if (fabs(z) < 1e-30) z = 0;
    const int bit4 = ((y > 0) || ((y == 0.0) && (x < 0))) ? 1 : 0;
    const int bit3 = ((x > 0) || ((x == 0.0) && (y > 0))) ? !bit4 : bit4;
    int di = ((bit4 << 3) + (bit3 << 2)) << 1;
    if ((x * y > 0) || (x == 0.0))
      di += (fabs(x) >= fabs(y)) ? 0 : 4;
    else
      di += (fabs(x) > fabs(y)) ? 4 : 0;
    di += z > 0 ? 1 : 0;
    di += (dist > r12) ? 2 : 0;
    const int sts = (int)floor(bds + 0.5), stc = (int)floor(bdc + 0.5);
    const int vs = di * 11, vc = SC_SHAPE + di * 31;
    bds -= sts;
    bdc -= stc;
    const double ws = 1 - fabs(bds), wc = 1 - fabs(bdc);
    if (bds > 0)
      sc_add(h, vs + ((sts + 1) % 10), bds, scale);
    else
      sc_add(h, vs + ((sts + 9) % 10), -bds, scale);
    if (bdc > 0)
      sc_add(h, vc + ((stc + 1) % 30), bdc, scale);
    else
      sc_add(h, vc + ((stc + 29) % 30), -bdc, scale);
    auto both = [&](int vol, double v) {  // the same spatial weight goes to both channels
      sc_add(h, vol * 11 + sts, v, scale);
      sc_add(h, SC_SHAPE + vol * 31 + stc, v, scale);
    };
    double w = 0;
    if (dist > r12) {
      const double rd = (dist - r34) / r12;
      if (dist > r34)
        w += 1 - rd;
      else {
        w += 1 + rd;
        both(di - 2, -rd);
      }
    } else {
      const double rd = (dist - r14) / r12;
      if (dist < r14)
        w += 1 + rd;
      else {
        w += 1 - rd;
        both(di + 2, rd);
      }
    }
    const float ic = (float)fmin(1.0, fmax(-1.0, z / dist));
    const float inc = acosf(ic);
    if (z <= 0) {  // == (inc > 90deg || (|inc - 90deg| < 1e-30 && z <= 0)) in exact arithmetic
      const float e = (inc - RAD135) / RAD90;
      if (inc > RAD135)
        w += 1 - e;
      else {
        w += 1 + e;
        both(di + 1, -e);
      }
    } else {
      const float e = (inc - RAD45) / RAD90;
      if (inc < RAD45)
        w += 1 + e;
      else {
        w += 1 - e;
        both(di - 1, e);
      }
    }
    if (y != 0.0 || x != 0.0) {
      const float az = atan2f((float)y, (float)x);
      const int sel = di >> 2;
      float ad = (az - (-RAD_PI_7_8 + RAD45 * sel)) / RAD45;
      ad = fmaxf(-0.5f, fminf(ad, 0.5f));
      if (ad > 0) {
        w += 1 - ad;
        both((di + 4) % 32, ad);
      } else {
        w += 1 + ad;
        both((di + 28) % 32, -ad);
      }
    }
    sc_add(h, vs + sts, ws + w, scale);
    sc_add(h, vc + stc, wc + w, scale);
  }
  __syncwarp();
  const float inv_scale = 1.0f / scale;
  double acc = 0.0;
  for (int c = lane; c < SC_LEN; c += 32) {
    const float v = (float)h[c] * inv_scale;
    acc += (double)__fmul_rn(v, v);
  }
  acc = warp_sum(acc);
  const float nrmv = (float)sqrt(acc);
  for (int c = lane; c < SC_LEN; c += 32) o[c] = __fdiv_rn((float)h[c] * inv_scale, nrmv);
}

static int lab_tables(Ctx* ctx) {
  if (ctx->lab_tab.p) return 0;
  LabTab host;  // PCL's two lookup tables, same libm calls (features/impl/shot.hpp, RGB2CIELAB)
  for (int i = 0; i < 256; ++i) {
    float f = static_cast<float>(i) / 255.0f;
    host.srgb[i] = (f > 0.04045) ? powf((f + 0.055f) / 1.055f, 2.4f) : f / 12.92f;
  }
  for (int i = 0; i < 4000; ++i) {
    float f = static_cast<float>(i) / 4000.0f;
    host.sxyz[i] = (f > 0.008856) ? static_cast<float>(powf(f, 0.3333f)) : static_cast<float>((7.787 * f) + (16.0 / 116.0));
  }
  PFX_CUDA(ctx->lab_tab.ensure(sizeof(LabTab)));
  PFX_CUDA(cudaMemcpyAsync(ctx->lab_tab.p, &host, sizeof(LabTab), cudaMemcpyHostToDevice, ctx->stream));
  PFX_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// packed colours (device pointer, byte stride) -> normalised Lab float4 rows
int colors_to_lab(Ctx* ctx, const unsigned char* rgb_dev, size_t stride, int n, DevBuf& lab) {
  PFX_TRY(lab_tables(ctx));
  PFX_CUDA(lab.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  if (n) PFX_LAUNCH(ctx, lab_kernel, div_up(n, 256), 256, 0, rgb_dev, stride, n, ctx->lab_tab.as<LabTab>(), lab.as<float4>());
  PFX_CUDA(cudaGetLastError());
  return 0;
}

// out_dev: rows of 1344 + 9 floats at stride_floats, caller query order
int shot_color_compute(Ctx* ctx, Grid* g, double radius, const float* rf9_dev, float* out_dev, size_t stride_floats) {
  const int nq = (int)ctx->num_queries();
  if (nq == 0) return 0;
  const float r2 = (float)(radius * radius);
  const float4* nrm = nullptr;
  PFX_TRY(normals_sorted_for_grid(ctx, g, &nrm));
  const int blocks = div_up(nq, SCW);
  if (ctx->q_is_surface)
    PFX_LAUNCH(ctx, shot_color_kernel<true>, blocks, SCW * 32, 0, g->view(), nullptr, nq, nrm, ctx->surf_lab.as<float4>(),
               nullptr, r2, radius, rf9_dev, out_dev, stride_floats);
  else
    PFX_LAUNCH(ctx, shot_color_kernel<false>, blocks, SCW * 32, 0, g->view(), ctx->qry.as<float4>(), nq, nrm,
               ctx->surf_lab.as<float4>(), ctx->qry_lab.as<float4>(), r2, radius, rf9_dev, out_dev, stride_floats);
  PFX_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace pfx

// shot_common.cuh — per-neighbour arithmetic of SHOT352 shared by the SHOT kernels
// (restates SHOTEstimation::createBinDistanceShape + interpolateSingleChannel).
#pragma once
#include "common.cuh"

namespace pfx {

// histogram slots are int32 fixed point (order-independent accumulation => bit-reproducible output)
__device__ __forceinline__ void shot_add(int* h, int slot, double v, float scale) {
  atomicAdd(&h[slot], __double2int_rn(v * (double)scale));
}

// power-of-two scale such that a bin that receives <= 4.5 per neighbour cannot overflow int32
__device__ __forceinline__ float shot_scale(int n_nb) {
  int bits = 32 - __clz(5 * n_nb + 8);
  return exp2f((float)(30 - bits));
}

// One neighbour p (squared distance d2, normal nj) of the query q with frame rf = (x, y, z axes).
// All DISCRETE decisions (volume index, cosine step, radial / elevation branch) replicate the CPU
// arithmetic exactly (float dots, double compares); only the continuous interpolation weights use
// float acosf / atan2f.
__device__ __forceinline__ void shot_accumulate_neighbor(int* h, float scale, float4 q, float4 p, float d2, float4 nj,
                                                         const float* rf, double R) {
  if (!finite3(nj.x, nj.y, nj.z)) return;
  const double r12 = R / 2, r14 = R / 4, r34 = 3 * R / 4;
  const float RAD45 = 0.78539816339744830962f, RAD90 = 1.57079632679489661923f, RAD135 = 2.35619449019234492885f,
              RAD_PI_7_8 = 2.7488935718910690836f;
  double cosd = (double)__fadd_rn(__fadd_rn(__fmul_rn(nj.x, rf[6]), __fmul_rn(nj.y, rf[7])), __fmul_rn(nj.z, rf[8]));
  cosd = fmin(1.0, fmax(-1.0, cosd));
  double bd = ((1.0 + cosd) * 10) / 2;
  float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
  double dist = sqrt((double)d2);
  if (fabs(dist) < 1e-15) return;
  double x = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[0]), __fmul_rn(dy, rf[1])), __fmul_rn(dz, rf[2]));
  double y = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[3]), __fmul_rn(dy, rf[4])), __fmul_rn(dz, rf[5]));
  double z = (double)__fadd_rn(__fadd_rn(__fmul_rn(dx, rf[6]), __fmul_rn(dy, rf[7])), __fmul_rn(dz, rf[8]));
  if (fabs(y) < 1e-30) y = 0;
  if (fabs(x) < 1e-30) x = 0;
  if (fabs(z) < 1e-30) z = 0;
  int bit4 = ((y > 0) || ((y == 0.0) && (x < 0))) ? 1 : 0;
  int bit3 = ((x > 0) || ((x == 0.0) && (y > 0))) ? !bit4 : bit4;
  int di = ((bit4 << 3) + (bit3 << 2)) << 1;
  if ((x * y > 0) || (x == 0.0))
    di += (fabs(x) >= fabs(y)) ? 0 : 4;
  else
    di += (fabs(x) > fabs(y)) ? 4 : 0;
  di += z > 0 ? 1 : 0;
  di += (dist > r12) ? 2 : 0;
  int step = (int)floor(bd + 0.5);
  int vol = di * 11;
  bd -= step;
  double w = 1 - fabs(bd);
  if (bd > 0)
    shot_add(h, vol + ((step + 1) % 10), bd, scale);
  else
    shot_add(h, vol + ((step - 1 + 10) % 10), -bd, scale);
  if (dist > r12) {
    double rd = (dist - r34) / r12;
    if (dist > r34)
      w += 1 - rd;
    else {
      w += 1 + rd;
      shot_add(h, (di - 2) * 11 + step, -rd, scale);
    }
  } else {
    double rd = (dist - r14) / r12;
    if (dist < r14)
      w += 1 + rd;
    else {
      w += 1 - rd;
      shot_add(h, (di + 2) * 11 + step, rd, scale);
    }
  }
  float ic = (float)fmin(1.0, fmax(-1.0, z / dist));
  float inc = acosf(ic);
  if (z <= 0) {  // == (inc > 90deg || (|inc - 90deg| < 1e-30 && z <= 0)) in exact arithmetic
    float e = (inc - RAD135) / RAD90;
    if (inc > RAD135)
      w += 1 - e;
    else {
      w += 1 + e;
      shot_add(h, (di + 1) * 11 + step, -e, scale);
    }
  } else {
    float e = (inc - RAD45) / RAD90;
    if (inc < RAD45)
      w += 1 + e;
    else {
      w += 1 - e;
      shot_add(h, (di - 1) * 11 + step, e, scale);
    }
  }
  if (y != 0.0 || x != 0.0) {
    float az = atan2f((float)y, (float)x);
    int sel = di >> 2;
    float ad = (az - (-RAD_PI_7_8 + RAD45 * sel)) / RAD45;
    ad = fmaxf(-0.5f, fminf(ad, 0.5f));
    if (ad > 0) {
      w += 1 - ad;
      shot_add(h, ((di + 4) % 32) * 11 + step, ad, scale);
    } else {
      w += 1 + ad;
      shot_add(h, ((di - 4 + 32) % 32) * 11 + step, -ad, scale);
    }
  }
  shot_add(h, vol + step, w, scale);
}

// Float variant of the above for the fused dense kernel.  The three frame projections keep the CPU's
// float arithmetic (separate multiply / add), so every sign / sector decision is the CPU's.  PCL adds the four
// interpolation weights of a neighbour into ITS bin (a sum, not a product): where the bin itself changes - the
// cosine step and the radial shell at R / 2 - the descriptor is discontinuous, and those two decisions are taken
// exactly like the CPU (the step in double, the shell from d2 against shot_d2_threshold).  The remaining branch
// points (R / 4, 3 R / 4, 45 / 135 degrees, the azimuth sector centre) are bin CENTRES, where the weight that
// leaves one bin enters its neighbour continuously: single precision and a 2-ulp square root are enough there.
// largest float f with sqrt((double)f) <= r: "sqrt((double)d2) > r" (the CPU's shell test) is then exactly "d2 > f"
__device__ __forceinline__ float shot_d2_threshold(double r) {
  float t = (float)(r * r);
  while (sqrt((double)t) > r) t = nextafterf(t, 0.f);
  while (sqrt((double)nextafterf(t, CUDART_INF_F)) <= r) t = nextafterf(t, CUDART_INF_F);
  return t;
}

__device__ __forceinline__ void shot_accumulate_neighbor_f(int* h, float scale, float4 q, float4 p, float d2, float4 nj,
                                                           const float* rf, float R, float t12) {
  if (!finite3(nj.x, nj.y, nj.z)) return;
  const float r14 = 0.25f * R, r34 = 0.75f * R, inv_r12 = 2.0f / R;
  const float RAD45 = 0.78539816339744830962f, RAD135 = 2.35619449019234492885f, RAD_PI_7_8 = 2.7488935718910690836f,
              INV_RAD90 = 0.63661977236758134308f, INV_RAD45 = 1.27323954473516268615f;
  float cosd = __fadd_rn(__fadd_rn(__fmul_rn(nj.x, rf[6]), __fmul_rn(nj.y, rf[7])), __fmul_rn(nj.z, rf[8]));
  cosd = fminf(1.0f, fmaxf(-1.0f, cosd));
  float bd = (1.0f + cosd) * 5.0f;
  const float dx = __fsub_rn(p.x, q.x), dy = __fsub_rn(p.y, q.y), dz = __fsub_rn(p.z, q.z);
  // sqrt through the reciprocal-square-root unit (2 ulp): distances only enter continuous weights and shell tests
  // that the interpolation makes continuous
  const float dist = (d2 > 0.f) ? d2 * rsqrtf(d2) : 0.f;
  if (dist < 1e-15f) return;
  float x = __fadd_rn(__fadd_rn(__fmul_rn(dx, rf[0]), __fmul_rn(dy, rf[1])), __fmul_rn(dz, rf[2]));
  float y = __fadd_rn(__fadd_rn(__fmul_rn(dx, rf[3]), __fmul_rn(dy, rf[4])), __fmul_rn(dz, rf[5]));
  float z = __fadd_rn(__fadd_rn(__fmul_rn(dx, rf[6]), __fmul_rn(dy, rf[7])), __fmul_rn(dz, rf[8]));
  if (fabsf(y) < 1e-30f) y = 0.f;
  if (fabsf(x) < 1e-30f) x = 0.f;
  if (fabsf(z) < 1e-30f) z = 0.f;
  const int bit4 = ((y > 0.f) || ((y == 0.f) && (x < 0.f))) ? 1 : 0;
  const int bit3 = ((x > 0.f) || ((x == 0.f) && (y > 0.f))) ? !bit4 : bit4;
  int di = ((bit4 << 3) + (bit3 << 2)) << 1;
  const bool same_sign = (x > 0.f && y > 0.f) || (x < 0.f && y < 0.f);  // x * y > 0 without underflow
  if (same_sign || (x == 0.f))
    di += (fabsf(x) >= fabsf(y)) ? 0 : 4;
  else
    di += (fabsf(x) > fabsf(y)) ? 4 : 0;
  di += z > 0.f ? 1 : 0;
  // the two decisions at which PCL's additive interpolation is NOT continuous (the neighbour's whole weight changes
  // bin) replicate the CPU exactly: the radial shell from d2 itself, the cosine step in double
  const bool outer = d2 > t12;
  di += outer ? 2 : 0;
  const int step = (int)floor(((1.0 + (double)cosd) * 10) / 2 + 0.5);
  const int vol = di * 11;
  bd -= (float)step;
  float w = 1.0f - fabsf(bd);
  {
    int nb_step = (bd > 0.f) ? step + 1 : step + 9;  // (step +- 1) mod 10, step in 0..10
    nb_step -= (nb_step >= 10) ? 10 : 0;
    atomicAdd(&h[vol + nb_step], __float2int_rn(fabsf(bd) * scale));
  }
  if (outer) {
    const float rd = (dist - r34) * inv_r12;
    if (dist > r34)
      w += 1.0f - rd;
    else {
      w += 1.0f + rd;
      atomicAdd(&h[(di - 2) * 11 + step], __float2int_rn(-rd * scale));
    }
  } else {
    const float rd = (dist - r14) * inv_r12;
    if (dist < r14)
      w += 1.0f + rd;
    else {
      w += 1.0f - rd;
      atomicAdd(&h[(di + 2) * 11 + step], __float2int_rn(rd * scale));
    }
  }
  const float rho2 = fmaf(x, x, y * y);
  const float rho = (rho2 > 0.f) ? rho2 * rsqrtf(rho2) : 0.f;  // inc = acos(z / dist) = atan2(|(x, y)|, z), in [0, pi]
  const float inc = fast_atan2f(rho, z);
  if (z <= 0.f) {
    const float e = (inc - RAD135) * INV_RAD90;
    if (inc > RAD135)
      w += 1.0f - e;
    else {
      w += 1.0f + e;
      atomicAdd(&h[(di + 1) * 11 + step], __float2int_rn(-e * scale));
    }
  } else {
    const float e = (inc - RAD45) * INV_RAD90;
    if (inc < RAD45)
      w += 1.0f + e;
    else {
      w += 1.0f - e;
      atomicAdd(&h[(di - 1) * 11 + step], __float2int_rn(e * scale));
    }
  }
  if (y != 0.f || x != 0.f) {
    const float az = fast_atan2f(y, x);
    const int sel = di >> 2;
    float ad = (az - (-RAD_PI_7_8 + RAD45 * (float)sel)) * INV_RAD45;
    ad = fmaxf(-0.5f, fminf(ad, 0.5f));
    if (ad > 0.f) {
      w += 1.0f - ad;
      atomicAdd(&h[((di + 4) & 31) * 11 + step], __float2int_rn(ad * scale));
    } else {
      w += 1.0f + ad;
      atomicAdd(&h[((di + 28) & 31) * 11 + step], __float2int_rn(-ad * scale));
    }
  }
  atomicAdd(&h[vol + step], __float2int_rn(w * scale));
}

// Eigenvectors of the largest (x) and smallest (z) eigenvalue of a symmetric 3x3 matrix given in double
// (a = xx xy xz yy yz zz).  Float Jacobi on the max-scaled matrix finds the vectors to ~1e-7; one double
// refinement step per vector (Rayleigh quotient, then the largest cross product of two rows of A - l I,
// pcl::eigen33's construction) brings them to ~1e-12 unless the eigenvalue is (nearly) repeated, where no
// two implementations agree anyway.  Returns false when the matrix is not finite.
__device__ __forceinline__ void refine_eigvec(const double c[6], double v[3]) {
  const double cx = c[0] * v[0] + c[1] * v[1] + c[2] * v[2];
  const double cy = c[1] * v[0] + c[3] * v[1] + c[4] * v[2];
  const double cz = c[2] * v[0] + c[4] * v[1] + c[5] * v[2];
  const double l = (v[0] * cx + v[1] * cy + v[2] * cz) / (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double r0[3] = {c[0] - l, c[1], c[2]}, r1[3] = {c[1], c[3] - l, c[4]}, r2[3] = {c[2], c[4], c[5] - l};
  const double e0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  const double e1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  const double e2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  const double l0 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2];
  const double l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
  const double l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
  double bx = e0[0], by = e0[1], bz = e0[2], bl = l0;
  if (l1 > bl) { bx = e1[0]; by = e1[1]; bz = e1[2]; bl = l1; }
  if (l2 > bl) { bx = e2[0]; by = e2[1]; bz = e2[2]; bl = l2; }
  if (bl > 1e-24) {  // (c is scaled to unit max-abs) otherwise rank < 2: keep the Jacobi vector
    const double il = rsqrt(bl);
    bx *= il; by *= il; bz *= il;
    if (bx * v[0] + by * v[1] + bz * v[2] < 0) { bx = -bx; by = -by; bz = -bz; }
    v[0] = bx; v[1] = by; v[2] = bz;
  }
}

__device__ __forceinline__ bool eig_extreme_refined(const double a[6], double x[3], double z[3]) {
  double sc = fmax(fmax(fabs(a[0]), fabs(a[1])), fmax(fmax(fabs(a[2]), fabs(a[3])), fmax(fabs(a[4]), fabs(a[5]))));
  if (!isfinite(sc)) return false;
  const double isc = (sc > 1e-300) ? 1.0 / sc : 1.0;
  double c[6];
  float af[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    c[i] = a[i] * isc;
    af[i] = (float)c[i];
  }
  float w[3], v[3][3];
  eig_sym3<float>(af, w, v, 8);
  x[0] = v[0][2]; x[1] = v[1][2]; x[2] = v[2][2];
  z[0] = v[0][0]; z[1] = v[1][0]; z[2] = v[2][0];
  refine_eigvec(c, x);
  refine_eigvec(c, z);
  return true;
}

// y = z x x in float (rf.row(1) = rf.row(2).cross(rf.row(0)))
__device__ __forceinline__ void lrf_to_float9(const double x[3], const double z[3], float* o) {
  float fx[3] = {(float)x[0], (float)x[1], (float)x[2]};
  float fz[3] = {(float)z[0], (float)z[1], (float)z[2]};
  o[0] = fx[0]; o[1] = fx[1]; o[2] = fx[2];
  o[6] = fz[0]; o[7] = fz[1]; o[8] = fz[2];
  o[3] = __fsub_rn(__fmul_rn(fz[1], fx[2]), __fmul_rn(fz[2], fx[1]));
  o[4] = __fsub_rn(__fmul_rn(fz[2], fx[0]), __fmul_rn(fz[0], fx[2]));
  o[5] = __fsub_rn(__fmul_rn(fz[0], fx[1]), __fmul_rn(fz[1], fx[0]));
}

}  // namespace pfx

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

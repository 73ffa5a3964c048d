// shot_color.cpp — CPU oracle: SHOT1344 (shape + colour).  TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// Restates pcl::SHOTColorEstimation<PointXYZRGB, Normal, SHOT1344> (reference evaluation.cpp:786-805 through
// features.h:181-195; upstream features/impl/shot.hpp: computePointSHOT, RGB2CIELAB, interpolateDoubleChannel).
// Per neighbour the shape channel is SHOT352's (cosine of the normal against the frame's z axis, 10 + 1 bins) and
// the colour channel bins the L1 distance of the CIELab colours, (|dL| + (|da| + |db|) / 2) / 3 with L / 100 and
// a, b / 120, into 30 + 1 bins; both are spread over the same 32 spatial volumes with the same quadrilinear
// weights; layout = 352 shape slots, then 32 x 31 colour slots; one L2 normalisation over all 1344.
// The frames come from the SHOT oracle (orc_shot_lrf).
//
// Definitions where upstream is undefined: the XYZ -> Lab table lookup int(v * 4000) is clamped to the table
// (upstream reads one element past it for white: y = 1.0); a neighbour with a non-finite normal is skipped.
#include "oracle_common.hpp"
#include "pcl_oracle.h"

using namespace orc;

namespace {

struct LabTables {
  float srgb[256];
  float sxyz[4000];
  LabTables() {
    for (int i = 0; i < 256; ++i) {
      float f = static_cast<float>(i) / 255.0f;
      if (f > 0.04045)
        srgb[i] = powf((f + 0.055f) / 1.055f, 2.4f);
      else
        srgb[i] = f / 12.92f;
    }
    for (int i = 0; i < 4000; ++i) {
      float f = static_cast<float>(i) / 4000.0f;
      if (f > 0.008856)
        sxyz[i] = static_cast<float>(powf(f, 0.3333f));
      else
        sxyz[i] = static_cast<float>((7.787 * f) + (16.0 / 116.0));
    }
  }
};
const LabTables& tables() {
  static LabTables t;
  return t;
}

inline int lutIndex(float v) {
  int i = (int)(v * 4000);
  return std::min(std::max(i, 0), 3999);
}

// normalised Lab: L / 100, a / 120, b / 120
void rgb2lab(uint32_t rgb, float lab[3]) {
  const LabTables& T = tables();
  const unsigned char R = (rgb >> 16) & 0xff, G = (rgb >> 8) & 0xff, B = rgb & 0xff;
  float fr = T.srgb[R], fg = T.srgb[G], fb = T.srgb[B];
  const float x = fr * 0.412453f + fg * 0.357580f + fb * 0.180423f;
  const float y = fr * 0.212671f + fg * 0.715160f + fb * 0.072169f;
  const float z = fr * 0.019334f + fg * 0.119193f + fb * 0.950227f;
  float vx = x / 0.95047f, vy = y, vz = z / 1.08883f;
  vx = T.sxyz[lutIndex(vx)];
  vy = T.sxyz[lutIndex(vy)];
  vz = T.sxyz[lutIndex(vz)];
  float L = 116.0f * vy - 16.0f;
  if (L > 100) L = 100.0f;
  float A = 500.0f * (vx - vy);
  if (A > 120) A = 120.0f; else if (A < -120) A = -120.0f;
  float Bv = 200.0f * (vy - vz);
  if (Bv > 120) Bv = 120.0f; else if (Bv < -120) Bv = -120.0f;
  lab[0] = L / 100.0f;
  lab[1] = A / 120.0f;
  lab[2] = Bv / 120.0f;
}

void shotColorRow(const float* surf, const uint32_t* rgb, const float* normals4, const float* c, uint32_t crgb,
                  const std::vector<Nbr>& nb, double R, const float rf[9], float* shot) {
  const int nbs = 10, nbc = 30, ss = nbs + 1, sc = nbc + 1, maxSectors = 32, stride = 32 * ss;
  const double r12 = R / 2, r14 = R / 4, r34 = 3 * R / 4;
  const double RAD45 = 0.78539816339744830961566084581988, RAD90 = 2 * RAD45, RAD135 = 3 * RAD45,
               RAD_PI_7_8 = 2.7488935718910690836548129603691;
  for (int i = 0; i < 1344; ++i) shot[i] = 0.f;
  const float *fx = rf, *fy = rf + 3, *fz = rf + 6;
  float ref[3];
  rgb2lab(crgb, ref);
  for (const Nbr& b : nb) {
    const float* nrm = normals4 + 4 * (size_t)b.idx;
    if (!std::isfinite(nrm[0]) || !std::isfinite(nrm[1]) || !std::isfinite(nrm[2])) continue;
    double cosd = (double)(nrm[0] * fz[0] + nrm[1] * fz[1] + nrm[2] * fz[2]);
    if (cosd > 1.0) cosd = 1.0;
    if (cosd < -1.0) cosd = -1.0;
    double bds = ((1.0 + cosd) * nbs) / 2;
    float lab[3];
    rgb2lab(rgb[b.idx], lab);
    double cd = (std::fabs((double)(ref[0] - lab[0])) + ((std::fabs((double)(ref[1] - lab[1])) + std::fabs((double)(ref[2] - lab[2]))) / 2)) / 3;
    if (cd > 1.0) cd = 1.0;
    if (cd < 0.0) cd = 0.0;
    double bdc = cd * nbc;

    const float* p = surf + 3 * (size_t)b.idx;
    float d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    double dist = std::sqrt((double)b.d2);
    if (std::fabs(dist) < 1e-15) continue;
    double x = (double)(d[0] * fx[0] + d[1] * fx[1] + d[2] * fx[2]);
    double y = (double)(d[0] * fy[0] + d[1] * fy[1] + d[2] * fy[2]);
    double z = (double)(d[0] * fz[0] + d[1] * fz[1] + d[2] * fz[2]);
    if (std::fabs(y) < 1e-30) y = 0;
    if (std::fabs(x) < 1e-30) x = 0;
    if (std::fabs(z) < 1e-30) z = 0;
    int bit4 = ((y > 0) || ((y == 0.0) && (x < 0))) ? 1 : 0;
    int bit3 = ((x > 0) || ((x == 0.0) && (y > 0))) ? !bit4 : bit4;
    int di = ((bit4 << 3) + (bit3 << 2)) << 1;
    if ((x * y > 0) || (x == 0.0))
      di += (std::fabs(x) >= std::fabs(y)) ? 0 : 4;
    else
      di += (std::fabs(x) > std::fabs(y)) ? 4 : 0;
    di += z > 0 ? 1 : 0;
    di += (dist > r12) ? 2 : 0;
    int sts = (int)std::floor(bds + 0.5), stc = (int)std::floor(bdc + 0.5);
    int vs = di * ss, vc = stride + di * sc;
    bds -= sts;
    bdc -= stc;
    double ws = 1 - std::fabs(bds), wc = 1 - std::fabs(bdc);
    if (bds > 0)
      shot[vs + ((sts + 1) % nbs)] += (float)bds;
    else
      shot[vs + ((sts - 1 + nbs) % nbs)] -= (float)bds;
    if (bdc > 0)
      shot[vc + ((stc + 1) % nbc)] += (float)bdc;
    else
      shot[vc + ((stc - 1 + nbc) % nbc)] -= (float)bdc;
    auto both = [&](int vol, double v) {  // the same spatial weight goes to both channels
      shot[vol * ss + sts] += (float)v;
      shot[stride + vol * sc + stc] += (float)v;
    };
    double w = 0;
    if (dist > r12) {
      double rd = (dist - r34) / r12;
      if (dist > r34)
        w += 1 - rd;
      else {
        w += 1 + rd;
        both(di - 2, -rd);
      }
    } else {
      double rd = (dist - r14) / r12;
      if (dist < r14)
        w += 1 + rd;
      else {
        w += 1 - rd;
        both(di + 2, rd);
      }
    }
    double ic = z / dist;
    if (ic < -1.0) ic = -1.0;
    if (ic > 1.0) ic = 1.0;
    double inc = std::acos(ic);
    if (inc > RAD90 || (std::fabs(inc - RAD90) < 1e-30 && z <= 0)) {
      double e = (inc - RAD135) / RAD90;
      if (inc > RAD135)
        w += 1 - e;
      else {
        w += 1 + e;
        both(di + 1, -e);
      }
    } else {
      double e = (inc - RAD45) / RAD90;
      if (inc < RAD45)
        w += 1 + e;
      else {
        w += 1 - e;
        both(di - 1, e);
      }
    }
    if (y != 0.0 || x != 0.0) {
      double az = std::atan2(y, x);
      int sel = di >> 2;
      double ad = (az - (-RAD_PI_7_8 + RAD45 * sel)) / RAD45;
      ad = std::max(-0.5, std::min(ad, 0.5));
      if (ad > 0) {
        w += 1 - ad;
        both((di + 4) % maxSectors, ad);
      } else {
        w += 1 + ad;
        both((di - 4 + maxSectors) % maxSectors, -ad);
      }
    }
    shot[vs + sts] += (float)(ws + w);
    shot[vc + stc] += (float)(wc + w);
  }
  double acc = 0;
  for (int j = 0; j < 1344; ++j) acc += (double)(shot[j] * shot[j]);
  acc = std::sqrt(acc);
  for (int j = 0; j < 1344; ++j) shot[j] /= (float)acc;
}

}  // namespace

// rgb / qrgb: packed 0x00RRGGBB (pcl::PointXYZRGB::rgba); lab_out (optional): n x 3 normalised Lab of the surface
extern "C" int orc_shot1344(const float* surf, const uint32_t* rgb, const float* normals4, int n, const float* q,
                            const uint32_t* qrgb, int nq, double radius, const float* lrf_in, float* out1344,
                            float* rf9, float* lab_out) {
  if (!(radius > 0)) return -1;
  if (lab_out)
    for (int i = 0; i < n; ++i) rgb2lab(rgb[i], lab_out + 3 * (size_t)i);
  if (lrf_in)
    std::memcpy(rf9, lrf_in, (size_t)nq * 9 * sizeof(float));
  else if (orc_shot_lrf(surf, n, q, nq, radius, rf9, nullptr) != 0)
    return -1;
  Searcher s;
  s.init(surf, n, radius, 0);
  const float kNaNv = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel
  {
    std::vector<Nbr> nb;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; ++i) {
      float* o = out1344 + 1344 * (size_t)i;
      float* rf = rf9 + 9 * (size_t)i;
      const float* c = q + 3 * (size_t)i;
      nb.clear();
      if (finite3(c)) s.query(c, nb);
      const bool ok = std::isfinite(rf[0]) && std::isfinite(rf[3]) && std::isfinite(rf[6]);
      if (!ok || nb.empty() || nb.size() < 5) {
        for (int d = 0; d < 1344; ++d) o[d] = kNaNv;
        if (!ok || nb.empty())
          for (int d = 0; d < 9; ++d) rf[d] = kNaNv;
        continue;
      }
      shotColorRow(surf, rgb, normals4, c, qrgb[i], nb, radius, rf, o);
    }
  }
  return 0;
}

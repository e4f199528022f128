// icet_b200/csrc/synth.h -- deterministic synthetic 64-channel LiDAR scans (bench / test utility).
//
// Not part of the reference; it realises the synthetic workload SURVEY.md 8(d) / BASELINE.md 2
// specify: a spinning LiDAR (ring centres linearly spaced inside +-22.5 deg elevation, `azim` azimuth steps)
// driving along a street canyon (ground plane z = -1.8 m, side walls |y| = 8..15 m piecewise constant
// in x, boxes and vertical cylinders hashed per 20 m tile), max range 120 m, 10 % dropped returns
// stored as (0,0,0), range noise N(0, 0.01 m).  Every quantity is a pure function of
// (seed, scan index, ring, azimuth step) through a counter-based hash, so any scan can be generated
// independently on the host or on the device.
//
// Compiles as plain C++ (host) and as CUDA (device): SYNTH_HD.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__
#else
#define SYNTH_HD
#endif

namespace synth {

SYNTH_HD inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
SYNTH_HD inline uint64_t key(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  return mix64(mix64(mix64(seed ^ 0xD1B54A32D192ED03ull) + a * 0x9E3779B97F4A7C15ull) + b * 0xC2B2AE3D27D4EB4Full) + c;
}
SYNTH_HD inline double u01(uint64_t k) { return (double)(mix64(k) >> 11) * (1.0 / 9007199254740992.0); }
SYNTH_HD inline double gauss(uint64_t k) {  // Box-Muller
  double u1 = u01(k), u2 = u01(k ^ 0x5851F42D4C957F2Dull);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

struct Pose {  // sensor pose in the world: p_world = Rw * p_sensor + tw
  double R[9];
  double t[3];
};

// One trajectory step (scan k -> k+1): dx ~ U(0.2,0.8), dy,dz ~ N(0,0.02), roll/pitch ~ N(0,0.002),
// yaw ~ N(0,0.01).  The pose of scan k is the composition of steps 0..k-1.
SYNTH_HD inline void step_motion(uint64_t seed, int k, double d[6]) {
  d[0] = 0.2 + 0.6 * u01(key(seed, 0xA0, (uint64_t)k, 0));
  d[1] = 0.02 * gauss(key(seed, 0xA1, (uint64_t)k, 0));
  d[2] = 0.02 * gauss(key(seed, 0xA2, (uint64_t)k, 0));
  d[3] = 0.002 * gauss(key(seed, 0xA3, (uint64_t)k, 0));
  d[4] = 0.002 * gauss(key(seed, 0xA4, (uint64_t)k, 0));
  d[5] = 0.01 * gauss(key(seed, 0xA5, (uint64_t)k, 0));
}
SYNTH_HD inline void rpy(double r, double p, double y, double R[9]) {
  double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}
// advance pose by one step (motion expressed in the sensor frame)
SYNTH_HD inline void advance(Pose& P, const double d[6]) {
  double Rs[9], Rn[9];
  rpy(d[3], d[4], d[5], Rs);
  for (int i = 0; i < 3; i++) P.t[i] += P.R[3 * i] * d[0] + P.R[3 * i + 1] * d[1] + P.R[3 * i + 2] * d[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Rn[3 * i + j] = P.R[3 * i] * Rs[j] + P.R[3 * i + 1] * Rs[3 + j] + P.R[3 * i + 2] * Rs[6 + j];
  for (int i = 0; i < 9; i++) P.R[i] = Rn[i];
}
// The driver: the random step of scan k plus a gentle correction towards the lane (y = 0, level, heading along +x),
// so that the trajectory stays inside the street for sequences of any length (a pure random walk of the heading and
// of the pitch leaves the canyon -- and the ground -- after a few hundred scans and the returns disappear).  The
// corrections are a fraction of the random terms per step; the pose of scan k stays a pure function of (seed, k).
SYNTH_HD inline void drive_step(uint64_t seed, int k, const Pose& P, double d[6]) {
  step_motion(seed, k, d);
  const double yaw = atan2(P.R[3], P.R[0]);
  const double pitch = -asin(P.R[6] > 1.0 ? 1.0 : (P.R[6] < -1.0 ? -1.0 : P.R[6]));
  const double roll = atan2(P.R[7], P.R[8]);
  d[5] += 0.15 * (-0.03 * P.t[1] - yaw);   // steer back to the lane centre
  d[4] += 0.15 * (-pitch);
  d[3] += 0.15 * (-roll);
  d[2] += 0.05 * (-P.t[2]);                // suspension: the sensor height stays near its nominal value
}
SYNTH_HD inline void pose_identity(Pose& P) {
  for (int i = 0; i < 9; i++) P.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  P.t[0] = P.t[1] = P.t[2] = 0.0;
}

constexpr double kGroundZ = -1.8;
constexpr double kMaxRange = 120.0;
constexpr double kTile = 20.0;     // objects are hashed per 20 m tile along x
constexpr int kBoxesPerTile = 8;
constexpr int kCylsPerTile = 4;
constexpr double kWallSeg = 10.0;  // wall offset is constant per 10 m segment

SYNTH_HD inline double wall_offset(uint64_t seed, long seg, int side) {
  return 8.0 + 7.0 * u01(key(seed, 0xB0 + (uint64_t)side, (uint64_t)(seg + (1l << 40)), 0));
}

// nearest hit of the world-frame ray o + s*dir (|dir| = 1), s in (0.3, kMaxRange]; returns <0 for no hit
SYNTH_HD inline double cast(uint64_t seed, const double o[3], const double dir[3]) {
  double best = kMaxRange + 1.0;
  // ground
  if (dir[2] < -1e-9) {
    double s = (kGroundZ - o[2]) / dir[2];
    if (s > 0.3 && s < best) best = s;
  }
  // walls: y = +-W(seg), valid where the hit's x falls into that 10 m segment
  long seg0 = (long)floor(o[0] / kWallSeg);
  for (int side = 0; side < 2; side++) {
    double sgn = side ? -1.0 : 1.0;
    if (dir[1] * sgn <= 1e-9) continue;
    for (long seg = seg0 - 13; seg <= seg0 + 13; seg++) {
      double W = sgn * wall_offset(seed, seg, side);
      double s = (W - o[1]) / dir[1];
      if (s > 0.3 && s < best) {
        double hx = o[0] + s * dir[0], hz = o[2] + s * dir[2];
        if (hx >= seg * kWallSeg && hx < (seg + 1) * kWallSeg && hz >= kGroundZ && hz <= kGroundZ + 12.0)
          best = s;
      }
    }
  }
  // boxes and cylinders of the five tiles around the sensor
  long t0 = (long)floor(o[0] / kTile);
  for (long t = t0 - 2; t <= t0 + 2; t++) {
    uint64_t tk = (uint64_t)(t + (1l << 40));
    for (int b = 0; b < kBoxesPerTile; b++) {
      double cx = (t + u01(key(seed, 0xC0, tk, (uint64_t)b))) * kTile;
      double cy = -7.0 + 14.0 * u01(key(seed, 0xC1, tk, (uint64_t)b));
      double hx = 0.4 + 1.6 * u01(key(seed, 0xC2, tk, (uint64_t)b));
      double hy = 0.4 + 1.6 * u01(key(seed, 0xC3, tk, (uint64_t)b));
      double hh = 0.8 + 3.2 * u01(key(seed, 0xC4, tk, (uint64_t)b));
      if (fabs(cy) < 2.0) cy += (cy < 0 ? -2.5 : 2.5);  // keep the driving lane clear
      double lo[3] = {cx - hx, cy - hy, kGroundZ}, hi[3] = {cx + hx, cy + hy, kGroundZ + hh};
      double s0 = 0.0, s1 = best;
      bool ok = true;
      for (int a = 0; a < 3 && ok; a++) {
        if (fabs(dir[a]) < 1e-12) {
          if (o[a] < lo[a] || o[a] > hi[a]) ok = false;
        } else {
          double inv = 1.0 / dir[a];
          double ta = (lo[a] - o[a]) * inv, tb = (hi[a] - o[a]) * inv;
          if (ta > tb) { double tt = ta; ta = tb; tb = tt; }
          if (ta > s0) s0 = ta;
          if (tb < s1) s1 = tb;
          if (s0 > s1) ok = false;
        }
      }
      if (ok && s0 > 0.3 && s0 < best) best = s0;
    }
    for (int c = 0; c < kCylsPerTile; c++) {
      double cx = (t + u01(key(seed, 0xD0, tk, (uint64_t)c))) * kTile;
      double cy = -7.5 + 15.0 * u01(key(seed, 0xD1, tk, (uint64_t)c));
      double rad = 0.15 + 0.45 * u01(key(seed, 0xD2, tk, (uint64_t)c));
      double hh = 2.0 + 6.0 * u01(key(seed, 0xD3, tk, (uint64_t)c));
      if (fabs(cy) < 2.0) cy += (cy < 0 ? -2.5 : 2.5);
      double ox = o[0] - cx, oy = o[1] - cy;
      double a = dir[0] * dir[0] + dir[1] * dir[1];
      if (a < 1e-12) continue;
      double bq = ox * dir[0] + oy * dir[1];
      double cq = ox * ox + oy * oy - rad * rad;
      double disc = bq * bq - a * cq;
      if (disc <= 0.0) continue;
      double s = (-bq - sqrt(disc)) / a;
      if (s > 0.3 && s < best) {
        double hz = o[2] + s * dir[2];
        if (hz >= kGroundZ && hz <= kGroundZ + hh) best = s;
      }
    }
  }
  return best <= kMaxRange ? best : -1.0;
}

// One return of scan `k`: ring `ring` of `rings`, azimuth step `az` of `azim`, sensor pose P.
SYNTH_HD inline void ray(uint64_t seed, int k, const Pose& P, int ring, int rings, int az, int azim,
                         float& x, float& y, float& z) {
  const double deg = 0.017453292519943295;
  // ring centres of `rings` equal slices of [-22.5, +22.5] deg: no ring sits exactly on an elevation-bin
  // edge of a (multiple of 24)-bin grid, which would turn a whole ring into "1-ulp edge points"
  double elev = (-22.5 + 45.0 * ((double)ring + 0.5) / (double)rings) * deg;
  double azr = 6.283185307179586 * (double)az / (double)azim;
  double ds[3] = {cos(elev) * cos(azr), cos(elev) * sin(azr), sin(elev)};
  double dw[3];
  for (int i = 0; i < 3; i++) dw[i] = P.R[3 * i] * ds[0] + P.R[3 * i + 1] * ds[1] + P.R[3 * i + 2] * ds[2];
  double s = cast(seed, P.t, dw);
  uint64_t kk = key(seed, (uint64_t)k + 0x100000ull, (uint64_t)ring, (uint64_t)az);
  bool drop = u01(kk ^ 0xE0) < 0.10;
  if (s < 0.0 || drop) {
    x = y = z = 0.f;
    return;
  }
  s += 0.01 * gauss(kk ^ 0xE1);
  x = (float)(s * ds[0]);
  y = (float)(s * ds[1]);
  z = (float)(s * ds[2]);
}

}  // namespace synth

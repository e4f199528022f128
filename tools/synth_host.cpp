// tools/synth_host.cpp -- host build of the synthetic scan generator (icet_b200/csrc/synth.h).
// Bench / test tooling only (lets the CPU-only tests and the reference arm of bench.py create
// the same kind of scans without touching the CUDA library).
#include <cstdint>
#include <thread>
#include <vector>

#include "../icet_b200/csrc/synth.h"

extern "C" int synth_host_scans(uint64_t seed, int first_scan, int nscans, int rings, int azim, int nthreads,
                                float* out) {
  if (nscans < 1 || rings < 1 || azim < 1 || first_scan < 0 || !out) return -1;
  std::vector<synth::Pose> poses(nscans);
  synth::Pose P;
  synth::pose_identity(P);
  for (int k = 0; k < first_scan + nscans; k++) {
    if (k >= first_scan) poses[k - first_scan] = P;
    double d[6];
    synth::drive_step(seed, k, P, d);
    synth::advance(P, d);
  }
  const int npts = rings * azim;
  const long total = (long)nscans * npts;
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int tid) {
    for (long g = tid; g < total; g += nthreads) {
      const int s = (int)(g / npts), i = (int)(g % npts);
      float x, y, z;
      synth::ray(seed, first_scan + s, poses[s], i / azim, rings, i % azim, azim, x, y, z);
      float* o = out + (size_t)s * 3 * npts;
      o[i] = x; o[npts + i] = y; o[2 * (size_t)npts + i] = z;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  return 0;
}

// relative motion between consecutive scans k -> k+1 as the generator defines it (dx dy dz roll pitch yaw)
extern "C" void synth_host_motion(uint64_t seed, int k, double d[6]) {
  synth::Pose P;
  synth::pose_identity(P);
  double e[6];
  for (int j = 0; j < k; j++) {
    synth::drive_step(seed, j, P, e);
    synth::advance(P, e);
  }
  synth::drive_step(seed, k, P, d);
}

// the motions of `n` consecutive steps k0, k0+1, ... in one walk of the trajectory
extern "C" void synth_host_motions(uint64_t seed, int k0, int n, double* out /* [n][6] */) {
  synth::Pose P;
  synth::pose_identity(P);
  double e[6];
  for (int j = 0; j < k0 + n; j++) {
    synth::drive_step(seed, j, P, e);
    if (j >= k0)
      for (int c = 0; c < 6; c++) out[(size_t)(j - k0) * 6 + c] = e[c];
    synth::advance(P, e);
  }
}

// BASELINE.json configs[4]: an accumulated map in the frame of scan `first_scan + nscans` (the scan that is matched
// against it): `per_scan` returns of each of the `nscans` earlier scans (the simpleMapMaker.cpp:150-160 recipe: a
// fixed-size sample of every scan), taken back through the generator's exact poses.  The sample of scan s is the
// arithmetic progression  (977 s + 65 j) mod (rings * azim), j < per_scan  (distinct rays, a pure function of s), so
// the map is bit-reproducible wherever this library is built.  Dropped returns are skipped; returns the point count;
// out: 3 planes of `cap` floats.
extern "C" long synth_host_map(uint64_t seed, int first_scan, int nscans, int per_scan, int rings, int azim, int nthreads,
                               float* out, long cap) {
  if (nscans < 1 || per_scan < 1 || rings < 1 || azim < 1 || first_scan < 0 || !out) return -1;
  std::vector<synth::Pose> poses(nscans + 1);
  synth::Pose P;
  synth::pose_identity(P);
  for (int k = 0; k <= first_scan + nscans; k++) {
    if (k >= first_scan) poses[k - first_scan] = P;
    double d[6];
    synth::drive_step(seed, k, P, d);
    synth::advance(P, d);
  }
  const synth::Pose& C = poses[nscans];  // the current sensor frame
  const int npts = rings * azim;
  std::vector<float> tmp((size_t)nscans * per_scan * 3);
  std::vector<char> ok((size_t)nscans * per_scan, 0);
  if (nthreads < 1) nthreads = 1;
  auto work = [&](int tid) {
    for (int s = tid; s < nscans; s += nthreads) {
      const synth::Pose& S = poses[s];
      for (int j = 0; j < per_scan; j++) {
        const int i = (int)(((long)977 * (first_scan + s) + (long)65 * j) % npts);
        float x, y, z;
        synth::ray(seed, first_scan + s, S, i / azim, rings, i % azim, azim, x, y, z);
        if (x == 0.f && y == 0.f && z == 0.f) continue;
        double w[3], c[3];
        for (int a = 0; a < 3; a++) w[a] = S.R[3 * a] * x + S.R[3 * a + 1] * y + S.R[3 * a + 2] * z + S.t[a] - C.t[a];
        for (int a = 0; a < 3; a++) c[a] = C.R[a] * w[0] + C.R[3 + a] * w[1] + C.R[6 + a] * w[2];  // R^T
        float* o = tmp.data() + ((size_t)s * per_scan + j) * 3;
        o[0] = (float)c[0]; o[1] = (float)c[1]; o[2] = (float)c[2];
        ok[(size_t)s * per_scan + j] = 1;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  long n = 0;
  for (size_t k = 0; k < ok.size(); k++) {
    if (!ok[k]) continue;
    if (n >= cap) return -2;
    out[n] = tmp[3 * k]; out[cap + n] = tmp[3 * k + 1]; out[2 * cap + n] = tmp[3 * k + 2];
    n++;
  }
  return n;
}

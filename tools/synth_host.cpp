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

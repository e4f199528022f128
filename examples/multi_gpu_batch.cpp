// examples/multi_gpu_batch.cpp -- a C++ caller sharding a batch of scan pairs over every GPU of the box through the C ABI
// (include/icet_b200.h "multi-GPU"): ONE process, one context per device, contiguous pair ranges, one final
// ncclAllGather of 48 floats per pair.  What "run multiple ICETs at once" (reference include/icet.h:43) becomes when
// the ICETs are a batch.
//
//   multi_gpu_batch [npairs] [ndev]     synthetic 64-channel sequence (icet_b200_synth_scans_device), X0 = 0
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "icet_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ < 0) { std::fprintf(stderr, "%s: %s\n", #call, icet_b200_last_error()); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  const int npairs = argc > 1 ? std::atoi(argv[1]) : 64;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  if (argc > 2) ndev = std::atoi(argv[2]);
  if (ndev < 1) { std::fprintf(stderr, "no CUDA device (icet_b200 has no CPU fallback)\n"); return 1; }
  const int n = 64 * 2048;
  icet_b200_multi* m = nullptr;
  CHECK(icet_b200_multi_create(nullptr, ndev, &m));
  icet_b200_params p = {7, 24, 75, 25, 0.1f, 0.1f, 0, 0};
  // every device generates the scans of its own pair range (lo .. hi, one boundary scan duplicated)
  std::vector<const float*> shards(ndev);
  for (int d = 0; d < ndev; d++) {
    const int lo = (int)((long long)npairs * d / ndev), hi = (int)((long long)npairs * (d + 1) / ndev);
    cudaSetDevice(d);
    float* buf = nullptr;
    cudaMalloc(&buf, (size_t)(hi - lo + 1) * 3 * n * sizeof(float));
    CHECK(icet_b200_synth_scans_device(icet_b200_multi_context(m, d), 20240, lo, hi - lo + 1, 64, 2048, buf));
    CHECK(icet_b200_synchronize(icet_b200_multi_context(m, d)));
    shards[d] = buf;
  }
  CHECK(icet_b200_register_sequence_multi_device(m, &p, npairs + 1, shards.data(), n));
  // every device now holds every result: read them back from the last one
  const float* rows = nullptr;
  int per = 0;
  CHECK(icet_b200_multi_gathered(m, ndev - 1, &rows, &per));
  std::vector<float> h((size_t)ndev * per * 48);
  cudaSetDevice(ndev - 1);
  cudaMemcpy(h.data(), rows, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
  for (int d = 0; d < ndev; d++) {
    const int lo = (int)((long long)npairs * d / ndev), hi = (int)((long long)npairs * (d + 1) / ndev);
    for (int k = lo; k < hi && k < lo + 2; k++) {
      const float* r = h.data() + ((size_t)d * per + (k - lo)) * 48;
      std::printf("pair %4d (device %d): X = %.5f %.5f %.5f %.6f %.6f %.6f  pred_stds[0] = %.2e\n", k, d, r[0], r[1], r[2],
                  r[3], r[4], r[5], r[6]);
    }
  }
  for (int d = 0; d < ndev; d++) { cudaSetDevice(d); cudaFree((void*)shards[d]); }
  icet_b200_multi_destroy(m);
  return 0;
}

// icet_b200/csrc/multi_abi.inl -- multi-GPU entry points of the C ABI: ONE process, one context per
// device, the pairs of a batch sharded by contiguous range, no exchange inside the registration, and ONE ncclAllGather
// of 48 floats per pair (X 6 | pred_stds 6 | Q 36) at the end, after which every device holds every result.
// Included by icet_b200.cu inside extern "C".  NCCL is loaded at run time (dlopen of libnccl.so.2): the single-GPU
// library has no link-time dependency on it.
}  // extern "C" (closed around the C++ helpers of this file; reopened below)

struct icet_b200_multi {
  int ndev = 0;
  std::vector<int> dev;
  std::vector<icet_b200_ctx*> ctx;
  std::vector<ncclComm_t> comm;
  std::vector<DevBuf> sendbuf, recvbuf, resbuf;
  int shard_rows = 0;  // rows per shard of the last gather (the largest shard; smaller ones are zero padded)
  int npairs = 0;
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

#define NCK(call)                                                                                      \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess)                                                                             \
      return fail(ICET_B200_E_CUDA, std::string(#call) + ": " + (m->GetErrorString ? m->GetErrorString(r_) : "NCCL error")); \
  } while (0)

int icet_b200_multi_gathered_host_impl(icet_b200_multi* m, int32_t d, float* out) {
  if (!m || d < 0 || d >= m->ndev || !out) return fail(ICET_B200_E_INVALID, "bad argument");
  if (m->shard_rows == 0) return 0;
  CK(cudaSetDevice(m->dev[d]));
  CK(cudaMemcpy(out, m->recvbuf[d].p, (size_t)m->ndev * m->shard_rows * 48 * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

template <class F>
static int multi_fan_out(icet_b200_multi* m, F&& per_device) {
  std::vector<int> rc(m->ndev, 0);
  std::vector<std::string> err(m->ndev);
  std::vector<std::thread> th;
  for (int d = 0; d < m->ndev; d++)
    th.emplace_back([&, d]() {
      cudaSetDevice(m->dev[d]);
      rc[d] = per_device(d);
      if (rc[d]) err[d] = g_err;  // (thread-local message of the worker)
    });
  for (auto& t : th) t.join();
  for (int d = 0; d < m->ndev; d++)
    if (rc[d]) return fail(rc[d], "device " + std::to_string(m->dev[d]) + ": " + err[d]);
  return 0;
}

extern "C" {

int icet_b200_multi_create(const int32_t* devices, int32_t ndev, icet_b200_multi** out) {
  if (!out) return fail(ICET_B200_E_INVALID, "out-pointer is NULL");
  *out = nullptr;
  if (ndev < 1 || ndev > 64) return fail(ICET_B200_E_INVALID, "ndev out of range");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have <= 0) {
    cudaGetLastError();
    return fail(ICET_B200_E_NODEVICE, "no CUDA device available; icet_b200 has no CPU fallback");
  }
  icet_b200_multi* m = new icet_b200_multi();
  m->ndev = ndev;
  for (int d = 0; d < ndev; d++) {
    const int id = devices ? devices[d] : d;
    if (id < 0 || id >= have) { delete m; return fail(ICET_B200_E_INVALID, "device index out of range"); }
    for (int e = 0; e < d; e++)
      if (m->dev[e] == id) { delete m; return fail(ICET_B200_E_INVALID, "duplicate device in the list"); }
    m->dev.push_back(id);
  }
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    m->h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (m->h) break;
  }
  if (!m->h) { delete m; return fail(ICET_B200_E_NODEVICE, "libnccl.so.2 not found (needed for the multi-GPU entry points)"); }
  m->CommInitAll = (decltype(m->CommInitAll))dlsym(m->h, "ncclCommInitAll");
  m->CommDestroy = (decltype(m->CommDestroy))dlsym(m->h, "ncclCommDestroy");
  m->AllGather = (decltype(m->AllGather))dlsym(m->h, "ncclAllGather");
  m->GroupStart = (decltype(m->GroupStart))dlsym(m->h, "ncclGroupStart");
  m->GroupEnd = (decltype(m->GroupEnd))dlsym(m->h, "ncclGroupEnd");
  m->GetErrorString = (decltype(m->GetErrorString))dlsym(m->h, "ncclGetErrorString");
  if (!m->CommInitAll || !m->CommDestroy || !m->AllGather || !m->GroupStart || !m->GroupEnd) {
    delete m;
    return fail(ICET_B200_E_NODEVICE, "libnccl.so.2 lacks the expected symbols");
  }
  m->ctx.assign(ndev, nullptr);
  for (int d = 0; d < ndev; d++) {
    int rc = icet_b200_create(m->dev[d], &m->ctx[d]);
    if (rc) {
      for (int e = 0; e < d; e++) icet_b200_destroy(m->ctx[e]);
      delete m;
      return rc;
    }
  }
  m->comm.assign(ndev, nullptr);
  m->sendbuf.resize(ndev); m->recvbuf.resize(ndev); m->resbuf.resize(ndev);
  ncclResult_t r = m->CommInitAll(m->comm.data(), ndev, m->dev.data());
  if (r != ncclSuccess) {
    const std::string msg = std::string("ncclCommInitAll: ") + (m->GetErrorString ? m->GetErrorString(r) : "NCCL error");
    for (int d = 0; d < ndev; d++) icet_b200_destroy(m->ctx[d]);
    delete m;
    return fail(ICET_B200_E_CUDA, msg);
  }
  *out = m;
  return 0;
}

int icet_b200_multi_destroy(icet_b200_multi* m) {
  if (!m) return 0;
  for (int d = 0; d < m->ndev; d++) {
    cudaSetDevice(m->dev[d]);
    if (m->comm[d]) m->CommDestroy(m->comm[d]);
    m->sendbuf[d].release(); m->recvbuf[d].release(); m->resbuf[d].release();
    icet_b200_destroy(m->ctx[d]);
  }
  delete m;  // (the NCCL library stays loaded)
  return 0;
}

int icet_b200_multi_devices(icet_b200_multi* m) { return m ? m->ndev : 0; }
icet_b200_ctx* icet_b200_multi_context(icet_b200_multi* m, int32_t d) { return (m && d >= 0 && d < m->ndev) ? m->ctx[d] : nullptr; }

// contiguous pair range of device d: pairs [P d / G, P (d + 1) / G)
static inline void multi_range(int npairs, int d, int ndev, int* lo, int* hi) {
  *lo = (int)(((long long)npairs * d) / ndev);
  *hi = (int)(((long long)npairs * (d + 1)) / ndev);
}

// the closing collective: rows of 48 floats out of every device's result records, one all-gather
static int multi_gather(icet_b200_multi* m, int npairs, const std::vector<const icet_b200_result*>& d_res) {
  int rows = 0;
  for (int d = 0; d < m->ndev; d++) {
    int lo, hi;
    multi_range(npairs, d, m->ndev, &lo, &hi);
    rows = std::max(rows, hi - lo);
  }
  m->shard_rows = rows;
  m->npairs = npairs;
  if (rows == 0) return 0;
  const size_t row_b = 48 * sizeof(float);
  for (int d = 0; d < m->ndev; d++) {
    CK(cudaSetDevice(m->dev[d]));
    int rc = m->sendbuf[d].ensure((size_t)rows * row_b);
    if (rc) return rc;
    rc = m->recvbuf[d].ensure((size_t)m->ndev * rows * row_b);
    if (rc) return rc;
    int lo, hi;
    multi_range(npairs, d, m->ndev, &lo, &hi);
    cudaStream_t st = m->ctx[d]->stream;
    CK(cudaMemsetAsync(m->sendbuf[d].p, 0, (size_t)rows * row_b, st));
    if (hi > lo)
      CK(cudaMemcpy2DAsync(m->sendbuf[d].p, row_b, d_res[d], sizeof(icet_b200_result), row_b, (size_t)(hi - lo),
                           cudaMemcpyDeviceToDevice, st));
  }
  NCK(m->GroupStart());
  for (int d = 0; d < m->ndev; d++)
    NCK(m->AllGather(m->sendbuf[d].p, m->recvbuf[d].p, (size_t)rows * 48, ncclFloat, m->comm[d], m->ctx[d]->stream));
  NCK(m->GroupEnd());
  for (int d = 0; d < m->ndev; d++) {
    CK(cudaSetDevice(m->dev[d]));
    CK(cudaStreamSynchronize(m->ctx[d]->stream));
  }
  return 0;
}

int icet_b200_multi_gathered(icet_b200_multi* m, int32_t d, const float** rows, int32_t* rows_per_shard) {
  if (!m || d < 0 || d >= m->ndev || !rows || !rows_per_shard) return fail(ICET_B200_E_INVALID, "bad argument");
  *rows = (const float*)m->recvbuf[d].p;
  *rows_per_shard = m->shard_rows;
  return 0;
}

int icet_b200_multi_gathered_host(icet_b200_multi* m, int32_t d, float* out) {
  return icet_b200_multi_gathered_host_impl(m, d, out);
}

int icet_b200_register_batch_multi(icet_b200_multi* m, const icet_b200_params* p, int32_t npairs, const float* const* scan1,
                                   const int32_t* n1, const float* const* scan2, const int32_t* n2, const float* x0,
                                   icet_b200_result* out) {
  if (!m) return fail(ICET_B200_E_INVALID, "multi handle is NULL");
  int rc = validate(p);
  if (rc) return rc;
  if (p->flags & ICET_B200_FLAG_CHAIN_X0)
    return fail(ICET_B200_E_INVALID, "ICET_B200_FLAG_CHAIN_X0 serialises the pairs: use one device");
  if (npairs < 0 || (npairs > 0 && (!scan1 || !scan2 || !n1 || !n2 || !out))) return fail(ICET_B200_E_INVALID, "NULL argument");
  rc = multi_fan_out(m, [&](int d) -> int {
    int lo, hi;
    multi_range(npairs, d, m->ndev, &lo, &hi);
    if (hi == lo) return 0;
    return icet_b200_register_batch(m->ctx[d], p, hi - lo, scan1 + lo, n1 + lo, scan2 + lo, n2 + lo,
                                    x0 ? x0 + (size_t)lo * 6 : nullptr, out + lo);
  });
  if (rc) return rc;
  std::vector<const icet_b200_result*> d_res(m->ndev);
  for (int d = 0; d < m->ndev; d++) d_res[d] = (const icet_b200_result*)m->ctx[d]->resbuf.p;  // device copy of the shard's results
  return multi_gather(m, npairs, d_res);
}

int icet_b200_register_sequence_multi_device(icet_b200_multi* m, const icet_b200_params* p, int32_t nscans,
                                             const float* const* shard_scans, int32_t n) {
  if (!m) return fail(ICET_B200_E_INVALID, "multi handle is NULL");
  int rc = validate(p);
  if (rc) return rc;
  if (p->flags & ICET_B200_FLAG_CHAIN_X0)
    return fail(ICET_B200_E_INVALID, "ICET_B200_FLAG_CHAIN_X0 serialises the pairs: use one device");
  if (nscans < 1 || n < 0 || !shard_scans) return fail(ICET_B200_E_INVALID, "bad argument");
  const int npairs = nscans - 1;
  rc = multi_fan_out(m, [&](int d) -> int {
    int lo, hi;
    multi_range(npairs, d, m->ndev, &lo, &hi);
    if (hi == lo) return 0;
    if (!shard_scans[d]) return fail(ICET_B200_E_INVALID, "shard pointer is NULL");
    int r = m->resbuf[d].ensure((size_t)(hi - lo) * sizeof(icet_b200_result));
    if (r) return r;
    r = icet_b200_register_sequence_device(m->ctx[d], p, hi - lo + 1, shard_scans[d], n, (icet_b200_result*)m->resbuf[d].p);
    if (r) return r;
    return icet_b200_synchronize(m->ctx[d]);
  });
  if (rc) return rc;
  std::vector<const icet_b200_result*> d_res(m->ndev);
  for (int d = 0; d < m->ndev; d++) d_res[d] = (const icet_b200_result*)m->resbuf[d].p;
  return multi_gather(m, npairs, d_res);
}

// icet_b200/csrc/callers_abi.inl -- C ABI of the callers layer (include/icet_b200.h "callers either side of the
// path"), included by icet_b200.cu inside its extern "C" block.  Host code only enqueues; see callers.cuh.

struct icet_b200_node {
  icet_b200_ctx* ctx = nullptr;
  icet_b200_params p{};
  icet_b200_odometry_params op{};
  int cap = 0;  // rows per scan the buffers were sized for
  bool initialized = false;
  bool have_result = false;
  DevBuf state;    // NodeState
  DevBuf slots;    // [m + 1][3][cap] filtered clouds of the current call (slot 0 = prev_pcl_matrix)
  DevBuf prev;     // [3][cap] prev_pcl_matrix between calls
  DevBuf kept;     // int32 [m + 1]
  DevBuf tilecnt;  // int32 [m + 1][ntile]
  DevBuf jobs;     // FilterJob [m + 1]
  DevBuf desc;     // PairDesc [m]
  DevBuf hres, hpose, hscan;  // device side of the one-scan host callback
  DevBuf last;     // icet_b200_result of the most recent registration
  void* pin = nullptr;  // pinned host staging of the job list
  size_t pin_cap = 0;
  cudaEvent_t ev_jobs = nullptr;  // the job list upload of the previous call has finished
  const float* cur_scan = nullptr;
  const int32_t* cur_n = nullptr;
};

struct icet_b200_map {
  icet_b200_ctx* ctx = nullptr;
  int cap = 0, ldr = 0;
  DevBuf ring;   // [3][ldr]
  DevBuf state;  // MapState
  DevBuf idx;    // uploaded row lists
  void* pin = nullptr;
  size_t pin_cap = 0;
  cudaEvent_t ev_idx = nullptr;
};

static int pin_ensure(void** pin, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*pin) cudaFreeHost(*pin);
  *pin = nullptr;
  *cap = 0;
  CK(cudaMallocHost(pin, bytes));
  *cap = bytes;
  return 0;
}

int icet_b200_node_create(icet_b200_ctx* c, const icet_b200_params* p, const icet_b200_odometry_params* op,
                          int32_t max_points, const float* X_homo0, const float* x0, icet_b200_node** out) {
  if (!c || !op || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  int rc = validate(p);
  if (rc) return rc;
  if (max_points < 1) return fail(ICET_B200_E_INVALID, "max_points must be >= 1");
  if (!(op->rate_hz == op->rate_hz)) return fail(ICET_B200_E_INVALID, "rate_hz is NaN");
  CK(cudaSetDevice(c->device));
  icet_b200_node* nd = new icet_b200_node();
  nd->ctx = c;
  nd->p = *p;
  nd->op = *op;
  nd->cap = (max_points + 3) & ~3;
  if (op->chain_x0) nd->p.flags |= ICET_B200_FLAG_CHAIN_X0; else nd->p.flags &= ~ICET_B200_FLAG_CHAIN_X0;
  auto bail = [&](int r) { icet_b200_node_destroy(nd); return r; };
  if ((rc = nd->state.ensure(sizeof(NodeState)))) return bail(rc);
  if ((rc = nd->prev.ensure((size_t)3 * nd->cap * sizeof(float)))) return bail(rc);
  if ((rc = nd->last.ensure(sizeof(icet_b200_result)))) return bail(rc);
  if ((rc = pin_ensure(&nd->pin, &nd->pin_cap, 4096))) return bail(rc);
  if (cudaEventCreateWithFlags(&nd->ev_jobs, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(ICET_B200_E_CUDA, "cudaEventCreate failed"));
  float* h = (float*)nd->pin;
  if (X_homo0) memcpy(h, X_homo0, 16 * sizeof(float));
  if (x0) memcpy(h + 16, x0, 6 * sizeof(float));
  float* d_tmp = (float*)nd->prev.p;  // scratch for the initial values (prev is unused until the first scan)
  CK(cudaMemcpyAsync(d_tmp, h, 22 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  k_node_init<<<1, 32, 0, c->stream>>>((NodeState*)nd->state.p, X_homo0 ? d_tmp : nullptr, x0 ? d_tmp + 16 : nullptr);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  *out = nd;
  return 0;
}

int icet_b200_node_destroy(icet_b200_node* nd) {
  if (!nd) return 0;
  cudaSetDevice(nd->ctx->device);
  cudaStreamSynchronize(nd->ctx->stream);
  nd->state.release(); nd->slots.release(); nd->prev.release(); nd->kept.release(); nd->tilecnt.release();
  nd->jobs.release(); nd->desc.release(); nd->hres.release(); nd->hpose.release(); nd->hscan.release();
  nd->last.release();
  if (nd->pin) cudaFreeHost(nd->pin);
  if (nd->ev_jobs) cudaEventDestroy(nd->ev_jobs);
  delete nd;
  return 0;
}

int icet_b200_node_push_device(icet_b200_node* nd, int32_t nscans, const float* scans, int32_t n, icet_b200_result* res,
                               icet_b200_pose* poses) {
  if (!nd) return fail(ICET_B200_E_INVALID, "node is NULL");
  if (nscans < 0 || n < 0 || n > nd->cap || (nscans > 0 && !scans)) return fail(ICET_B200_E_INVALID, "bad argument");
  if (nscans == 0) return 0;
  icet_b200_ctx* c = nd->ctx;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int cap = nd->cap;
  const int nslots = nscans + (nd->initialized ? 1 : 0);  // slot 0: prev_pcl_matrix
  const int npairs = nslots - 1;
  if (npairs > 0 && (!res || !poses)) return fail(ICET_B200_E_INVALID, "res / poses is NULL");
  int rc;
  if ((rc = nd->slots.ensure((size_t)nslots * 3 * cap * sizeof(float)))) return rc;
  if ((rc = nd->kept.ensure((size_t)nslots * sizeof(int32_t)))) return rc;
  const int ntile = std::max(1, (std::max(n, cap) + FILT_TILE - 1) / FILT_TILE);
  if ((rc = nd->tilecnt.ensure((size_t)nslots * ntile * sizeof(int32_t)))) return rc;
  if ((rc = nd->jobs.ensure((size_t)nslots * sizeof(FilterJob)))) return rc;
  if ((rc = nd->desc.ensure((size_t)std::max(1, npairs) * sizeof(PairDesc)))) return rc;
  CK(cudaEventSynchronize(nd->ev_jobs));
  if ((rc = pin_ensure(&nd->pin, &nd->pin_cap, (size_t)nslots * sizeof(FilterJob)))) return rc;
  FilterJob* hj = (FilterJob*)nd->pin;
  float* slots = (float*)nd->slots.p;
  int32_t* kept = (int32_t*)nd->kept.p;
  NodeState* stt = (NodeState*)nd->state.p;
  for (int s = 0; s < nslots; s++) {
    FilterJob j;
    j.dst = slots + (size_t)s * 3 * cap;
    j.ld_dst = cap;
    j.kept = kept + s;
    if (nd->initialized && s == 0) {  // prev_pcl_matrix of the previous callback, already filtered
      j.src = (const float*)nd->prev.p; j.n = cap; j.ld = cap; j.n_dev = &stt->prev_n; j.min_d = -1.f;
    } else {
      const int k = nd->initialized ? s - 1 : s;
      j.src = scans + (size_t)k * 3 * n; j.n = n; j.ld = n; j.n_dev = nullptr;
      // the first scan ever is stored unfiltered (odometry.cpp:47-52, simpleMapMaker.cpp:86-92)
      j.min_d = (!nd->initialized && s == 0) ? -1.f : nd->op.min_range;
    }
    hj[s] = j;
  }
  CK(cudaMemcpyAsync(nd->jobs.p, hj, (size_t)nslots * sizeof(FilterJob), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(nd->ev_jobs, st));
  k_range_count<<<dim3(ntile, nslots), FILT_THREADS, 0, st>>>((const FilterJob*)nd->jobs.p, ntile, (int32_t*)nd->tilecnt.p);
  k_range_scatter<<<dim3(ntile, nslots), FILT_THREADS, 0, st>>>((const FilterJob*)nd->jobs.p, ntile,
                                                               (const int32_t*)nd->tilecnt.p);
  c->launches += 2;
  if (npairs > 0) {
    k_seq_desc<<<(npairs + 127) / 128, 128, 0, st>>>((PairDesc*)nd->desc.p, npairs, slots, cap, kept);
    c->launches++;
    CK(cudaGetLastError());
    rc = batch_device_impl(c, &nd->p, npairs, nullptr, nd->op.chain_x0 ? stt->X0 : nullptr, res, false,
                           (const PairDesc*)nd->desc.p, cap);
    if (rc) return rc;
    k_node_poses<<<1, 32, 0, st>>>(stt, res, npairs, kept, nd->op, poses);
    c->launches++;
    CK(cudaMemcpyAsync(nd->last.p, res + (npairs - 1), sizeof(icet_b200_result), cudaMemcpyDeviceToDevice, st));
    nd->have_result = true;
  }
  // prev_pcl_matrix = pcl_matrix (odometry.cpp:89, simpleMapMaker.cpp:162)
  k_copy_cloud<<<std::max(1, std::min((cap + 255) / 256, c->sm_count * 4)), 256, 0, st>>>(
      slots + (size_t)(nslots - 1) * 3 * cap, (float*)nd->prev.p, cap, kept + (nslots - 1));
  c->launches++;
  CK(cudaMemcpyAsync(&stt->prev_n, kept + (nslots - 1), sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  CK(cudaGetLastError());
  nd->initialized = true;
  nd->cur_scan = (const float*)nd->prev.p;
  nd->cur_n = &stt->prev_n;
  return npairs;
}

int icet_b200_node_push(icet_b200_node* nd, const float* scan, int32_t n, int32_t ld, icet_b200_result* res,
                        icet_b200_pose* pose) {
  if (!nd) return fail(ICET_B200_E_INVALID, "node is NULL");
  if (n < 0 || ld < n || n > nd->cap || (n > 0 && !scan)) return fail(ICET_B200_E_INVALID, "bad argument");
  icet_b200_ctx* c = nd->ctx;
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = nd->hscan.ensure((size_t)3 * std::max(1, n) * sizeof(float)))) return rc;
  if ((rc = nd->hres.ensure(sizeof(icet_b200_result)))) return rc;
  if ((rc = nd->hpose.ensure(sizeof(icet_b200_pose)))) return rc;
  CK(cudaStreamSynchronize(c->stream));  // hscan of the previous callback is no longer read
  if (n > 0) {
    if (ld == n) {
      CK(cudaMemcpyAsync(nd->hscan.p, scan, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    } else {
      CK(cudaMemcpy2DAsync(nd->hscan.p, (size_t)n * sizeof(float), scan, (size_t)ld * sizeof(float),
                           (size_t)n * sizeof(float), 3, cudaMemcpyHostToDevice, c->stream));
    }
  }
  const int np = icet_b200_node_push_device(nd, 1, (const float*)nd->hscan.p, n, (icet_b200_result*)nd->hres.p,
                                            (icet_b200_pose*)nd->hpose.p);
  if (np < 0) return np;
  if (np > 0) {
    if (res) CK(cudaMemcpyAsync(res, nd->hres.p, sizeof(icet_b200_result), cudaMemcpyDeviceToHost, c->stream));
    if (pose) CK(cudaMemcpyAsync(pose, nd->hpose.p, sizeof(icet_b200_pose), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  rc = check_loop_watchdog(c);
  if (rc) return rc;
  return np;
}

int icet_b200_node_current_scan(icet_b200_node* nd, const float** scan, const int32_t** n_dev, int32_t* ld) {
  if (!nd || !nd->initialized) return fail(ICET_B200_E_INVALID, "node has not seen a scan yet");
  if (scan) *scan = nd->cur_scan;
  if (n_dev) *n_dev = nd->cur_n;
  if (ld) *ld = nd->cap;
  return 0;
}

int icet_b200_node_last_result(icet_b200_node* nd, const icet_b200_result** res_dev) {
  if (!nd || !res_dev) return fail(ICET_B200_E_INVALID, "NULL argument");
  *res_dev = nd->have_result ? (const icet_b200_result*)nd->last.p : nullptr;
  return 0;
}

int icet_b200_map_create(icet_b200_ctx* c, int32_t capacity, icet_b200_map** out) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (capacity < 1) return fail(ICET_B200_E_INVALID, "capacity must be >= 1");
  CK(cudaSetDevice(c->device));
  icet_b200_map* m = new icet_b200_map();
  m->ctx = c;
  m->cap = capacity;
  m->ldr = (capacity + 3) & ~3;
  int rc;
  auto bail = [&](int r) { icet_b200_map_destroy(m); return r; };
  if ((rc = m->ring.ensure((size_t)3 * m->ldr * sizeof(float)))) return bail(rc);
  if ((rc = m->state.ensure(sizeof(MapState)))) return bail(rc);
  if (cudaEventCreateWithFlags(&m->ev_idx, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(ICET_B200_E_CUDA, "cudaEventCreate failed"));
  CK(cudaMemsetAsync(m->ring.p, 0, (size_t)3 * m->ldr * sizeof(float), c->stream));  // MatrixXf matrix(maxSize, 3)
  CK(cudaMemsetAsync(m->state.p, 0, sizeof(MapState), c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *out = m;
  return 0;
}

int icet_b200_map_destroy(icet_b200_map* m) {
  if (!m) return 0;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  m->ring.release(); m->state.release(); m->idx.release();
  if (m->pin) cudaFreeHost(m->pin);
  if (m->ev_idx) cudaEventDestroy(m->ev_idx);
  delete m;
  return 0;
}

int icet_b200_map_add_scan_device(icet_b200_map* m, const float* scan, int32_t n, int32_t ld, const int32_t* n_dev,
                                  const int32_t* idx, int32_t count, const float* X, float guard_trans,
                                  float guard_rot) {
  if (!m || !X) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (n < 0 || ld < n || count < 0 || (count > 0 && n > 0 && !scan)) return fail(ICET_B200_E_INVALID, "bad argument");
  icet_b200_ctx* c = m->ctx;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int32_t* d_idx = nullptr;
  if (idx && count > 0) {
    for (int k = 0; k < count; k++)
      if (idx[k] < 0 || idx[k] >= n) return fail(ICET_B200_E_INVALID, "row index out of range");
    int rc = m->idx.ensure((size_t)count * sizeof(int32_t));
    if (rc) return rc;
    CK(cudaEventSynchronize(m->ev_idx));
    if ((rc = pin_ensure(&m->pin, &m->pin_cap, (size_t)count * sizeof(int32_t)))) return rc;
    memcpy(m->pin, idx, (size_t)count * sizeof(int32_t));
    CK(cudaMemcpyAsync(m->idx.p, m->pin, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(m->ev_idx, st));
    d_idx = (const int32_t*)m->idx.p;
  }
  k_map_enqueue<<<1, 256, 0, st>>>((MapState*)m->state.p, (float*)m->ring.p, m->cap, m->ldr, scan, ld, n_dev, n, d_idx,
                                   count, X, guard_trans, guard_rot);
  const int nv = (m->cap + 3) / 4;
  const int grid = std::max(1, std::min((nv + 255) / 256, c->sm_count * 8));
  k_map_reexpress<<<grid, 256, 0, st>>>((const MapState*)m->state.p, (float*)m->ring.p, m->cap, m->ldr);
  c->launches += 2;
  CK(cudaGetLastError());
  return 0;
}

int icet_b200_map_get_device(icet_b200_map* m, float* out, int32_t ld_out, int32_t* n_out) {
  if (!m || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (ld_out < 1) return fail(ICET_B200_E_INVALID, "bad ld_out");
  icet_b200_ctx* c = m->ctx;
  CK(cudaSetDevice(c->device));
  const int grid = std::max(1, std::min((m->cap + 255) / 256, c->sm_count * 8));
  k_map_get<<<grid, 256, 0, c->stream>>>((const MapState*)m->state.p, (const float*)m->ring.p, m->cap, m->ldr, out, ld_out,
                                         n_out);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int icet_b200_map_get(icet_b200_map* m, float* out, int32_t ld_out, int32_t* n_out) {
  if (!m || !out || !n_out) return fail(ICET_B200_E_INVALID, "NULL argument");
  icet_b200_ctx* c = m->ctx;
  CK(cudaSetDevice(c->device));
  MapState hs;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(&hs, m->state.p, sizeof(MapState), cudaMemcpyDeviceToHost));
  const int rows = hs.filled ? m->cap : hs.pos;
  if (ld_out < rows) return fail(ICET_B200_E_INVALID, "ld_out smaller than the number of stored rows");
  *n_out = rows;
  if (rows == 0) return 0;
  DevBuf tmp;
  int rc = tmp.ensure((size_t)3 * rows * sizeof(float));
  if (rc) return rc;
  rc = icet_b200_map_get_device(m, (float*)tmp.p, rows, nullptr);
  if (rc) { tmp.release(); return rc; }
  cudaError_t e = cudaMemcpy2DAsync(out, (size_t)ld_out * sizeof(float), tmp.p, (size_t)rows * sizeof(float),
                                    (size_t)rows * sizeof(float), 3, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  tmp.release();
  if (e != cudaSuccess) return fail(ICET_B200_E_CUDA, std::string("map_get: ") + cudaGetErrorString(e));
  return 0;
}

int icet_b200_transform_cloud_device(icet_b200_ctx* c, const float* cloud, int32_t n, int32_t ld, const int32_t* n_dev,
                                     const float* X, int32_t mode, float* out, int32_t ld_out) {
  if (!c || !X) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (n < 0 || ld < n || ld_out < n || (mode != 0 && mode != 1) || (n > 0 && (!cloud || !out)))
    return fail(ICET_B200_E_INVALID, "bad argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(c->device));
  const int grid = std::max(1, std::min((n + 255) / 256, c->sm_count * 8));
  k_transform_cloud<<<grid, 256, 0, c->stream>>>(cloud, n, ld, n_dev, X, mode, out, ld_out);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int icet_b200_transform_cloud(icet_b200_ctx* c, const float* cloud, int32_t n, int32_t ld, const float* X, int32_t mode,
                              float* out, int32_t ld_out) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (n < 0 || ld_out < n) return fail(ICET_B200_E_INVALID, "bad argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  int rc = c->rawbuf[1].ensure((size_t)3 * n * sizeof(float));
  if (rc) return rc;
  rc = icet_b200_transform_cloud_device(c, cloud, n, ld, nullptr, X, mode, (float*)c->rawbuf[1].p, n);
  if (rc) return rc;
  CK(cudaMemcpy2DAsync(out, (size_t)ld_out * sizeof(float), c->rawbuf[1].p, (size_t)n * sizeof(float),
                       (size_t)n * sizeof(float), 3, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// -- ingest ------------------------------------------------------------------------------------------------------
static int cloud_bytes(const icet_b200_cloud* c, size_t* bytes) {
  if (!c) return fail(ICET_B200_E_INVALID, "cloud is NULL");
  if (c->n < 0 || (c->n > 0 && !c->data)) return fail(ICET_B200_E_INVALID, "bad cloud");
  if (c->dtype != ICET_B200_F32 && c->dtype != ICET_B200_F64 && c->dtype != ICET_B200_I32)
    return fail(ICET_B200_E_INVALID, "unknown dtype");
  const int es = c->dtype == ICET_B200_F64 ? 8 : 4;
  if (c->plane_stride > 0) {
    if (c->plane_stride < c->n) return fail(ICET_B200_E_INVALID, "plane_stride smaller than n");
    *bytes = ((size_t)2 * c->plane_stride + c->n) * es;
    return 0;
  }
  if (c->point_step < es || c->point_step % es) return fail(ICET_B200_E_INVALID, "point_step must be a multiple of the element size");
  for (int k = 0; k < 3; k++)
    if (c->off[k] < 0 || c->off[k] % es || c->off[k] + es > c->point_step)
      return fail(ICET_B200_E_INVALID, "field offset outside the record or misaligned");
  *bytes = (size_t)c->n * c->point_step;
  return 0;
}

// stages the raw bytes in `raw` (device) and converts into `out`; all on the context's stream
static int ingest_into(icet_b200_ctx* c, const icet_b200_cloud* cl, DevBuf& raw, float* out, int32_t ld) {
  size_t bytes = 0;
  int rc = cloud_bytes(cl, &bytes);
  if (rc) return rc;
  if (ld < cl->n) return fail(ICET_B200_E_INVALID, "ld smaller than n");
  if (cl->n == 0) return 0;
  if (!out) return fail(ICET_B200_E_INVALID, "out is NULL");
  if ((rc = raw.ensure(bytes))) return rc;
  CK(cudaMemcpyAsync(raw.p, cl->data, bytes, cudaMemcpyHostToDevice, c->stream));
  IngestDesc d;
  d.raw = (const unsigned char*)raw.p;
  d.n = cl->n; d.step = cl->point_step; d.dtype = cl->dtype; d.plane_stride = cl->plane_stride; d.divide = cl->divide;
  for (int k = 0; k < 3; k++) d.off[k] = cl->off[k];
  k_ingest<<<(cl->n + 255) / 256, 256, 0, c->stream>>>(d, out, ld);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

int icet_b200_ingest(icet_b200_ctx* c, const icet_b200_cloud* cl, float* out, int32_t ld) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));  // the staging buffer of the previous call is no longer read
  int rc = ingest_into(c, cl, c->rawbuf[0], out, ld);
  if (rc) return rc;
  // pageable source: the runtime has copied it out before returning; pinned source: wait for the DMA
  cudaPointerAttributes at;
  if (cl->n > 0 && cudaPointerGetAttributes(&at, cl->data) == cudaSuccess && at.type == cudaMemoryTypeHost)
    CK(cudaStreamSynchronize(c->stream));
  cudaGetLastError();
  return 0;
}

int icet_b200_register_clouds(icet_b200_ctx* c, const icet_b200_params* p, const icet_b200_cloud* s1,
                              const icet_b200_cloud* s2, const float x0[6], icet_b200_result* out) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  int rc = validate(p);
  if (rc) return rc;
  size_t b1, b2;
  if ((rc = cloud_bytes(s1, &b1)) || (rc = cloud_bytes(s2, &b2))) return rc;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const int n1 = s1->n, n2 = s2->n;
  const size_t l1 = ((size_t)n1 + 3) & ~(size_t)3, l2 = ((size_t)n2 + 3) & ~(size_t)3;
  if ((rc = c->planebuf.ensure((3 * l1 + 3 * l2 + 8) * sizeof(float) + sizeof(icet_b200_result)))) return rc;
  float* d1 = (float*)c->planebuf.p;
  float* d2 = d1 + 3 * l1;
  float* dx = d2 + 3 * l2;
  icet_b200_result* dres = (icet_b200_result*)(dx + 8);
  if ((rc = ingest_into(c, s1, c->rawbuf[0], d1, (int32_t)l1))) return rc;
  if ((rc = ingest_into(c, s2, c->rawbuf[1], d2, (int32_t)l2))) return rc;
  if (x0) CK(cudaMemcpyAsync(dx, x0, 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  PairDesc d{d1, d2, n1, (int)l1, n2, (int)l2};
  const bool dump = c->dump_on != 0;
  if (dump && (rc = ensure_dump(c, p))) return rc;
  rc = batch_device_impl(c, p, 1, &d, x0 ? dx : nullptr, dres, dump);
  if (rc) return rc;
  if (dump) { c->dump_params = *p; c->dump_valid = true; }
  c->last_n2 = n2;
  c->last_runlen = p->runlen;
  c->last_valid = true;  // (the planes of both clouds stay in planebuf until the next call)
  CK(cudaMemcpyAsync(out, dres, sizeof(icet_b200_result), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_loop_watchdog(c);
}

int icet_b200_node_push_cloud(icet_b200_node* nd, const icet_b200_cloud* cl, icet_b200_result* res, icet_b200_pose* pose) {
  if (!nd || !cl) return fail(ICET_B200_E_INVALID, "NULL argument");
  size_t bytes;
  int rc = cloud_bytes(cl, &bytes);
  if (rc) return rc;
  if (cl->n > nd->cap) return fail(ICET_B200_E_INVALID, "cloud larger than the node's max_points");
  icet_b200_ctx* c = nd->ctx;
  CK(cudaSetDevice(c->device));
  const int n = cl->n;
  if ((rc = nd->hscan.ensure((size_t)3 * std::max(1, n) * sizeof(float)))) return rc;
  if ((rc = nd->hres.ensure(sizeof(icet_b200_result)))) return rc;
  if ((rc = nd->hpose.ensure(sizeof(icet_b200_pose)))) return rc;
  CK(cudaStreamSynchronize(c->stream));
  if ((rc = ingest_into(c, cl, c->rawbuf[0], (float*)nd->hscan.p, n))) return rc;
  const int np = icet_b200_node_push_device(nd, 1, (const float*)nd->hscan.p, n, (icet_b200_result*)nd->hres.p,
                                            (icet_b200_pose*)nd->hpose.p);
  if (np < 0) return np;
  if (np > 0) {
    if (res) CK(cudaMemcpyAsync(res, nd->hres.p, sizeof(icet_b200_result), cudaMemcpyDeviceToHost, c->stream));
    if (pose) CK(cudaMemcpyAsync(pose, nd->hpose.p, sizeof(icet_b200_pose), cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  rc = check_loop_watchdog(c);
  if (rc) return rc;
  return np;
}

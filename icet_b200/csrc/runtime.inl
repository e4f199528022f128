// icet_b200/csrc/runtime.inl -- host side: device buffers, the context, workspace layout, bin tables and the per-chunk
// launch sequence (run_chunk).  Included by icet_b200.cu INSIDE its anonymous namespace; the context struct lives at
// global scope (it is the opaque type of the C ABI), so this file closes and reopens the namespace around it.
// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
      if (e != cudaSuccess) return fail(ICET_B200_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

constexpr int ICET_LOOP_MAX_PAIRS = 1;  // chunks up to this size run the Gauss-Newton loop as one persistent kernel
constexpr int ICET_DEFAULT_CHUNK = 512;  // pairs per chunk of a device-resident batch (measured r02: 512 -> +1.3 % over 256)
constexpr int ICET_NSLOT = 4;  // staging slots of the host-buffer pipeline
constexpr int ICET_NLANE = 8;  // compute lanes: consecutive chunks rotate over up to eight streams (each with its own
                               // workspace) so that the latency-bound ends of one chunk's kernels overlap the other's

struct icet_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t lanes[ICET_NLANE] = {};  // compute lanes 1.. (lane 0 is `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join[ICET_NLANE] = {};
  cudaEvent_t ev_aux[2] = {};  // single-pair chunks: prepScan2 runs beside the scan-1 kernels on lane 1
  int nlanes_default = 8;
  int nlanes = 8;
  cudaEvent_t ev_copy[ICET_NSLOT] = {};
  cudaEvent_t ev_done[ICET_NSLOT] = {};
  int chunk_pairs = ICET_DEFAULT_CHUNK;
  int host_chunk = 64;  // pairs per chunk of the host-buffer pipeline (upload of chunk k+1 || registration of chunk k)
  int64_t launches = 0;
  int dump_on = 0;
  int sm_count = 148;
  int pass_smem_set = 0;  // dynamic shared memory the pass kernels are currently allowed
  int* loop_dbg[ICET_NLANE] = {};  // watchdog record of the last k_loop launch per lane
  unsigned long long loop_timeout_ns = 20000000000ull;  // ICET_B200_LOOP_TIMEOUT_MS
  float inc_max_sa = INC_MAX_SA, inc_max_sb = INC_MAX_SB;  // rebuild bounds of the incremental loop (ICET_B200_INC_SA / _SB: A/B runs)
  int first_tiles_wide = 1;  // single pair: the tiles of iteration 0 by a GPU-wide k_pass2 launch (ICET_B200_FIRST_TILES)
  int cluster_helpers = 7;  // helper clusters of a single / chained pair (ICET_B200_CLUSTER_HELPERS: A/B runs, 0 = none)
  int cluster_cs = 0, cluster_max = 0, cluster_nT = -1, cluster_nP = -1;  // k_loop_cluster: cluster size, clusters resident at once
  int loop_occ[2] = {0, 0};  // resident blocks per SM of k_loop<PASS_K>, k_loop<PASS_K_SMALL>
  // per-kernel timing (icet_b200_set_profile): events around every launch, summed on request
  int profile_on = 0;
  std::vector<cudaEvent_t> prof_ev;   // pairs (begin, end)
  std::vector<int> prof_id;
  size_t prof_used = 0;
  double prof_ms[ICET_B200_NKERNELS] = {0};
  int64_t prof_n[ICET_B200_NKERNELS] = {0};
  // workspace
  DevBuf ws[ICET_NLANE];  // one slab per compute lane, carved per chunk
  DevBuf zero_ws;   // (part of ws) -- region that must be cleared per chunk is contiguous
  DevBuf edges;     // azE | elE
  int edges_nT = -1, edges_nP = -1;
  DevBuf stage[ICET_NSLOT];  // host-input staging of scans (double buffered)
  DevBuf descbuf[ICET_NSLOT];
  DevBuf lane_desc[ICET_NLANE];        // device-resident batches: pair descriptors of the chunk a lane is working on
  cudaEvent_t ev_lane[ICET_NLANE] = {};  //   ... and "its upload has been consumed"
  DevBuf x0buf[ICET_NSLOT];
  DevBuf resbuf;    // device results for host-facing calls
  DevBuf dumpbuf;
  DevBuf posebuf;
  DevBuf rawbuf[2];  // ingest: device copies of the callers' raw records
  DevBuf planebuf;   // ingest: planes of the two clouds of icet_b200_register_clouds
  void* pinned = nullptr;  // pinned host bounce for results / descriptors
  size_t pinned_cap = 0;
  // dump bookkeeping
  icet_b200_params dump_params{};
  bool dump_valid = false;
  Dump dump_ptrs{};
  // the most recent single-pair chunk (for icet_b200_get_points2)
  bool last_valid = false;
  int last_n2 = 0, last_runlen = 0;
  char last_ck[768];
};

namespace {

struct Carve {
  char* base;
  size_t off = 0;
  explicit Carve(void* b) : base((char*)b) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

// Layout of the chunk workspace.  The first region (cnt1, cntz, cursor, acc) must be zero at chunk start.
// Fixed-point bits of the per-voxel accumulators for a cloud of up to nmax points: offsets up to 2^bits, products up to
// 2^(2 bits), 64-bit sums over at most nmax points: 2 bits + ceil(log2 nmax) <= 62.
inline int fp_bits(int nmax) {
  int lg = 1;
  while ((1ll << lg) < (long long)std::max(nmax, 2)) lg++;
  return std::max(12, std::min(23, (62 - lg) / 2));
}

inline bool huge_path(int n1max, bool shipped) { return !shipped && n1max > 8 * HUGE_MIN; }

size_t carve_chunk(void* base, int P, int ncell, int n1max, int n2max, int runlen, Chunk& ck, size_t* zero_bytes,
                   bool shipped = false) {
  const int vt = (ncell + 31) / 32;
  Carve c(base);
  ck.cnt1 = c.take<int32_t>((size_t)P * ncell);
  ck.cntz = c.take<int32_t>((size_t)P * ncell);
  ck.cursor = c.take<int32_t>((size_t)P * ncell);
  ck.acc = c.take<unsigned long long>((size_t)2 * P * ncell * NQ);
  ck.n2c = c.take<int32_t>((size_t)P);
  ck.nz2 = c.take<int32_t>((size_t)P);
  ck.ticket = c.take<unsigned>(1);
  ck.tiles_done = c.take<unsigned>((size_t)P);
  ck.iter_done = c.take<int>((size_t)P);
  ck.vox_done = c.take<unsigned>((size_t)P * std::max(1, runlen));
  ck.vmask = c.take<unsigned>((size_t)P * ((vt + 31) / 32));
  ck.dbg = c.take<int>(8);
  ck.nhuge = c.take<int32_t>(1);
  ck.hflag = c.take<int32_t>(HUGE_SLOTS);
  ck.hslot = c.take<int32_t>((size_t)P * ncell);
  c.off = (c.off + 255) & ~(size_t)255;
  if (zero_bytes) *zero_bytes = c.off;
  ck.off = c.take<int32_t>((size_t)P * ncell);
  ck.work = c.take<int32_t>((size_t)P * ncell);
  ck.nwork = c.take<int32_t>((size_t)P);
  ck.nbig = c.take<int32_t>((size_t)P);
  ck.pm = c.take<PairMode>((size_t)P);
  ck.rec = c.take<CellRec>((size_t)P * ncell);
  ck.vox = c.take<Vox1>((size_t)P * ncell);
  ck.vlist = c.take<int32_t>((size_t)P * ncell);
  ck.nvox = c.take<int32_t>((size_t)P);
  ck.cellid1 = c.take<int32_t>((size_t)P * n1max);
  ck.r1 = c.take<float>((size_t)P * n1max);
  ck.th1 = c.take<float>((size_t)P * n1max);
  ck.ph1 = c.take<float>((size_t)P * n1max);
  ck.rbuf = c.take<float>((size_t)P * n1max);
  ck.cellg = shipped ? nullptr : c.take<int32_t>((size_t)P * n1max);
  ck.thg = shipped ? nullptr : c.take<float>((size_t)P * n1max);
  ck.phg = shipped ? nullptr : c.take<float>((size_t)P * n1max);
  ck.n1g = c.take<int32_t>((size_t)P);
  ck.kbuf = shipped ? c.take<unsigned long long>((size_t)P * n1max) : nullptr;
  ck.pos1 = shipped ? c.take<int32_t>((size_t)P * n1max) : nullptr;
  ck.pog = c.take<float>((size_t)P * 3 * n2max);
  ck.mrec = c.take<uint2>((size_t)P * n2max);
  ck.anch = c.take<float4>((size_t)P * ncell);
  ck.X = c.take<float>((size_t)P * 6);
  ck.TR = c.take<float>((size_t)P * 12);
  ck.TRprev = c.take<float>((size_t)P * 12);
  ck.J = c.take<float>((size_t)P * 27);
  ck.part = c.take<double>((size_t)P * vt * 28);  // per vox group (k_loop) / per 64-voxel block (split loop)
  ck.hpair = c.take<int32_t>(HUGE_SLOTS);
  ck.hcell = c.take<int32_t>(HUGE_SLOTS);
  ck.hbkt = huge_path(n1max, shipped) ? c.take<int32_t>((size_t)HUGE_SLOTS * 3 * HUGE_NB) : nullptr;
  return (c.off + 255) & ~(size_t)255;
}

int validate(const icet_b200_params* p) {
  if (!p) return fail(ICET_B200_E_INVALID, "params is NULL");
  if (p->runlen < 0 || p->runlen > 10000) return fail(ICET_B200_E_INVALID, "runlen out of range");
  if (p->bins_phi < 1 || p->bins_theta < 1 || p->bins_phi + p->bins_theta > 4096 ||
      (long long)p->bins_phi * p->bins_theta > (1 << 20))
    return fail(ICET_B200_E_INVALID, "bins_phi/bins_theta out of range");
  if (p->n < 1) return fail(ICET_B200_E_INVALID, "n must be >= 1");
  if (!(p->thresh >= 0.f) || !(p->buff >= 0.f)) return fail(ICET_B200_E_INVALID, "thresh/buff must be >= 0");
  return 0;
}

// smallest fp32 a >= 0 with int((double(a)/period)*nb) >= k  (binary search over the fp32 bit patterns,
// which are ordered like the values for a >= 0)
float bin_threshold(int k, double period, int nb) {
  auto f = [&](float a) { return static_cast<int>(((double)a / period) * nb); };
  uint32_t lo = 0, hi = 0x41000000u;  // +0.0f .. 8.0f
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    float a;
    memcpy(&a, &mid, 4);
    if (f(a) >= k) hi = mid; else lo = mid + 1;
  }
  float a;
  memcpy(&a, &lo, 4);
  return a;
}

// device tables that depend only on the bin counts: fp32 box edges (src/icet.cpp:136-139) and the
// exact bin-lookup thresholds (src/icet.cpp:545-546)
int ensure_edges(icet_b200_ctx* ctx, int nT, int nP) {
  if (ctx->edges_nT == nT && ctx->edges_nP == nP) return 0;
  const size_t nbase = (size_t)2 * (nT + nP) + 6;
  std::vector<float> e(((nbase + 3) & ~(size_t)3) + (size_t)6 * (nT + nP + 4));
  float* azE = e.data();
  float* elE = azE + nT + 1;
  float* Tth = elE + nP + 1;
  float* Tph = Tth + nT + 2;
  // src/icet.cpp:136-139: float divide, double multiply, float store
  for (int t = 0; t <= nT; t++) azE[t] = (static_cast<float>(t) / nT) * (2 * M_PI);
  for (int q = 0; q <= nP; q++) elE[q] = (static_cast<float>(q) / nP) * (M_PI);
  for (int k = 0; k <= nT; k++) Tth[k] = bin_threshold(k, 2 * M_PI, nT);
  for (int k = 0; k <= nP; k++) Tph[k] = bin_threshold(k, M_PI, nP);
  Tth[nT + 1] = INFINITY;
  Tph[nP + 1] = INFINITY;
  // bin + box records (bin_box): {T[k], T[k+1], max(T[k], E[k]), min(pred(T[k+1]), E[k+1])}; record nb is empty
  float* rec = e.data() + ((nbase + 3) & ~(size_t)3);
  auto fill = [](float* out, const float* T, const float* E, int nb, double period) {
    const float beyond = std::nextafterf((float)period, INFINITY);
    for (int k = 0; k < nb; k++) {
      out[4 * k + 0] = T[k];
      out[4 * k + 1] = T[k + 1];
      out[4 * k + 2] = std::max(T[k], E[k]);
      out[4 * k + 3] = std::min(std::nextafterf(T[k + 1], -INFINITY), E[k + 1]);
    }
    out[4 * nb + 0] = T[nb];
    out[4 * nb + 1] = beyond;
    out[4 * nb + 2] = INFINITY;
    out[4 * nb + 3] = -INFINITY;
    out[4 * nb + 4] = beyond;
    out[4 * nb + 5] = INFINITY;
    out[4 * nb + 6] = INFINITY;
    out[4 * nb + 7] = -INFINITY;
  };
  fill(rec, Tth, azE, nT, 2 * M_PI);
  fill(rec + 4 * (nT + 2), Tph, elE, nP, M_PI);
  // "sure" intervals of the filtered evaluation (kernels_pass2.cuh): an approximate angle strictly inside
  // (lo + tau, hi - tau) of record k is in bin k AND inside the bin's fp32 box for the exact pipeline as well
  float* sure = rec + 4 * (nT + nP + 4);
  auto fill_sure = [](float* out, const float* r4, int nb, float tau) {
    for (int k = 0; k < nb + 2; k++) {
      const float lo = r4[4 * k + 2], hi = r4[4 * k + 3];
      out[2 * k + 0] = (k < nb) ? std::nextafterf(lo + tau, INFINITY) : INFINITY;
      out[2 * k + 1] = (k < nb) ? std::nextafterf(hi - tau, -INFINITY) : -INFINITY;
    }
  };
  fill_sure(sure, rec, nT, 3.0e-6f);                          // TAU_TH
  fill_sure(sure + 2 * (nT + 2), rec + 4 * (nT + 2), nP, 1.5e-6f);  // TAU_PH
  int rc = ctx->edges.ensure(e.size() * sizeof(float));
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->edges.p, e.data(), e.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->edges_nT = nT;
  ctx->edges_nP = nP;
  return 0;
}

void fill_tables(icet_b200_ctx* ctx, int nT, int nP, const float** azE, const float** elE, icet::BinTable* bth,
                 icet::BinTable* bph, const float** binrec = nullptr) {
  const float* base = (const float*)ctx->edges.p;
  if (binrec) *binrec = base + ((((size_t)2 * (nT + nP) + 6) + 3) & ~(size_t)3);
  *azE = base;
  *elE = base + nT + 1;
  bth->T = base + nT + 1 + nP + 1;
  bth->nb = nT;
  bth->scale = (float)((double)nT / (2 * M_PI));
  bth->amax = (float)(2 * M_PI);
  bth->acap = (float)((nT + 1.25) / ((double)nT / (2 * M_PI)));
  bth->sbin = static_cast<int>(((double)1000.0f / (2 * M_PI)) * nT) % nT;
  bph->T = bth->T + nT + 2;
  bph->nb = nP;
  bph->scale = (float)((double)nP / M_PI);
  bph->amax = (float)M_PI;
  bph->acap = (float)((nP + 1.25) / ((double)nP / M_PI));
  bph->sbin = static_cast<int>(((double)1000.0f / M_PI) * nP) % nP;
}

int prof_events(icet_b200_ctx* ctx, int id, cudaEvent_t* e0, cudaEvent_t* e1) {
  if (ctx->prof_used + 2 > ctx->prof_ev.size()) {
    for (int k = 0; k < 2; k++) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return fail(ICET_B200_E_CUDA, "cudaEventCreate failed");
      ctx->prof_ev.push_back(e);
    }
  }
  *e0 = ctx->prof_ev[ctx->prof_used];
  *e1 = ctx->prof_ev[ctx->prof_used + 1];
  ctx->prof_used += 2;
  ctx->prof_id.push_back(id);
  return 0;
}

// Enqueue the whole registration of one chunk (descriptors already on the device).
int run_chunk(icet_b200_ctx* ctx, const icet_b200_params* p, int P, const PairDesc* d_desc, int n1max, int n2max,
              const float* d_x0, icet_b200_result* d_res, bool dump, int lane = 0) {
  n2max = (n2max + 3) & ~3;  // plane stride of pog: every plane / tile row on a 16-byte boundary (bulk async copies)
  const int nT = p->bins_theta, nP = p->bins_phi, ncell = nT * nP;
  int rc = ensure_edges(ctx, nT, nP);
  if (rc) return rc;
  Chunk ck;
  memset(&ck, 0, sizeof(ck));
  size_t zero_bytes = 0;
  const bool shipped = (p->flags & ICET_B200_FLAG_SHIPPED_ORDER) != 0;
  size_t need = carve_chunk(nullptr, P, ncell, n1max, n2max, p->runlen, ck, &zero_bytes, shipped);
  rc = ctx->ws[lane].ensure(need);
  if (rc) return rc;
  carve_chunk(ctx->ws[lane].p, P, ncell, n1max, n2max, p->runlen, ck, &zero_bytes, shipped);
  ck.desc = d_desc;
  ck.npairs = P; ck.ncell = ncell; ck.nT = nT; ck.nP = nP; ck.n = p->n; ck.runlen = p->runlen;
  ck.flags = p->flags; ck.thresh = p->thresh; ck.buff = p->buff;
  ck.n1max = n1max; ck.n2max = n2max;
  ck.fl1 = 1 << fp_bits(n1max); ck.fs1 = ldexpf(1.0f, fp_bits(n1max) - FPB);
  ck.fl2 = 1 << fp_bits(n2max); ck.fs2 = ldexpf(1.0f, fp_bits(n2max) - FPB);
  fill_tables(ctx, nT, nP, &ck.azE, &ck.elE, &ck.bth, &ck.bph, &ck.binrec);
  ck.x0 = d_x0;
  ck.res = d_res;
  ck.loop_timeout_ns = ctx->loop_timeout_ns;
  ck.inc_max_sa = ctx->inc_max_sa;
  ck.inc_max_sb = ctx->inc_max_sb;
  ck.dump_on = dump ? 1 : 0;
  if (dump) ck.dump = ctx->dump_ptrs;
  cudaStream_t st = lane == 0 ? ctx->stream : ctx->lanes[lane];
  ctx->loop_dbg[lane] = nullptr;  // the lane's workspace is re-carved: a watchdog record of an earlier chunk is gone
  CK(cudaMemsetAsync(ctx->ws[lane].p, 0, zero_bytes, st));
  // (k_prep2 strides over its pair: one block per 256 points for small chunks, at least 32 blocks per pair always)
  const dim3 g1((n1max + 255) / 256, P), g2(std::max(1, std::min((n2max + 255) / 256, std::max(32, 8192 / P))), P);
  const int nblk = (ncell + VOX_THREADS - 1) / VOX_THREADS;
  // shape of the loop kernel: big tiles (16 points per lane) for throughput; small tiles (4 points per lane) when
  // the chunk has too few big tiles to keep every resident block busy for several rounds
  // pass kernels of the split loop: 12 rows per warp tile, 4 blocks / SM, coordinates prefetched 2 rows ahead
  // (measured against 8 / 12 rows with 4-5 blocks and against no prefetch: profiles/r01_pass_variants.txt)
  const int tile1 = pass_tile_points(PASS_K);
  const dim3 gp1((n1max + tile1 - 1) / tile1, P), gp2((n2max + tile1 - 1) / tile1, P);
  const int psm = pass_smem_bytes(nT, nP, PASS_K);
  if (psm > ctx->pass_smem_set) {
    CK(cudaFuncSetAttribute(k_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaFuncSetAttribute(k_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaFuncSetAttribute(k_pass2<>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaFuncSetAttribute(k_pass2<PASS_K_SMALL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaFuncSetAttribute(k_loop<PASS_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaFuncSetAttribute(k_loop<PASS_K_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->loop_occ[0], k_loop<PASS_K>, PASS_THREADS, psm));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->loop_occ[1], k_loop<PASS_K_SMALL>, PASS_THREADS,
                                                     pass_smem_bytes(nT, nP, PASS_K_SMALL)));
    ctx->pass_smem_set = psm;
  }
  // warp tiles: 32*K points.  Small K when the chunk has too few big tiles to keep every resident warp busy for
  // several rounds per iteration (small batches, single-pair latency).
  const int wt_big = 32 * PASS_K;
  const long long big_tiles = (long long)P * ((n2max + wt_big - 1) / wt_big);
  const bool chain = (p->flags & ICET_B200_FLAG_CHAIN_X0) != 0;  // one pair at a time is in flight: latency shape
  const bool small_batch = big_tiles < 4LL * ctx->sm_count * std::max(1, ctx->loop_occ[0]) * PASS_WARPS;
  const bool small = chain || small_batch;
  const int K2 = small ? PASS_K_SMALL : PASS_K;
  const int tiles2 = std::max(1, (n2max + 32 * K2 - 1) / (32 * K2));  // >= 1: tile 0 carries the dropped returns
  const int vt = (ncell + 31) / 32;
  const int psm2 = pass_smem_bytes(nT, nP, K2);
  // LAUNCH(id, kernel<<<...>>>(...)): counts the launch and, when profiling, brackets it with events
#define LAUNCH(id, ...)                                                   \
  do {                                                                    \
    cudaEvent_t e0_ = nullptr, e1_ = nullptr;                             \
    if (ctx->profile_on) {                                                \
      if (prof_events(ctx, id, &e0_, &e1_)) return ICET_B200_E_CUDA;      \
      cudaEventRecord(e0_, st);                                           \
    }                                                                     \
    __VA_ARGS__;                                                          \
    if (e1_) cudaEventRecord(e1_, st);                                    \
    ctx->launches++;                                                      \
  } while (0)
  // Latency shape: kernel -> kernel edges of the set-up sequence are programmatic dependent launches (pdl_prologue)
  const bool pdl = small_batch && !shipped && !ctx->profile_on;
  auto launch_ex = [&](auto kernel, dim3 grid, dim3 block, size_t smem, bool dependent, auto... args) -> cudaError_t {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && dependent) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
  };
  // Latency shape (one pair): prepScan2 does not depend on the scan-1 kernels, so it runs beside them on lane 1.
  const bool prep_aside = P == 1 && n2max > 0 && lane == 0 && !ctx->profile_on && ctx->lanes[1] != nullptr;
  if (prep_aside) {
    CK(cudaEventRecord(ctx->ev_aux[0], st));  // after the workspace has been cleared and the descriptor uploaded
    CK(cudaStreamWaitEvent(ctx->lanes[1], ctx->ev_aux[0], 0));
    k_prep2<<<g2, 256, 0, ctx->lanes[1]>>>(ck);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_aux[1], ctx->lanes[1]));
  }
  if (n1max > 0) LAUNCH(0, CK(launch_ex(k_scan1_bin, g1, dim3(256), 0, false, ck)));  // (follows a memset: plain launch)
  LAUNCH(1, CK(launch_ex(k_cell_scan, dim3(P), dim3(256), 0, n1max > 0, ck)));
  if (n1max > 0) {
    if (shipped) {
      // The row order the reference ends up with (src/icet.cpp:72-83), reproduced on the host from the ranges the
      // device computed: the same index sort by range (std::sort; the reference's std::execution::par falls back
      // to it without TBB) and the same swap loop, which is NOT a valid permutation application.
      // One host round trip per chunk (this mode is for validation / strict reference compatibility, not for speed);
      // the pairs of the chunk are ordered on all host threads.
      std::vector<float> hr((size_t)P * n1max);
      std::vector<PairDesc> hd((size_t)P);
      CK(cudaMemcpyAsync(hr.data(), ck.r1, hr.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(hd.data(), d_desc, hd.size() * sizeof(PairDesc), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<int32_t> pos((size_t)P * n1max, 0);
      auto order_pair = [&](int pr) {
        const int n1 = std::min(hd[pr].n1, n1max);
        const float* r = hr.data() + (size_t)pr * n1max;
        std::vector<int> index((size_t)n1), orig((size_t)n1);
        for (int i = 0; i < n1; i++) index[i] = orig[i] = i;
        std::sort(index.begin(), index.end(), [&](int a, int b) { return r[a] < r[b]; });
        for (int i = 0; i < n1; i++) {
          if (index[i] != i) {
            const int j = index[i];
            std::swap(orig[i], orig[j]);    // points1Spherical.row(i).swap(points1Spherical.row(index[i]))
            std::swap(index[i], index[j]);  // std::swap(index[i], index[index[i]])
          }
        }
        int32_t* ps = pos.data() + (size_t)pr * n1max;
        for (int i = 0; i < n1; i++) ps[orig[i]] = i;
      };
      {
        const int nth = std::max(1, std::min(P, (int)std::thread::hardware_concurrency()));
        if (nth == 1) {
          for (int pr = 0; pr < P; pr++) order_pair(pr);
        } else {
          std::vector<std::thread> th;
          for (int t = 0; t < nth; t++)
            th.emplace_back([&, t]() { for (int pr = t; pr < P; pr += nth) order_pair(pr); });
          for (auto& t : th) t.join();
        }
      }
      CK(cudaMemcpyAsync(ck.pos1, pos.data(), pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      LAUNCH(2, k_off_shipped<<<P, 256, 0, st>>>(ck));
      LAUNCH(2, k_scatter_shipped<<<g1, 256, 0, st>>>(ck));
      LAUNCH(3, k_cluster_shipped<<<dim3(std::max(1, std::min(ncell, 1024)), P), 128, 0, st>>>(ck));
    } else {
    LAUNCH(2, CK(launch_ex(k_scatter, g1, dim3(256), 0, true, ck)));
    if (ck.hbkt) {  // accumulated maps: cells with tens of thousands of ranges are clustered by many CTAs each
      LAUNCH(3, k_huge_init<<<dim3(HUGE_SPLIT, HUGE_SLOTS), 256, 0, st>>>(ck));
      LAUNCH(3, k_huge_hist<<<dim3(HUGE_SPLIT, HUGE_SLOTS), 256, 0, st>>>(ck));
      LAUNCH(3, k_huge_walk<<<HUGE_SLOTS, HW_THREADS, 0, st>>>(ck));
    }
    // one warp per listed cell; enough CTAs to cover a typical work list (~25 % of the cells) in one pass
    int gx = std::max(1, std::min((ncell + CLUSTER_WARPS - 1) / CLUSTER_WARPS,
                                  std::max(32, (ctx->sm_count * 16 + P - 1) / P)));
    LAUNCH(3, CK(launch_ex(k_cluster, dim3(gx, P), dim3(CLUSTER_WARPS * 32), 0, !ck.hbkt, ck)));
    }
    if (small_batch) {  // latency shape: 128 points per warp
      const int tile_s = pass_tile_points(PASS_K_SMALL);
      LAUNCH(4, CK(launch_ex(k_pass<false, PASS_K_SMALL, 3, 1, PASS_K_SMALL>, dim3((n1max + tile_s - 1) / tile_s, P),
                             dim3(PASS_THREADS), psm2, true, ck)));
    } else {
      LAUNCH(4, k_pass<false><<<gp1, PASS_THREADS, psm, st>>>(ck));
    }
  }
  LAUNCH(5, CK(launch_ex(k_fit1, dim3((ncell + 127) / 128, P), dim3(128), 0, n1max > 0, ck)));
  if (prep_aside) CK(cudaStreamWaitEvent(st, ctx->ev_aux[1], 0));
  else if (n2max > 0) LAUNCH(6, k_prep2<<<g2, 256, 0, st>>>(ck));
  const bool use_loop = chain || (p->flags & ICET_B200_FLAG_PERSISTENT_LOOP) ||
                        (!(p->flags & ICET_B200_FLAG_UNFUSED_LOOP) && P <= ICET_LOOP_MAX_PAIRS);
  // latency shape: the loop of a pair inside one thread-block cluster (kernels_cluster.cuh) unless the global-flag
  // persistent kernel is asked for explicitly
  bool use_cluster = use_loop && p->runlen > 0 && !(p->flags & ICET_B200_FLAG_PERSISTENT_LOOP);
  if (p->flags & ICET_B200_FLAG_CLUSTER_LOOP) use_cluster = p->runlen > 0;
  if (use_cluster) {
    if (ctx->cluster_cs == 0 || ctx->cluster_nT != nT || ctx->cluster_nP != nP) {
      CK(cudaFuncSetAttribute(k_loop_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      ctx->cluster_cs = 0;
      for (int cs : {16, 8, 4, 2, 1}) {
        const int csm = cluster_smem_bytes(nT, nP, cs);
        if (cudaFuncSetAttribute(k_loop_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, csm) != cudaSuccess) {
          cudaGetLastError();
          continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(CL_THREADS);
        cfg.dynamicSmemBytes = csm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, k_loop_cluster, &cfg) == cudaSuccess && ncl >= 1) {
          ctx->cluster_cs = cs;
          ctx->cluster_max = ncl;
          ctx->cluster_nT = nT;
          ctx->cluster_nP = nP;
          break;
        }
        cudaGetLastError();
      }
      if (ctx->cluster_cs == 0) use_cluster = false;
    }
  }
  if (use_cluster) {
    const int cs = ctx->cluster_cs;
    const int ncl = chain ? 1 : std::max(1, std::min(P, ctx->cluster_max));
    // single pairs: the tiles of iteration 0 (always a rebuild) run GPU-wide first, the cluster starts at its voxel phase
    const int first_tiles = (!chain && P == 1 && n2max > 0 && !(p->flags & ICET_B200_FLAG_EXACT_PASS) && ctx->first_tiles_wide) ? 1 : 0;
    if (first_tiles) {
      const int tile_s = pass_tile_points(PASS_K_SMALL);
      LAUNCH(7, CK(launch_ex(k_pass2<PASS_K_SMALL, 3>, dim3((n2max + tile_s - 1) / tile_s, P), dim3(PASS_THREADS), psm2,
                             !prep_aside, ck)));
    }
    // single pair / chained pairs: helper clusters share the tiles of the rebuild iterations (kernels_cluster.cuh)
    const int nhelp = ((chain || P == 1) && !(p->flags & ICET_B200_FLAG_EXACT_PASS) && n2max > 0)
                          ? std::max(0, std::min(ctx->cluster_helpers, ctx->cluster_max - 1)) : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((nhelp > 0 ? 1 + nhelp : ncl) * cs);
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = cluster_smem_bytes(nT, nP, cs);
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && (first_tiles || !prep_aside)) ? 2 : 1;  // (after an event wait the edge is an ordinary one)
    LAUNCH(10, CK(cudaLaunchKernelEx(&cfg, k_loop_cluster, ck, first_tiles, nhelp)));
  } else if (!use_loop) {
    // the per-voxel algebra walks the pair's ACTIVE voxels (a sixth of the cells): VOX_GRID blocks per pair instead of
    // one per 64 cells (a 256-pair launch had 7 424 blocks in six waves, most of them without a voxel).  With per-voxel
    // dumps every cell is visited (the dump records the inactive ones as well).
    const int vgrid = dump ? nblk : std::min(nblk, VOX_GRID);
    if (p->runlen > 0) LAUNCH(5, k_vox_list<<<P, 256, 0, st>>>(ck));
    for (int it = 0; it < p->runlen; it++) {
      if (n2max > 0) {
        if (p->flags & ICET_B200_FLAG_EXACT_PASS) LAUNCH(7, k_pass<true><<<gp2, PASS_THREADS, psm, st>>>(ck));
        else LAUNCH(7, k_pass2<><<<gp2, PASS_THREADS, psm, st>>>(ck));
      }
      LAUNCH(8, k_vox2<<<dim3(vgrid, P), VOX_THREADS, 0, st>>>(ck, it));
      LAUNCH(9, k_solve6<<<P, SOLVE_THREADS, 0, st>>>(ck, it, vgrid));
    }
  } else if (p->runlen > 0) {
    // persistent: as many blocks as can be resident (more would only queue behind them)
    const long long tasks = (long long)P * (tiles2 + vt) * p->runlen;
    if (tasks >= (1LL << 32)) return fail(ICET_B200_E_INVALID, "chunk too large: reduce icet_b200_set_chunk");
    const int occ = std::max(1, ctx->loop_occ[small ? 1 : 0]);
    const int grid = (int)std::min<long long>((tasks + PASS_WARPS - 1) / PASS_WARPS, (long long)ctx->sm_count * occ);
    ctx->loop_dbg[lane] = ck.dbg;
    if (small) LAUNCH(10, k_loop<PASS_K_SMALL><<<grid, PASS_THREADS, psm2, st>>>(ck, tiles2, vt));
    else LAUNCH(10, k_loop<PASS_K><<<grid, PASS_THREADS, psm2, st>>>(ck, tiles2, vt));
  }
#undef LAUNCH
  CK(cudaGetLastError());
  static_assert(sizeof(Chunk) <= sizeof(icet_b200_ctx::last_ck), "Chunk too large for last_ck");
  // icet_b200_get_points2 / classify_scan2 refer to the most recent single-pair icet_b200_register call only: that
  // entry point validates the record (together with last_n2 / last_runlen) after this returns; any other chunk --
  // the 1-pair tail of a batch, a node push -- invalidates it (its descriptors point into buffers that get reused)
  ctx->last_valid = false;
  if (P == 1) memcpy(ctx->last_ck, &ck, sizeof(Chunk));
  return 0;
}

// after a blocking call: did a wait inside k_loop give up?  (never expected; turns a would-be hang into an error)
int check_loop_watchdog(icet_b200_ctx* ctx) {
  for (int lane = 0; lane < ICET_NLANE; lane++) {
    if (!ctx->loop_dbg[lane]) continue;
    int d[8];
    CK(cudaMemcpy(d, ctx->loop_dbg[lane], sizeof(d), cudaMemcpyDeviceToHost));
    ctx->loop_dbg[lane] = nullptr;
    if (d[0])
      return fail(ICET_B200_E_CUDA, "persistent loop kernel: wait timed out (kind " + std::to_string(d[1]) + ", pair " +
                                        std::to_string(d[2]) + ", iteration " + std::to_string(d[3]) + ", seen " +
                                        std::to_string(d[4]) + ", need " + std::to_string(d[5]) + ", ticket " +
                                        std::to_string(d[6]) + ")");
  }
  return 0;
}

int ensure_pinned(icet_b200_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return 0;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_cap = 0;
  CK(cudaMallocHost(&ctx->pinned, bytes));
  ctx->pinned_cap = bytes;
  return 0;
}

int ensure_dump(icet_b200_ctx* ctx, const icet_b200_params* p) {
  const size_t ncell = (size_t)p->bins_phi * p->bins_theta, rl = (size_t)std::max(1, p->runlen);
  Carve c(nullptr);
  auto lay = [&](Carve& cv, Dump& d) {
    d.nin1 = cv.take<int32_t>(ncell); d.has1 = cv.take<uint8_t>(ncell); d.mu1 = cv.take<float>(ncell * 3);
    d.sigma1 = cv.take<float>(ncell * 9); d.evec1 = cv.take<float>(ncell * 9); d.eval1 = cv.take<float>(ncell * 3);
    d.lmask = cv.take<uint8_t>(ncell * 3);
    d.cnt2 = cv.take<int32_t>(rl * ncell); d.nin2 = cv.take<int32_t>(rl * ncell); d.used2 = cv.take<uint8_t>(rl * ncell);
    d.mu2 = cv.take<float>(rl * ncell * 3); d.sigma2 = cv.take<float>(rl * ncell * 9);
    d.Xit = cv.take<float>(rl * 6); d.HTWH = cv.take<float>(rl * 36); d.HTWdz = cv.take<float>(rl * 6);
    d.TRit = cv.take<float>(rl * 12); d.testpts = cv.take<float>(ncell * 18);
    d.tl = cv.take<unsigned long long>(rl * 16 + 6144);  // + begin / end / mid of up to 2048 tiles of iteration 3
  };
  Dump tmp;
  lay(c, tmp);
  size_t need = c.off + 256;
  int rc = ctx->dumpbuf.ensure(need);
  if (rc) return rc;
  Carve c2(ctx->dumpbuf.p);
  lay(c2, ctx->dump_ptrs);
  CK(cudaMemsetAsync(ctx->dumpbuf.p, 0, need, ctx->stream));
  return 0;
}


// icet_b200/csrc/icet_b200.cu -- sm_100a kernels and the C ABI (include/icet_b200.h) of the
// B200-native ICET registration hot path.
//
// Replaces the work of the reference constructor ICET::ICET (src/icet.cpp:29-63):
//   fitScan1  (:68-107)  -> k_scan1_bin, k_cell_scan, k_scatter, k_cluster, k_pass<false>, k_fit1
//   prepScan2 (:254-277) -> k_prep2
//   fitScan2  (:372-436) -> per iteration: k_pass<true>, k_vox2, k_solve6 (batches), or all iterations in the
//                           persistent k_loop (single pairs, chained pairs)
// All pairs of a chunk advance together through these kernels (bulk-synchronous over the batch),
// everything stays on the device between iterations; the host only enqueues.
// callers.cuh / callers_abi.inl (included below) hold the thin callers around the path: range filter, pose
// bookkeeping, FIFO map, cloud re-expression, ingest.
//
// Determinism: per-voxel statistics are accumulated as 64-bit fixed-point integers (run sums in registers, then
// RED.64 to L2), so results do not depend on thread / block / GPU partitioning or on atomic ordering.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/icet_b200.h"
#include "icet_math.cuh"
#include "synth.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ICET_B200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

constexpr int FPB = 21;                 // fixed-point: |d * scale| <= 2^FPB
constexpr int FP_LIM = (1 << FPB);
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t F_STAT1 = 1u;        // scan-1 statistics wanted for this cell
constexpr uint32_t F_ACTIVE2 = 2u;      // voxel takes part in the scan-2 loop
constexpr int SORT_SMEM = 4096;         // cells up to this many points are sorted in shared memory
constexpr int NQ = 12;                  // accumulator words per cell
constexpr int CLUSTER_WARPS = 4;
constexpr int WSORT_MAX = 1024;         // cells up to this many non-zero ranges are sorted by one warp in registers

struct PairDesc {
  const float* s1;
  const float* s2;
  int n1, ld1, n2, ld2;
};

struct __align__(16) CellRec {  // read by the pass kernels: first half by every point, second half per run of inside points
  float inner, outer;           // clusterBounds columns 4,5 (src/icet.cpp:149)
  uint32_t flags;
  float scale;                  // power-of-two scale of the voxel's fixed-point frame
  float refx, refy, refz;       // reference point of the fixed-point frame
  int32_t cnt1;                 // points of scan 1 in this angular bin
};

struct Vox1 {     // scan-1 Gaussian, constants of the iteration loop (sigma1/mu1/U/L, include/icet.h:89-94)
  double mu[3];
  double S1n[6];  // sigma1 / (cnt1 - 1), upper triangle xx xy xz yy yz zz
  double LV[9];   // L * U^T = L * V  (rows of V, zeroed where L is 0)
  int lmask;      // bit k: L(k,k) == 1
  int pad;
};

struct Dump {  // optional per-voxel recording (device memory), single-pair debugging only
  int32_t* nin1; uint8_t* has1; float* mu1; float* sigma1; float* evec1; float* eval1; uint8_t* lmask;
  int32_t* cnt2; int32_t* nin2; uint8_t* used2; float* mu2; float* sigma2; float* Xit; float* HTWH; float* HTWdz;
  unsigned long long* tl;  // [runlen][16] globaltimer stamps of the loop kernel (debug) + per-tile stamps of iteration 3
};

struct Chunk {  // everything a kernel needs, passed by value
  const PairDesc* desc;
  int npairs, ncell, nT, nP, n, runlen, flags;
  float thresh, buff;
  int n1max, n2max;
  // scan 1
  int32_t* cellid1;  // [P][n1max]
  float* r1;         // [P][n1max]
  float* th1;        // [P][n1max]  theta, phi of scan 1 (K1 -> K3: the second pass over scan 1 does not redo the
  float* ph1;        // [P][n1max]  spherical conversion)
  float* rbuf;       // [P][n1max]  non-zero ranges grouped by cell
  unsigned long long* kbuf;  // [n1max]  ICET_B200_FLAG_SHIPPED_ORDER: (row position << 32 | range bits) grouped by cell
  int32_t* pos1;     // [n1max]  ... position of every row of scan 1 in the reference's shipped row order
  int32_t* cnt1;     // [P][ncell]
  int32_t* cntz;     // [P][ncell]  zero-range points
  int32_t* off;      // [P][ncell]
  int32_t* cursor;   // [P][ncell]
  int32_t* work;     // [P][ncell]  cells with cnt1 >= n
  int32_t* nwork;    // [P]
  int32_t* nbig;     // [P]  cells of the work list with more than WSORT_MAX non-zero ranges
  CellRec* rec;      // [P][ncell]
  unsigned long long* acc;  // [P][ncell][NQ]
  Vox1* vox;         // [P][ncell]
  const float* azE;  // [nT+1] float azimuth bin edges  (src/icet.cpp:136-137)
  const float* elE;  // [nP+1] float elevation bin edges (src/icet.cpp:138-139)
  icet::BinTable bth, bph;  // exact bin lookup tables (src/icet.cpp:545-546)
  const float* binrec;      // [(nT+1) + (nP+1)][4] bin + box records of the pass kernels (see bin_box)
  float* TR;         // [P][12] translation (3) and rotation R(X) (9) of the current iteration
  float* TRprev;     // [P][12] transform the LAST iteration used (the reference's public `points2`)
  float* J;          // [P][27] get_H derivative matrices Jx | Jy | Jz of the current iteration
  double* part;      // [P][nblk][NRED] per-block partial sums of the voxel contributions
  // scan 2
  float* pog;        // [P][3][n2max]  points2_OG without the dropped returns (compacted, any order)
  int32_t* n2c;      // [P] points stored in pog
  int32_t* nz2;      // [P] dropped returns of scan 2 (points2_OG == 0)
  float* X;          // [P][6]
  const float* x0;   // [P][6] or null
  icet_b200_result* res;  // [P] device
  // control words of the persistent Gauss-Newton kernel (k_loop)
  unsigned* ticket;      // [1]  next task
  unsigned* tiles_done;  // [P]  scan-2 tiles finished so far (all iterations)
  int* iter_done;        // [P]  iterations whose solve has been published
  unsigned* vox_done;    // [P][runlen] vox tasks finished per iteration
  unsigned* vmask;       // [P][ceil(vt/32)] vox groups that wrote a partial sum (current iteration)
  int* dbg;              // [8] watchdog record of k_loop: {tripped, kind, pair, iter, seen, need, ticket, -}
  Dump dump;
  int dump_on;
};

// Atomics on Chunk memory, with the address space spelled out.  Kernels that hand the Chunk to non-inlined device
// functions by reference (k_loop) keep it in local memory, and the compiler then no longer knows which address space
// the pointers it loads from there refer to: atomicAdd() becomes a GENERIC atomic that waits for a predicate from the
// memory system (one L2 round trip each instead of a fire-and-forget RED) plus a shared-memory CAS fallback.
__device__ __forceinline__ void red_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add(unsigned* p, unsigned v) {
  asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or(unsigned* p, unsigned v) {
  asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned atom_add(unsigned* p, unsigned v) {
  unsigned r;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}

// ----------------------------------------------------------------------------------------------
// Fixed-point accumulators acc[cell][NQ] (64-bit integers, RED.64 to L2):
//   q[0] += points in the angular bin, q[1] += points inside the cluster box,
//   q[2..4] += sum d, q[5..10] += sum d d^T (xx xy xz yy yz zz),   d = round((p - ref) * scale)
// Exact integer arithmetic => the sums do not depend on the order or grouping of the additions.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_of(const Chunk& ck, float th, float ph, int& bt, int& bp) {
  bt = icet::bin_lookup(th, ck.bth, 2 * M_PI);
  bp = icet::bin_lookup(ph, ck.bph, M_PI);
}
// ----------------------------------------------------------------------------------------------
// K1: scan 1 -> spherical, cell index, per-cell histogram.
// utils::cartesianToSpherical (src/utils.cpp:93-119) + sortSphericalCoordinates (src/icet.cpp:534-554)
// ----------------------------------------------------------------------------------------------
constexpr int32_t CELL_INBOX = 0x40000000;  // cellid1 flag: the point passes the fp32 az / el box test of its own bin
__device__ __forceinline__ int bin_box(float a, const float4* rec, const icet::BinTable& bt, bool& inbox);

__global__ void __launch_bounds__(256) k_scan1_bin(const Chunk ck) {
  const int pair = blockIdx.y;
  const PairDesc d = ck.desc[pair];
  if ((int)(blockIdx.x * blockDim.x) >= d.n1) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int cell = -1;
  bool zero = false;
  if (i < d.n1) {
    float x = __ldg(d.s1 + i), y = __ldg(d.s1 + d.ld1 + i), z = __ldg(d.s1 + 2 * (size_t)d.ld1 + i);
    float r, th, ph;
    icet::c2s(x, y, z, r, th, ph);
    // bin and box test in one look-up (same records as the pass kernels, read through L1 here)
    const float4* tth = reinterpret_cast<const float4*>(ck.binrec);
    const float4* tph = tth + ck.nT + 2;
    bool bt_in, bp_in;
    const int bt = bin_box(th, tth, ck.bth, bt_in);
    const int bp = bin_box(ph, tph, ck.bph, bp_in);
    cell = ck.nT * bp + bt;
    zero = (r == 0.0f);
    const size_t o = (size_t)pair * ck.n1max + i;
    ck.cellid1[o] = cell | ((bt_in && bp_in) ? CELL_INBOX : 0);
    ck.r1[o] = r;
    ck.th1[o] = th;
    ck.ph1[o] = ph;
  }
  // warp-aggregated histogram
  const int lane = threadIdx.x & 31;
  const unsigned act = __ballot_sync(FULL, cell >= 0);
  if (cell >= 0) {
    const unsigned m = __match_any_sync(act, cell);
    const unsigned mz = __ballot_sync(m, zero);
    if (lane == __ffs(m) - 1) {
      atomicAdd(&ck.cnt1[(size_t)pair * ck.ncell + cell], __popc(m));
      if (mz) atomicAdd(&ck.cntz[(size_t)pair * ck.ncell + cell], __popc(mz));
    }
  }
}

// ----------------------------------------------------------------------------------------------
// K2a: per pair: exclusive scan of the non-zero counts (offsets into rbuf), work list of cells with
// cnt1 >= n (src/icet.cpp:115), default cell records (the else-branch :243-251: inner = outer = 0).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_scan(const Chunk ck) {
  const int pair = blockIdx.x;
  __shared__ int s_ws[8], s_ww[8];
  const int per = (ck.ncell + 255) / 256;
  const int c0 = threadIdx.x * per, c1 = min(ck.ncell, c0 + per);
  const int32_t* cnt1 = ck.cnt1 + (size_t)pair * ck.ncell;
  const int32_t* cntz = ck.cntz + (size_t)pair * ck.ncell;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int s = 0, w = 0, big = 0;
  for (int c = c0; c < c1; c++) {
    const int m = cnt1[c] - cntz[c];
    s += m;
    w += (cnt1[c] >= ck.n) ? 1 : 0;
    big += (cnt1[c] >= ck.n && m > WSORT_MAX) ? 1 : 0;
  }
  {
    const int anybig = __syncthreads_count(big > 0);  // (number of threads that own a big cell: only zero / non-zero matters)
    if (threadIdx.x == 0) ck.nbig[pair] = anybig;
  }
  // exclusive block scan of (s, w): inclusive warp scans, then the warp totals
  int is = s, iw = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int ts = __shfl_up_sync(FULL, is, o), tw = __shfl_up_sync(FULL, iw, o);
    if (lane >= o) { is += ts; iw += tw; }
  }
  if (lane == 31) { s_ws[warp] = is; s_ww[warp] = iw; }
  // meanwhile: the per-pair state.  Warp 1 owns the result record, thread 0 the transform.
  if (warp == 1) {
    icet_b200_result* R = ck.res + pair;
    // chained pairs (ICET_B200_FLAG_CHAIN_X0, odometry.cpp:82): x0 holds ONE seed, that of pair 0; the later pairs are
    // seeded by the last solve of their predecessor (chain_seed_next).  Without iterations the seed is the answer.
    const bool chain = (ck.flags & ICET_B200_FLAG_CHAIN_X0) != 0;
    const bool seeded = ck.x0 && (!chain || pair == 0 || ck.runlen == 0);
    if (lane < 6) {
      const float x = seeded ? ck.x0[(chain ? 0 : pair * 6) + lane] : 0.f;
      ck.X[pair * 6 + lane] = x;
      R->X[lane] = x;
      R->pred_stds[lane] = 0.f;
    }
    for (int k = lane; k < 36; k += 32) R->Q[k] = 0.f;
    if (lane == 6) { R->status = 0; R->n_gauss1 = 0; R->n_used = 0; R->n_dropped = 0; R->cond = 0.f; }
    if (lane >= 7 && lane < 10) R->reserved[lane - 7] = 0;
  }
  if (threadIdx.x == 0) {
    const bool chain = (ck.flags & ICET_B200_FLAG_CHAIN_X0) != 0;
    const bool seeded = ck.x0 && (!chain || pair == 0 || ck.runlen == 0);
    float X[6];
    for (int k = 0; k < 6; k++) X[k] = seeded ? ck.x0[(chain ? 0 : pair * 6) + k] : 0.f;
    float* TR = ck.TR + (size_t)pair * 12;
    TR[0] = X[0]; TR[1] = X[1]; TR[2] = X[2];
    icet::rotR(X[3], X[4], X[5], TR + 3);
    icet::getH_J(X[3], X[4], X[5], ck.J + (size_t)pair * 27);
    for (int k = 0; k < 12; k++) ck.TRprev[(size_t)pair * 12 + k] = TR[k];
  }
  __syncthreads();
  int a = is - s, b = iw - w;
  for (int q = 0; q < warp; q++) { a += s_ws[q]; b += s_ww[q]; }
  if (threadIdx.x == 255) ck.nwork[pair] = b + w;
  for (int c = c0; c < c1; c++) {
    ck.off[(size_t)pair * ck.ncell + c] = a;
    a += cnt1[c] - cntz[c];
    if (cnt1[c] >= ck.n) ck.work[(size_t)pair * ck.ncell + (b++)] = c;
    CellRec rc;
    rc.inner = 0.f; rc.outer = 0.f; rc.refx = rc.refy = rc.refz = 0.f; rc.scale = 0.f;
    rc.flags = 0; rc.cnt1 = cnt1[c];
    ck.rec[(size_t)pair * ck.ncell + c] = rc;
  }
}

// K2b: group the non-zero ranges by cell
__global__ void __launch_bounds__(256) k_scatter(const Chunk ck) {
  const int pair = blockIdx.y;
  const PairDesc d = ck.desc[pair];
  if ((int)(blockIdx.x * blockDim.x) >= d.n1) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int cell = -1;
  float r = 0.f;
  if (i < d.n1) {
    r = ck.r1[(size_t)pair * ck.n1max + i];
    if (r != 0.0f) cell = ck.cellid1[(size_t)pair * ck.n1max + i] & ~CELL_INBOX;
  }
  const int lane = threadIdx.x & 31;
  const unsigned act = __ballot_sync(FULL, cell >= 0);
  if (cell >= 0) {
    const unsigned m = __match_any_sync(act, cell);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&ck.cursor[(size_t)pair * ck.ncell + cell], __popc(m));
    base = __shfl_sync(m, base, leader);
    const int rank = __popc(m & ((1u << lane) - 1));
    ck.rbuf[(size_t)pair * ck.n1max + ck.off[(size_t)pair * ck.ncell + cell] + base + rank] = r;
  }
}

// ----------------------------------------------------------------------------------------------
// K2c: per cell with cnt1 >= n: sort the ranges ascending (what src/icet.cpp:71-83 intends) and run
// ICET::findCluster (src/icet.cpp:557-607) on them; write the cell record.
// ----------------------------------------------------------------------------------------------
// Bitonic-style network with all comparators ascending ("mirror" first step): works for any m,
// indices >= m behave as +inf and are never touched.
template <class Ptr>
__device__ inline void block_sort_asc(Ptr a, int m) {
  int p2 = 1;
  while (p2 < m) p2 <<= 1;
  for (int k = 2; k <= p2; k <<= 1) {
    // mirror step: i with partner i ^ (k-1)
    for (int t = threadIdx.x; t < p2 / 2; t += blockDim.x) {
      int blk = t / (k / 2), o = t % (k / 2);
      int i = blk * k + o, j = blk * k + (k - 1 - o);
      if (j < m) {
        auto x = a[i];
        auto y = a[j];
        if (y < x) { a[i] = y; a[j] = x; }
      }
    }
    __syncthreads();
    for (int j2 = k / 4; j2 >= 1; j2 >>= 1) {
      for (int t = threadIdx.x; t < p2 / 2; t += blockDim.x) {
        int i = (t / j2) * (2 * j2) + (t % j2), j = i + j2;
        if (j < m) {
          auto x = a[i];
          auto y = a[j];
          if (y < x) { a[i] = y; a[j] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// findCluster over the virtual sequence seq = [0 x nz, a[0..m)] (ascending); executed by warp 0,
// every lane computes the same result.  A "break" at position i means seq[i] does not extend the
// current run (reference :572); the run before a break is returned if it has >= n points (:577-582,
// no zero check), the run that reaches the end of the data goes through the zero check (:592-603).
template <class Ptr>
__device__ inline void find_cluster_warp(Ptr a, int m, int nz, int n, float thresh, float buff, float& inner,
                                         float& outer) {
  const int lane = threadIdx.x & 31;
  const int total = nz + m;
  auto seq = [&](int i) -> float { return i < nz ? 0.0f : a[i - nz]; };
  int start = 0;  // first element of the current run; the nz leading zeros never break (thresh >= 0)
  bool found = false;
  float fi = 0.f, fo = 0.f;
  for (int base = nz; base < total && !found; base += 32) {
    const int i = base + lane;
    bool brk = false;
    if (i < total && i > 0) brk = !(fabsf(seq(i - 1) - seq(i)) <= thresh);
    unsigned mask = __ballot_sync(FULL, brk);
    while (mask && !found) {
      const int pos = base + __ffs(mask) - 1;
      mask &= mask - 1;
      if (pos - start >= n) {
        fi = seq(start) - buff;
        fo = seq(pos - 1) + buff;
        found = true;
      } else {
        start = pos;
      }
    }
  }
  if (!found && total > 0 && total - start >= n) {
    if (seq(start) != 0.0f) {
      fi = seq(start) - buff;
      fo = seq(total - 1) + buff;
    }
  }
  inner = fi;
  outer = fo;
}

// writes the cell record of a clustered cell (clusterBounds columns 4,5 + the voxel's fixed-point frame)
__device__ inline void write_cluster_rec(const Chunk& ck, int pair, int cell, int cnt, float inner, float outer) {
  const int bt = cell % ck.nT, bp = cell / ck.nT;
  CellRec rc;
  rc.inner = inner;
  rc.outer = outer;
  rc.cnt1 = cnt;
  rc.flags = ((double)outer > 0.1) ? F_STAT1 : 0u;  // `outerDistance > 0.1` src/icet.cpp:158
  // fixed-point frame of the voxel: reference point = centre of the spherical box, scale from
  // a bound on the box diameter
  const float azl = ck.azE[bt], azh = ck.azE[bt + 1], ell = ck.elE[bp], elh = ck.elE[bp + 1];
  const float rm = 0.5f * (inner + outer), tm = 0.5f * (azl + azh), pm = 0.5f * (ell + elh);
  icet::s2c(rm, tm, pm, rc.refx, rc.refy, rc.refz);
  float D = (outer - inner) + fabsf(outer) * ((azh - azl) + (elh - ell));
  int e;
  frexpf(fmaxf(D, 1e-20f), &e);          // D < 2^e
  rc.scale = ldexpf(1.0f, FPB - e - 1);  // |d| <= D  =>  |d*scale| < 2^(FPB-1)
  ck.rec[(size_t)pair * ck.ncell + cell] = rc;
}

// Bitonic sort of 32*EPL floats held EPL per lane; element index = lane*EPL + j.  Compare-exchange distances
// below EPL stay inside a lane (registers), larger ones are one shuffle per element.  Fully unrolled.
template <int EPL>
__device__ __forceinline__ void warp_sort_regs(float (&v)[EPL]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
    for (int d = k >> 1; d > 0; d >>= 1) {
      if (d >= EPL) {
        const int ld = d / EPL;                                             // partner lane distance
        const bool asc = (k >= 32 * EPL) || ((lane & (k / EPL)) == 0);
        const bool keep_min = (((lane & ld) == 0) == asc);
#pragma unroll
        for (int j = 0; j < EPL; j++) {
          const float o = __shfl_xor_sync(FULL, v[j], ld);
          v[j] = keep_min ? fminf(v[j], o) : fmaxf(v[j], o);
        }
      } else {
#pragma unroll
        for (int j = 0; j < EPL; j++) {
          if ((j & d) == 0) {
            // direction bit k of the element index: in the register index for k < EPL, in the lane above
            const bool asc = (k >= 32 * EPL) || (k < EPL ? ((j & k) == 0) : ((lane & (k / EPL)) == 0));
            const float lo = fminf(v[j], v[j ^ d]), hi = fmaxf(v[j], v[j ^ d]);
            v[j] = asc ? lo : hi;
            v[j ^ d] = asc ? hi : lo;
          }
        }
      }
    }
  }
}

// loads the m ranges of a cell (any order), sorts them in registers and leaves them ascending in the warp's
// shared-memory row (index i stored at i + i/32: conflict-free for the blocked write and for consecutive reads)
template <int EPL>
__device__ __forceinline__ void warp_sort_cell(const float* __restrict__ g, int m, float* srow) {
  const int lane = threadIdx.x & 31;
  float v[EPL];
#pragma unroll
  for (int j = 0; j < EPL; j++) {
    const int i = j * 32 + lane;  // coalesced; which register an unsorted value lands in is irrelevant
    v[j] = i < m ? __ldg(g + i) : INFINITY;
  }
  warp_sort_regs<EPL>(v);
#pragma unroll
  for (int j = 0; j < EPL; j++) {
    const int i = lane * EPL + j;
    srow[i + (i >> 5)] = v[j];
  }
  __syncwarp();
}

struct PaddedRow {  // view of a shared-memory row written by warp_sort_cell
  const float* p;
  __device__ __forceinline__ float operator[](int i) const { return p[i + (i >> 5)]; }
};

// K2c: ONE WARP per cell with cnt1 >= n (a cell of a 64-ring scan holds ~260 ranges): register bitonic sort, then
// ICET::findCluster on the sorted row.  Cells with more than WSORT_MAX ranges take the CTA path at the end of the kernel.
__global__ void __launch_bounds__(CLUSTER_WARPS * 32) k_cluster(const Chunk ck) {
  const int pair = blockIdx.y;
  constexpr int ROW = WSORT_MAX + WSORT_MAX / 32;
  constexpr int SM_FLOATS = CLUSTER_WARPS * ROW > SORT_SMEM ? CLUSTER_WARPS * ROW : SORT_SMEM;
  __shared__ float s_all[SM_FLOATS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* srow = s_all + warp * ROW;
  const int nw = ck.nwork[pair];
  for (int w = blockIdx.x * CLUSTER_WARPS + warp; w < nw; w += gridDim.x * CLUSTER_WARPS) {
    const int cell = ck.work[(size_t)pair * ck.ncell + w];
    const int cnt = ck.cnt1[(size_t)pair * ck.ncell + cell];
    const int nz = ck.cntz[(size_t)pair * ck.ncell + cell];
    const int m = cnt - nz;
    if (m > WSORT_MAX) continue;
    const float* g = ck.rbuf + (size_t)pair * ck.n1max + ck.off[(size_t)pair * ck.ncell + cell];
    if (m <= 128) warp_sort_cell<4>(g, m, srow);
    else if (m <= 256) warp_sort_cell<8>(g, m, srow);
    else if (m <= 512) warp_sort_cell<16>(g, m, srow);
    else warp_sort_cell<32>(g, m, srow);
    float inner, outer;
    find_cluster_warp(PaddedRow{srow}, m, nz, ck.n, ck.thresh, ck.buff, inner, outer);
    if (lane == 0) write_cluster_rec(ck, pair, cell, cnt, inner, outer);
    __syncwarp();
  }
  // The big cells (more than WSORT_MAX ranges: an accumulated map as scan 1; k_cell_scan counted them), the whole CTA per
  // cell.  findCluster only needs to know where consecutive SORTED ranges are more than `thresh` apart, so no sort:
  // the ranges are dropped into buckets half a threshold wide (count, min, max per bucket; ranges inside one bucket
  // are closer than the threshold by construction), windows of NBKT buckets are walked in ascending order, and the
  // walk stops at the first run that qualifies.  O(m) per window instead of the O(m log^2 m) of a bitonic network
  // (a 2 M-point map: 217 ms -> well under 1 ms).  Fallback for thresholds / ranges the buckets cannot cover: sort.
  if (ck.nbig[pair] == 0) return;  // block-uniform
  __syncthreads();
  constexpr int NBKT = (SM_FLOATS - 8) / 3;
  int* b_cnt = reinterpret_cast<int*>(s_all);
  int* b_min = b_cnt + NBKT;
  int* b_max = b_min + NBKT;
  int* s_ctl = b_max + NBKT;  // [0] points seen so far, [1] done, [2..3] result bits
  for (int w = blockIdx.x; w < nw; w += gridDim.x) {
    const int cell = ck.work[(size_t)pair * ck.ncell + w];
    const int cnt = ck.cnt1[(size_t)pair * ck.ncell + cell];
    const int nz = ck.cntz[(size_t)pair * ck.ncell + cell];
    const int m = cnt - nz;
    if (m <= WSORT_MAX) continue;  // block-uniform
    float* g = ck.rbuf + (size_t)pair * ck.n1max + ck.off[(size_t)pair * ck.ncell + cell];
    // largest range of the cell (the ranges are >= 0: their bit patterns order like the values)
    __syncthreads();
    if (threadIdx.x == 0) s_ctl[0] = 0;
    __syncthreads();
    {
      int mx = 0;
      for (int i0 = threadIdx.x; i0 < m; i0 += 8 * blockDim.x) {  // eight loads in flight per thread
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * blockDim.x;
          v[u] = i < m ? __float_as_int(g[i]) : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) mx = max(mx, v[u]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
      if (lane == 0) atomicMax(&s_ctl[0], mx);
    }
    __syncthreads();
    const float rmax = __int_as_float(s_ctl[0]);
    const float wdt = 0.5f * ck.thresh, inv_w = 1.0f / wdt;
    const bool buckets_ok = ck.thresh > 1e-6f && rmax * inv_w < 64.0f * NBKT;  // false for inf / NaN as well
    float inner = 0.f, outer = 0.f;
    if (buckets_ok) {
      // state of the walk (warp 0, uniform across its lanes)
      int idx = nz, start = 0;
      float start_val = 0.f, prev_val = 0.f;
      bool found = false;
      __syncthreads();
      if (threadIdx.x == 0) { s_ctl[0] = 0; s_ctl[1] = 0; }
      for (int win = 0;; win++) {
        for (int k = threadIdx.x; k < NBKT; k += blockDim.x) { b_cnt[k] = 0; b_min[k] = 0x7f800000; b_max[k] = 0; }
        __syncthreads();
        if (s_ctl[1]) break;  // found, or every range has been walked
        const int lo = win * NBKT;
        for (int i0 = threadIdx.x; i0 < m; i0 += 8 * blockDim.x) {  // eight loads in flight per thread
          float rv[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int i = i0 + u * blockDim.x;
            rv[u] = i < m ? g[i] : -1.0f;
          }
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const float r = rv[u];
            const int k = __float2int_rd(r * inv_w) - lo;
            if (r >= 0.0f && k >= 0 && k < NBKT) {
              atomicAdd(&b_cnt[k], 1);
              atomicMin(&b_min[k], __float_as_int(r));
              atomicMax(&b_max[k], __float_as_int(r));
            }
          }
        }
        __syncthreads();
        if (warp == 0) {
          int seen = 0;
          for (int base = 0; base < NBKT && !found; base += 32) {
            const int k = base + lane;
            const int c = k < NBKT ? b_cnt[k] : 0;
            unsigned mask = __ballot_sync(FULL, c > 0);
            while (mask && !found) {
              const int kb = base + __ffs(mask) - 1;
              mask &= mask - 1;
              const int bc = b_cnt[kb];
              const float mn = __int_as_float(b_min[kb]), mxv = __int_as_float(b_max[kb]);
              if (idx > 0) {
                if (!(fabsf(prev_val - mn) <= ck.thresh)) {  // a break in front of this bucket (reference :572)
                  if (idx - start >= ck.n) {
                    inner = start_val - ck.buff;             // :577-582, no zero check
                    outer = prev_val + ck.buff;
                    found = true;
                    break;
                  }
                  start = idx;
                  start_val = mn;
                }
              } else {
                start_val = mn;
              }
              idx += bc;
              seen += bc;
              prev_val = mxv;
            }
          }
          if (lane == 0) {
            s_ctl[0] += seen;
            if (found || s_ctl[0] >= m) s_ctl[1] = 1;
          }
        }
        __syncthreads();
      }
      if (warp == 0 && !found && nz + m - start >= ck.n && start_val != 0.0f) {  // end of the data (:592-603)
        inner = start_val - ck.buff;
        outer = prev_val + ck.buff;
      }
    } else {
      float* s_r = s_all;
      if (m <= SORT_SMEM) {
        for (int i = threadIdx.x; i < m; i += blockDim.x) s_r[i] = g[i];
        __syncthreads();
        block_sort_asc(s_r, m);
        if (threadIdx.x < 32) find_cluster_warp(s_r, m, nz, ck.n, ck.thresh, ck.buff, inner, outer);
      } else {
        __syncthreads();
        block_sort_asc(g, m);
        if (threadIdx.x < 32) find_cluster_warp(g, m, nz, ck.n, ck.thresh, ck.buff, inner, outer);
      }
    }
    if (threadIdx.x == 0) write_cluster_rec(ck, pair, cell, cnt, inner, outer);
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// ICET_B200_FLAG_SHIPPED_ORDER (single pair, validation): clustering in the row order the reference ends up with
// after its broken permutation loop (src/icet.cpp:72-83); ck.pos1 holds that position for every row of scan 1.
// Zero ranges are ordinary members of the sequence here.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_off_shipped(const Chunk ck) {  // offsets of ALL points per cell
  const int pair = blockIdx.x;
  __shared__ int s_part[256];
  const int per = (ck.ncell + 255) / 256;
  const int c0 = threadIdx.x * per, c1 = min(ck.ncell, c0 + per);
  const int32_t* cnt1 = ck.cnt1 + (size_t)pair * ck.ncell;
  int s = 0;
  for (int c = c0; c < c1; c++) s += cnt1[c];
  s_part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0;
    for (int t = 0; t < 256; t++) { const int v = s_part[t]; s_part[t] = a; a += v; }
  }
  __syncthreads();
  int a = s_part[threadIdx.x];
  for (int c = c0; c < c1; c++) {
    ck.off[(size_t)pair * ck.ncell + c] = a;
    ck.cursor[(size_t)pair * ck.ncell + c] = 0;
    a += cnt1[c];
  }
}

__global__ void __launch_bounds__(256) k_scatter_shipped(const Chunk ck) {
  const int pair = blockIdx.y;
  const PairDesc d = ck.desc[pair];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n1) return;
  const size_t o = (size_t)pair * ck.n1max + i;
  const int cell = ck.cellid1[o] & ~CELL_INBOX;
  const int slot = atomicAdd(&ck.cursor[(size_t)pair * ck.ncell + cell], 1);
  const unsigned long long key = ((unsigned long long)(unsigned)ck.pos1[o] << 32) | (unsigned)__float_as_uint(ck.r1[o]);
  ck.kbuf[(size_t)pair * ck.n1max + ck.off[(size_t)pair * ck.ncell + cell] + slot] = key;
}

struct KeyRanges {  // the ranges of a cell in row order: low words of the keys sorted by position
  const unsigned long long* k;
  __device__ __forceinline__ float operator[](int i) const { return __uint_as_float((unsigned)(k[i] & 0xffffffffull)); }
};

__global__ void __launch_bounds__(128) k_cluster_shipped(const Chunk ck) {
  const int pair = blockIdx.y;
  const int nw = ck.nwork[pair];
  for (int w = blockIdx.x; w < nw; w += gridDim.x) {
    const int cell = ck.work[(size_t)pair * ck.ncell + w];
    const int m = ck.cnt1[(size_t)pair * ck.ncell + cell];
    unsigned long long* g = ck.kbuf + (size_t)pair * ck.n1max + ck.off[(size_t)pair * ck.ncell + cell];
    __syncthreads();
    block_sort_asc(g, m);  // by row position (the high word is unique)
    float inner = 0.f, outer = 0.f;
    if (threadIdx.x < 32) find_cluster_warp(KeyRanges{g}, m, 0, ck.n, ck.thresh, ck.buff, inner, outer);
    if (threadIdx.x == 0) write_cluster_rec(ck, pair, cell, m, inner, outer);
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// K3 / K5a: one pass over the points of a scan: [transform,] spherical, cell, cluster-box test,
// sph->cart round trip, fixed-point accumulation of count / sum / sum of products per voxel.
//   SCAN2 = false: scan 1 (filterPointsInsideCluster + mean/cov of fitCells1, src/icet.cpp:155-162)
//   SCAN2 = true : scan 2, one Gauss-Newton iteration (src/icet.cpp:375-388 + fitCells2 :290-306)
// ----------------------------------------------------------------------------------------------
constexpr int PASS_THREADS = 256;
constexpr int PASS_WARPS = PASS_THREADS / 32;
constexpr int PASS_K = 12;       // rows of 32 points per warp tile (throughput shape): 48 KB of entry tiles per CTA,
constexpr int PASS_MINB = 4;     // so that four CTAs (32 warps, 64 registers per thread) fit an SM (r01g: +3 % over 16 / 3)
constexpr int PASS_K_SMALL = 4;  // same for small batches (latency shape: more, smaller tiles; 2 rows: the same
                                 // single-pair latency within 1 %, with twice the tasks)

__host__ __device__ constexpr int pass_wslots(int K) { return 32 * K; }  // 16-byte entry slots per warp tile
__host__ __device__ constexpr int pass_tile_points(int K) { return PASS_WARPS * 32 * K; }
// shared memory: entry tiles, then the angular tables: (nT + 2) + (nP + 2) records {T[k], T[k+1], lo[k], hi[k]}
__host__ __device__ inline int pass_tab_floats(int nT, int nP) { return 4 * (nT + nP + 4); }
__host__ __device__ inline int pass_smem_bytes(int nT, int nP, int K) {
  return PASS_WARPS * pass_wslots(K) * 16 + pass_tab_floats(nT, nP) * 4;
}

// Angular bin + box test in one table look-up.  rec[k] = {T[k], T[k+1], lo[k], hi[k]}:
//   T    exact thresholds of int((double(a)/period)*nb) (src/icet.cpp:545-546): bin k  <=>  T[k] <= a < T[k+1]
//   lo/hi the part of the bin that also passes the reference's inclusive fp32 box test against the bin edges
//        (src/icet.cpp:136-139, :632-633): lo = max(T[k], E[k]), hi = min(pred(T[k+1]), E[k+1]).
// Record nb (a == fp32(period), bin index nb % nb = 0) and record nb + 1 (everything beyond the period, i.e. the
// NaN sentinel 1000.0, whose bin `sbin` comes from the double formula on the host) have an empty [lo, hi].
// The fp32 estimate of k is off by at most one (checked against T).
__device__ __forceinline__ int bin_box(float a, const float4* rec, const icet::BinTable& bt, bool& inbox) {
  int k = __float2int_rz(fminf(a, bt.acap) * bt.scale);
  float4 e = rec[k];
  if (a < e.x || a >= e.y) {
    k += (a < e.x) ? -1 : 1;
    e = rec[k];
  }
  inbox = a >= e.z && a <= e.w;
  return k < bt.nb ? k : (k == bt.nb ? 0 : bt.sbin);
}

// Stage 1 of a point: [transform,] spherical coordinates, cell, gates.  active = the cell takes part
// (has a cluster / an active voxel); in = the point passes ICET::filterPointsInsideCluster (src/icet.cpp:632-634).
template <bool SCAN2>
__device__ __forceinline__ void point_stage1(const Chunk& ck, const float4* tth, const float4* tph, const CellRec* recs,
                                             const float* tr, float x, float y, float z, int& c, bool& active,
                                             bool& in, float& r, float& th, float& ph) {
  if (SCAN2) icet::transform(x, y, z, tr, tr + 3, x, y, z);
  icet::c2s(x, y, z, r, th, ph);
  bool bt_in, bp_in;
  const int bt = bin_box(th, tth, ck.bth, bt_in);
  const int bp = bin_box(ph, tph, ck.bph, bp_in);
  c = ck.nT * bp + bt;
  const float4 ra = __ldg(reinterpret_cast<const float4*>(recs + c));  // inner, outer, flags, scale
  active = (__float_as_uint(ra.z) & (SCAN2 ? F_ACTIVE2 : F_STAT1)) != 0;
  in = active && bt_in && bp_in && r >= ra.x && r <= ra.y;
}

// Stage 2 of an inside point: sph->cart round trip (statistics use round-tripped points, src/icet.cpp:159 / :303)
// and conversion to the voxel's fixed-point frame.
__device__ __forceinline__ void point_stage2(float r, float th, float ph, float refx, float refy, float refz, float sc,
                                             int& fx, int& fy, int& fz) {
  float cx, cy, cz;
  icet::s2c(r, th, ph, cx, cy, cz);
  fx = max(-FP_LIM, min(FP_LIM, __float2int_rn((cx - refx) * sc)));
  fy = max(-FP_LIM, min(FP_LIM, __float2int_rn((cy - refy) * sc)));
  fz = max(-FP_LIM, min(FP_LIM, __float2int_rn((cz - refz) * sc)));
}

// The dropped returns of scan 2: points2_OG == (0,0,0) for all of them, so they all land on t * R
// (src/icet.cpp:377-378) -- evaluated once per iteration, weighted with their number.
__device__ inline void pass_dropped_returns(const Chunk& ck, const float4* tth, const float4* tph, const CellRec* recs,
                                            const float* tr, unsigned long long* accp, long long nz) {
  if (nz <= 0) return;
  int c;
  bool active, in;
  float r, th, ph;
  point_stage1<true>(ck, tth, tph, recs, tr, 0.0f, 0.0f, 0.0f, c, active, in, r, th, ph);
  if (!active) return;
  unsigned long long* q = accp + (size_t)c * NQ;
  red_add(q, (unsigned long long)nz);
  if (in) {
    const CellRec rc = recs[c];
    int ix, iy, iz;
    point_stage2(r, th, ph, rc.refx, rc.refy, rc.refz, rc.scale, ix, iy, iz);
    const long long fx = ix, fy = iy, fz = iz;
    red_add(q + 1, (unsigned long long)nz);
    red_add(q + 2, (unsigned long long)(nz * fx));
    red_add(q + 3, (unsigned long long)(nz * fy));
    red_add(q + 4, (unsigned long long)(nz * fz));
    red_add(q + 5, (unsigned long long)(nz * fx * fx));
    red_add(q + 6, (unsigned long long)(nz * fx * fy));
    red_add(q + 7, (unsigned long long)(nz * fx * fz));
    red_add(q + 8, (unsigned long long)(nz * fy * fy));
    red_add(q + 9, (unsigned long long)(nz * fy * fz));
    red_add(q + 10, (unsigned long long)(nz * fz * fz));
  }
}

// publishes the sums one lane collected over a run of consecutive inside points of one cell
__device__ __forceinline__ void flush_in_run(unsigned long long* accp, int cell, int nin, int sx, int sy, int sz,
                                             long long pxx, long long pxy, long long pxz, long long pyy, long long pyz,
                                             long long pzz) {
  if (cell < 0) return;
  unsigned long long* q = accp + (size_t)cell * NQ;
  red_add(q + 1, (unsigned long long)nin);
  red_add(q + 2, (unsigned long long)(long long)sx);
  red_add(q + 3, (unsigned long long)(long long)sy);
  red_add(q + 4, (unsigned long long)(long long)sz);
  red_add(q + 5, (unsigned long long)pxx);
  red_add(q + 6, (unsigned long long)pxy);
  red_add(q + 7, (unsigned long long)pxz);
  red_add(q + 8, (unsigned long long)pyy);
  red_add(q + 9, (unsigned long long)pyz);
  red_add(q + 10, (unsigned long long)pzz);
}

// One warp tile of 32*K consecutive points.
//  Phase A (lane per point, K rows, coalesced loads): stage 1 of every point.  Points of one cell that are
//    neighbours in the row form runs (LiDAR scans list points ring by ring: ~27 per run at 2048 azimuth steps /
//    75 bins); the head lane of a run adds the run length to the cell's bin count (one RED per run).  Inside points
//    are COMPACTED into the warp's shared-memory tile as {cell, r, theta, phi} -- typically 40 % of the points.
//  Phase B (lane per contiguous slice of the compacted list): stage 2 (the expensive round trip) only for inside
//    points and with all lanes busy; sums of a run of equal cell stay in registers, one flush per run.
// Only warp-level synchronisation inside.
//  G > 1 (latency shape): stage 1 of G rows is evaluated back to back before the warp-synchronous bookkeeping of
//    those rows, so that G independent dependency chains (each with a look-up of the cell record in the middle)
//    overlap inside one warp; the tile's latency, not its instruction count, is what a single pair waits for.
template <bool SCAN2, int K, int PF = 2, int G = 1>
__device__ __forceinline__ void pass_warp_tile(const Chunk& ck, int4* went /* the warp's pass_wslots(K) slots */,
                                               const float* tab, const CellRec* recs, const float* tr,
                                               const float* px_, size_t ld, int n, int w0,
                                               unsigned long long* accp, unsigned long long* dbg_stamp = nullptr,
                                               const int32_t* s1_cell = nullptr, const float* s1_th = nullptr,
                                               const float* s1_ph = nullptr) {
  const int lane = threadIdx.x & 31;
  if (w0 >= n) return;
  const float4* tth = reinterpret_cast<const float4*>(tab);
  const float4* tph = tth + ck.nT + 2;
  const unsigned lt = (1u << lane) - 1u;
  int nin_tile = 0;
  if (!SCAN2) {
    // ---- phase A, scan 1: K1 already stored cell (+ box flag), r, theta, phi of every point; only the range test
    // against the cluster bounds (known since K2c) is left.  px_ = r1 of the pair; th / ph are read for inside points.
    // Rows in groups of R: every load of the group is requested before the first use, so that the only dependent
    // look-up left (the cell record) overlaps across the rows of the group.
    constexpr int R = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);
#pragma unroll 1
    for (int j0 = 0; j0 < K; j0 += R) {
      int cid[R];
      float rr[R], tt[R], pp[R];
#pragma unroll
      for (int g = 0; g < R; g++) {
        const int i = w0 + (j0 + g) * 32 + lane;
        cid[g] = -1; rr[g] = 0.f; tt[g] = 0.f; pp[g] = 0.f;
        if (i < n) { cid[g] = __ldg(s1_cell + i); rr[g] = __ldg(px_ + i); tt[g] = __ldg(s1_th + i); pp[g] = __ldg(s1_ph + i); }
      }
      float4 ra[R];
#pragma unroll
      for (int g = 0; g < R; g++) {
        ra[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cid[g] >= 0) ra[g] = __ldg(reinterpret_cast<const float4*>(recs + (cid[g] & ~CELL_INBOX)));  // inner, outer, flags, scale
      }
#pragma unroll
      for (int g = 0; g < R; g++) {
        const bool in = cid[g] >= 0 && (cid[g] & CELL_INBOX) && (__float_as_uint(ra[g].z) & F_STAT1) && rr[g] >= ra[g].x &&
                        rr[g] <= ra[g].y;
        const unsigned im = __ballot_sync(FULL, in);
        if (in)
          went[nin_tile + __popc(im & lt)] = make_int4(cid[g] & ~CELL_INBOX, __float_as_int(rr[g]), __float_as_int(tt[g]), __float_as_int(pp[g]));
        nin_tile += __popc(im);
      }
    }
  } else if (G > 1) {
    static_assert(G == 1 || K % G == 0, "rows per group must divide the tile");
    // ---- phase A, grouped: all coordinates of the group in flight, then G stage-1 chains, then the bookkeeping
#pragma unroll 1
    for (int j0 = 0; j0 < K; j0 += G) {
      float gx[G], gy[G], gz[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int i = w0 + (j0 + g) * 32 + lane;
        gx[g] = gy[g] = gz[g] = 0.f;
        if (i < n) { gx[g] = __ldg(px_ + i); gy[g] = __ldg(px_ + ld + i); gz[g] = __ldg(px_ + 2 * ld + i); }
      }
      int gc[G];
      bool gact[G], gin[G];
      float gr[G], gth[G], gph[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int i = w0 + (j0 + g) * 32 + lane;
        gc[g] = -1; gact[g] = false; gin[g] = false; gr[g] = 0.f; gth[g] = 0.f; gph[g] = 0.f;
        if (i < n) point_stage1<SCAN2>(ck, tth, tph, recs, tr, gx[g], gy[g], gz[g], gc[g], gact[g], gin[g], gr[g], gth[g], gph[g]);
      }
#pragma unroll
      for (int g = 0; g < G; g++) {
        const int key = gact[g] ? gc[g] : -1;
        const int prev = __shfl_up_sync(FULL, key, 1);
        const bool head = (lane == 0) || (key != prev);
        const unsigned hm = __ballot_sync(FULL, head);
        if (head && key >= 0) {
          const unsigned nh = (lane == 31) ? 0u : (hm >> (lane + 1));
          const int len = nh ? __ffs(nh) : 32 - lane;
          red_add(accp + (size_t)key * NQ, (unsigned long long)len);
        }
        const unsigned im = __ballot_sync(FULL, gin[g]);
        if (gin[g])
          went[nin_tile + __popc(im & lt)] = make_int4(gc[g], __float_as_int(gr[g]), __float_as_int(gth[g]), __float_as_int(gph[g]));
        nin_tile += __popc(im);
      }
    }
  } else {
  // ---- phase A (the coordinates of row j + PF are requested before row j is worked on)
  float bx[PF > 0 ? PF : 1], by[PF > 0 ? PF : 1], bz[PF > 0 ? PF : 1];
#pragma unroll
  for (int p = 0; p < PF; p++) {
    const int i = w0 + p * 32 + lane;
    bx[p] = by[p] = bz[p] = 0.f;
    if (p < K && i < n) { bx[p] = __ldg(px_ + i); by[p] = __ldg(px_ + ld + i); bz[p] = __ldg(px_ + 2 * ld + i); }
  }
#pragma unroll(PF > 0 ? (K % (2 * PF) == 0 ? 2 * PF : PF) : 2)
  for (int j = 0; j < K; j++) {
    const int i = w0 + j * 32 + lane;
    int c = -1;
    bool active = false, in = false;
    float r = 0.f, th = 0.f, ph = 0.f;
    float x = 0.f, y = 0.f, z = 0.f;
    if (PF > 0) {
      constexpr int PFm = PF > 0 ? PF : 1;
      x = bx[j % PFm]; y = by[j % PFm]; z = bz[j % PFm];
      const int ip = i + PF * 32;
      if (j + PF < K && ip < n) {
        bx[j % PFm] = __ldg(px_ + ip); by[j % PFm] = __ldg(px_ + ld + ip); bz[j % PFm] = __ldg(px_ + 2 * ld + ip);
      }
    } else if (i < n) {
      x = __ldg(px_ + i); y = __ldg(px_ + ld + i); z = __ldg(px_ + 2 * ld + i);
    }
    if (i < n) point_stage1<SCAN2>(ck, tth, tph, recs, tr, x, y, z, c, active, in, r, th, ph);
    // bin counts: one RED per run of equal (participating) cell in this row
    const int key = active ? c : -1;
    const int prev = __shfl_up_sync(FULL, key, 1);
    const bool head = (lane == 0) || (key != prev);
    const unsigned hm = __ballot_sync(FULL, head);
    if (head && key >= 0) {
      const unsigned nh = (lane == 31) ? 0u : (hm >> (lane + 1));
      const int len = nh ? __ffs(nh) : 32 - lane;
      red_add(accp + (size_t)key * NQ, (unsigned long long)len);
    }
    // compaction of the inside points
    const unsigned im = __ballot_sync(FULL, in);
    if (in) went[nin_tile + __popc(im & lt)] = make_int4(c, __float_as_int(r), __float_as_int(th), __float_as_int(ph));
    nin_tile += __popc(im);
  }
  }  // G == 1
  __syncwarp();
  if (dbg_stamp && lane == 0) {  // debug timeline: end of phase A
    unsigned long long t_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    *dbg_stamp = t_;
  }
  // ---- phase B: lane takes entries [lane*q, lane*q + q); q odd => conflict-free 16-byte shared loads
  const int q = ((nin_tile + 31) >> 5) | 1;
  const int e0 = lane * q, e1 = min(nin_tile, e0 + q);
  int cur = -1, nin = 0, sx = 0, sy = 0, sz = 0;
  long long pxx = 0, pxy = 0, pxz = 0, pyy = 0, pyz = 0, pzz = 0;
  float refx = 0.f, refy = 0.f, refz = 0.f, sc = 0.f;
#pragma unroll 2
  for (int e = e0; e < e1; e++) {
    const int4 v = went[e];
    if (v.x != cur) {
      flush_in_run(accp, cur, nin, sx, sy, sz, pxx, pxy, pxz, pyy, pyz, pzz);
      cur = v.x;
      nin = sx = sy = sz = 0;
      pxx = pxy = pxz = pyy = pyz = pzz = 0;
      const float4* rp = reinterpret_cast<const float4*>(recs + cur);
      const float4 ra = __ldg(rp), rb = __ldg(rp + 1);
      sc = ra.w; refx = rb.x; refy = rb.y; refz = rb.z;
    }
    int fx, fy, fz;
    point_stage2(__int_as_float(v.y), __int_as_float(v.z), __int_as_float(v.w), refx, refy, refz, sc, fx, fy, fz);
    nin++;
    sx += fx; sy += fy; sz += fz;
    pxx += (long long)fx * fx; pxy += (long long)fx * fy; pxz += (long long)fx * fz;
    pyy += (long long)fy * fy; pyz += (long long)fy * fz; pzz += (long long)fz * fz;
  }
  flush_in_run(accp, cur, nin, sx, sy, sz, pxx, pxy, pxz, pyy, pyz, pzz);
  __syncwarp();
}

template <bool SCAN2, int K = PASS_K, int MINB = PASS_MINB, int PF = 2, int G = 1>
__global__ void __launch_bounds__(PASS_THREADS, MINB) k_pass(const Chunk ck) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int4* ent = reinterpret_cast<int4*>(smem_raw);
  float* tab = reinterpret_cast<float*>(smem_raw + PASS_WARPS * pass_wslots(K) * 16);
  const int pair = blockIdx.y;
  const PairDesc d = ck.desc[pair];
  const int n = SCAN2 ? ck.n2c[pair] : d.n1;
  const int tile0 = blockIdx.x * pass_tile_points(K);
  if (tile0 >= n && !(SCAN2 && blockIdx.x == 0)) return;
  {
    const int ntab = pass_tab_floats(ck.nT, ck.nP);
    for (int k = threadIdx.x; k < ntab; k += PASS_THREADS) tab[k] = __ldg(ck.binrec + k);
  }
  float tr[12];
  if (SCAN2) {
    const float4* tp = reinterpret_cast<const float4*>(ck.TR + (size_t)pair * 12);
    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    tr[0] = a.x; tr[1] = a.y; tr[2] = a.z; tr[3] = a.w; tr[4] = b.x; tr[5] = b.y; tr[6] = b.z; tr[7] = b.w;
    tr[8] = c.x; tr[9] = c.y; tr[10] = c.z; tr[11] = c.w;
  }
  const size_t o1 = (size_t)pair * ck.n1max;
  const float* px_ = SCAN2 ? ck.pog + (size_t)pair * 3 * ck.n2max : ck.r1 + o1;
  const size_t ld = SCAN2 ? (size_t)ck.n2max : (size_t)d.ld1;
  const CellRec* recs = ck.rec + (size_t)pair * ck.ncell;
  unsigned long long* accp = ck.acc + (size_t)pair * ck.ncell * NQ;
  __syncthreads();
  pass_warp_tile<SCAN2, K, PF, G>(ck, ent + (threadIdx.x >> 5) * pass_wslots(K), tab, recs, tr, px_, ld, n,
                           tile0 + (threadIdx.x >> 5) * 32 * K, accp, nullptr, SCAN2 ? nullptr : ck.cellid1 + o1,
                           SCAN2 ? nullptr : ck.th1 + o1, SCAN2 ? nullptr : ck.ph1 + o1);
  if (SCAN2 && blockIdx.x == 0 && threadIdx.x == 0)
    pass_dropped_returns(ck, reinterpret_cast<const float4*>(tab), reinterpret_cast<const float4*>(tab) + ck.nT + 2, recs, tr,
                         accp, ck.nz2[pair]);
}

// exact-sum -> mean / covariance (double) of a voxel
__device__ __forceinline__ void stats_from_acc(const unsigned long long* q, const CellRec& rc, double mean[3],
                                               double cov[6]) {
  const double nin = (double)(long long)q[1];
  const double inv = 1.0 / (double)rc.scale;  // exact: the scale is a power of two
  const double in_ = 1.0 / nin;
  const double sx = (double)(long long)q[2], sy = (double)(long long)q[3], sz = (double)(long long)q[4];
  const double mx = sx * in_, my = sy * in_, mz = sz * in_;
  mean[0] = (double)rc.refx + mx * inv;
  mean[1] = (double)rc.refy + my * inv;
  mean[2] = (double)rc.refz + mz * inv;
  const double f = inv * inv / (nin - 1.0);
  cov[0] = ((double)(long long)q[5] - sx * mx) * f;
  cov[1] = ((double)(long long)q[6] - sx * my) * f;
  cov[2] = ((double)(long long)q[7] - sx * mz) * f;
  cov[3] = ((double)(long long)q[8] - sy * my) * f;
  cov[4] = ((double)(long long)q[9] - sy * mz) * f;
  cov[5] = ((double)(long long)q[10] - sz * mz) * f;
}

// ----------------------------------------------------------------------------------------------
// K4: per voxel of scan 1: mean / covariance, 3x3 eigen-decomposition, sigma points, L mask
// (fitCells1 src/icet.cpp:158-232, testSigmaPoints :654-696); constants for the iteration loop.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fit1(const Chunk ck) {
  const int pair = blockIdx.y;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ck.ncell) return;
  const size_t ci = (size_t)pair * ck.ncell + cell;
  CellRec rc = ck.rec[ci];
  unsigned long long* q = ck.acc + ci * NQ;
  bool has = false;
  if (rc.flags & F_STAT1) {
    const long long nin = (long long)q[1];
    if (ck.dump_on) ck.dump.nin1[cell] = (int)nin;
    // `filteredPoints.size() >= n` with size() = 3 * rows  (src/icet.cpp:158)
    if (3 * nin >= ck.n && nin >= 2) {
      has = true;
      double mean[3], cov[6];
      stats_from_acc(q, rc, mean, cov);
      Vox1 v;
      for (int k = 0; k < 3; k++) v.mu[k] = mean[k];
      const double d1 = (double)(rc.cnt1 - 1);  // `indices1.size() - 1` (:315)
      for (int k = 0; k < 6; k++) v.S1n[k] = cov[k] / d1;
      float A[9] = {(float)cov[0], (float)cov[1], (float)cov[2], (float)cov[1], (float)cov[3],
                    (float)cov[4], (float)cov[2], (float)cov[4], (float)cov[5]};
      float ev[3], V[9];
      icet::eig3f(A, ev, V);
      // sigma points mu +- 2 sqrt(ev_k) * V.row(k)   (:187-202), tested in order 0+,0-,1+,1-,2+,2-
      const float mu[3] = {(float)mean[0], (float)mean[1], (float)mean[2]};
      const int bt = cell % ck.nT, bp = cell / ck.nT;
      const float azl = ck.azE[bt], azh = ck.azE[bt + 1], ell = ck.elE[bp], elh = ck.elE[bp + 1];
      bool inside[6] = {false, false, false, false, false, false};
      for (int j = 0; j < 6; j++) {
        const int k = j >> 1;
        const float al = 2.0f * sqrtf(ev[k]);
        float p[3];
        for (int c = 0; c < 3; c++) {
          float rot = al * V[3 * k + c];
          p[c] = (j & 1) ? mu[c] - rot : mu[c] + rot;
        }
        float r, th, ph;
        icet::c2s(p[0], p[1], p[2], r, th, ph);
        if (th >= azl && th <= azh && ph >= ell && ph <= elh && r >= rc.inner && r <= rc.outer) inside[j] = true;
        if (r > rc.outer) break;  // the early break of testSigmaPoints (:683-685)
      }
      int lm = 0;
      for (int k = 0; k < 3; k++)
        if (inside[2 * k] || inside[2 * k + 1]) lm |= (1 << k);
      v.lmask = lm;
      v.pad = 0;
      for (int k = 0; k < 3; k++)
        for (int c = 0; c < 3; c++) v.LV[3 * k + c] = (lm >> k & 1) ? (double)V[3 * k + c] : 0.0;
      ck.vox[ci] = v;
      if (ck.dump_on) {
        for (int k = 0; k < 3; k++) { ck.dump.mu1[3 * cell + k] = (float)mean[k]; ck.dump.eval1[3 * cell + k] = ev[k]; }
        for (int k = 0; k < 9; k++) { ck.dump.sigma1[9 * cell + k] = A[k]; ck.dump.evec1[9 * cell + k] = V[k]; }
        for (int k = 0; k < 3; k++) ck.dump.lmask[3 * cell + k] = (lm >> k) & 1;
      }
    }
  }
  if (ck.dump_on) ck.dump.has1[cell] = has ? 1 : 0;
  // gates of fitCells2 that do not depend on scan 2: `indices1.size() > n && bounds[5] > 1`
  // (src/icet.cpp:290); a voxel without a scan-1 Gaussian is skipped.
  uint32_t fl = rc.flags & ~F_ACTIVE2;
  if (has && rc.cnt1 > ck.n && rc.outer > 1.0f) fl |= F_ACTIVE2;
  if (fl != rc.flags) ck.rec[ci].flags = fl;
  if (rc.flags & F_STAT1)
    for (int k = 0; k < NQ; k++) q[k] = 0ull;  // hand the accumulators to the scan-2 loop
  if (has) atomicAdd(&ck.res[pair].n_gauss1, 1);
}

// ----------------------------------------------------------------------------------------------
// prepScan2 (src/icet.cpp:254-277): points2_OG = sphericalToCartesian(cartesianToSpherical(scan2)).
// (The radial re-ordering of scan 2 only changes the reference's summation order.)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep2(const Chunk ck) {
  const int pair = blockIdx.y;
  const PairDesc d = ck.desc[pair];
  if ((int)(blockIdx.x * blockDim.x) >= d.n2) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0.f, y = 0.f, z = 0.f;
  bool keep = false, zero = false;
  // Consecutive pairs of a sequence share a scan: scan 2 of this pair is scan 1 of the next one, whose spherical
  // coordinates K1 has already stored (same function, same inputs) -- read them instead of converting again.
  bool shared = false;
  if (pair + 1 < ck.npairs) {
    const PairDesc nx = ck.desc[pair + 1];
    shared = nx.s1 == d.s2 && nx.n1 == d.n2 && nx.ld1 == d.ld2;
  }
  if (i < d.n2) {
    float r, th, ph;
    if (shared) {
      const size_t o = (size_t)(pair + 1) * ck.n1max + i;
      r = __ldg(ck.r1 + o); th = __ldg(ck.th1 + o); ph = __ldg(ck.ph1 + o);
    } else {
      x = __ldg(d.s2 + i); y = __ldg(d.s2 + d.ld2 + i); z = __ldg(d.s2 + 2 * (size_t)d.ld2 + i);
      icet::c2s(x, y, z, r, th, ph);
    }
    icet::s2c(r, th, ph, x, y, z);
    // Dropped returns: (0,0,0) stays (+0,+0,+0).  They are all the same point in every iteration, so they are
    // counted here and evaluated once per iteration by k_pass<true> instead of being stored.
    zero = (__float_as_uint(x) | __float_as_uint(y) | __float_as_uint(z)) == 0u;
    keep = !zero;
  }
  // block-level compaction (the order of points2_OG is irrelevant: all sums over it are exact integers)
  __shared__ int s_cnt[8], s_zero[8], s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned mk = __ballot_sync(FULL, keep), mz = __ballot_sync(FULL, zero);
  if (lane == 0) { s_cnt[warp] = __popc(mk); s_zero[warp] = __popc(mz); }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0, totz = 0;
    for (int w = 0; w < 8; w++) { const int c = s_cnt[w]; s_cnt[w] = tot; tot += c; totz += s_zero[w]; }
    s_base = tot ? atomicAdd(&ck.n2c[pair], tot) : 0;
    if (totz) atomicAdd(&ck.nz2[pair], totz);
  }
  __syncthreads();
  if (keep) {
    const int o = s_base + s_cnt[warp] + __popc(mk & ((1u << lane) - 1));
    float* pg = ck.pog + (size_t)pair * 3 * ck.n2max;
    pg[o] = x;
    pg[ck.n2max + o] = y;
    pg[2 * (size_t)ck.n2max + o] = z;
  }
}

// ----------------------------------------------------------------------------------------------
// K5b: per voxel.  Scan-2 mean / covariance, R_noise, W, H_z and the voxel's contributions
// H^T W H_j (upper triangle, 21) and H^T W dz_j (6)  (fitCells2 src/icet.cpp:302-338), ADDED to acc[].
// Takes the voxel's integer accumulators (L2 reads: other SMs produced them) and clears them for the next
// iteration.
// ----------------------------------------------------------------------------------------------
constexpr int VOX_THREADS = 64;
constexpr int NRED = 28;  // 21 (upper triangle of H^T W H) + 6 (H^T W dz) + 1 (voxels used)

// Sum over the 32 lanes of each of the NRED (28) per-lane values; lane k < NRED returns the total of value k.
// Transposing butterfly: at distance o the lanes with bit o set keep the upper half of the remaining values and the
// others the lower half, so 16 + 8 + 4 + 2 + 1 shuffles replace 28 x 5.  The pairing of the additions is that of the
// plain xor butterfly (lanes L and L ^ o at every level), so the sums are bit-identical to it.
__device__ __forceinline__ double warp_sum_transposed(const double (&acc)[NRED], int lane) {
  double v[16];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const double hi = (i + 16 < NRED) ? acc[i + 16] : 0.0;
      const double send = up ? acc[i] : hi;
      const double keep = up ? hi : acc[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, 16);
    }
  }
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const double send = up ? v[i] : v[i + o];
      const double keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, o);
    }
  }
  return v[0];
}

// Gate of fitCells2 for one voxel (`indices2.size() > n` :290, `rows > n` :302).  Returns true if the voxel
// contributes; its accumulators are then left in place for vox_algebra, otherwise they are cleared here.
__device__ __forceinline__ bool vox_gate(const Chunk& ck, int pair, int cell, int iter) {
  const size_t ci = (size_t)pair * ck.ncell + cell;
  const uint32_t flags = ck.rec[ci].flags;
  if (!(flags & F_ACTIVE2)) {
    if (ck.dump_on) {
      ck.dump.cnt2[(size_t)iter * ck.ncell + cell] = -1;
      ck.dump.nin2[(size_t)iter * ck.ncell + cell] = -1;
      ck.dump.used2[(size_t)iter * ck.ncell + cell] = 0;
    }
    return false;
  }
  unsigned long long* qp = ck.acc + ci * NQ;
  const ulonglong2 q01 = __ldcg(reinterpret_cast<const ulonglong2*>(qp));
  const long long nbin = (long long)q01.x, nin = (long long)q01.y;
  const bool use = nbin > ck.n && nin > ck.n;
  if (ck.dump_on) {
    ck.dump.cnt2[(size_t)iter * ck.ncell + cell] = (int)nbin;
    ck.dump.nin2[(size_t)iter * ck.ncell + cell] = (nbin > ck.n) ? (int)nin : -1;
    ck.dump.used2[(size_t)iter * ck.ncell + cell] = use ? 1 : 0;
  }
  if (!use && nbin != 0) {
#pragma unroll
    for (int k = 0; k < NQ; k++) qp[k] = 0ull;
  }
  return use;
}

// The algebra of one contributing voxel; takes (and clears) its accumulators.
__device__ __forceinline__ void vox_algebra(const Chunk& ck, int pair, int cell, int iter, const float* Jm /* 27 */,
                                            double acc[NRED]) {
  const size_t ci = (size_t)pair * ck.ncell + cell;
  const CellRec rc = ck.rec[ci];
  unsigned long long q[NQ];
  unsigned long long* qp = ck.acc + ci * NQ;
#pragma unroll
  for (int k = 0; k < NQ; k += 2) {
    const ulonglong2 t = __ldcg(reinterpret_cast<const ulonglong2*>(qp + k));
    q[k] = t.x;
    q[k + 1] = t.y;
  }
#pragma unroll
  for (int k = 0; k < NQ; k++) qp[k] = 0ull;
  const long long nbin = (long long)q[0];
  double mean[3], cov[6];
  stats_from_acc(q, rc, mean, cov);
  const Vox1 v = ck.vox[ci];
  if (ck.dump_on) {
    float* m2 = ck.dump.mu2 + ((size_t)iter * ck.ncell + cell) * 3;
    float* s2 = ck.dump.sigma2 + ((size_t)iter * ck.ncell + cell) * 9;
    for (int k = 0; k < 3; k++) m2[k] = (float)mean[k];
    s2[0] = (float)cov[0]; s2[1] = (float)cov[1]; s2[2] = (float)cov[2];
    s2[3] = (float)cov[1]; s2[4] = (float)cov[3]; s2[5] = (float)cov[4];
    s2[6] = (float)cov[2]; s2[7] = (float)cov[4]; s2[8] = (float)cov[5];
  }
  // R_noise = sigma1/(|idx1|-1) + sigma2/(|idx2|-1)   (:315)
  const double id2 = 1.0 / (double)(nbin - 1);
  double Rn[9];
  {
    double r6[6];
#pragma unroll
    for (int k = 0; k < 6; k++) r6[k] = v.S1n[k] + cov[k] * id2;
    Rn[0] = r6[0]; Rn[1] = r6[1]; Rn[2] = r6[2]; Rn[3] = r6[1]; Rn[4] = r6[3]; Rn[5] = r6[4];
    Rn[6] = r6[2]; Rn[7] = r6[4]; Rn[8] = r6[5];
  }
  // M = (L U^T) R_noise (L U^T)^T   (:317)
  double T[9], M[9];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++)
      T[3 * a + b] = v.LV[3 * a] * Rn[b] + v.LV[3 * a + 1] * Rn[3 + b] + v.LV[3 * a + 2] * Rn[6 + b];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++)
      M[3 * a + b] = T[3 * a] * v.LV[3 * b] + T[3 * a + 1] * v.LV[3 * b + 1] + T[3 * a + 2] * v.LV[3 * b + 2];
  // W = pinv(M)  (:320-321)
  double W[9];
  if (!icet::masked_inv3(M, v.lmask, W)) icet::cod_pinv(M, 3, 3, W);
  // H_z = L U^T [ -I | Jx mu | Jy mu | Jz mu ]  (:324-329)
  double H[18];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    H[6 * a + 0] = (a == 0) ? -1.0 : 0.0;
    H[6 * a + 1] = (a == 1) ? -1.0 : 0.0;
    H[6 * a + 2] = (a == 2) ? -1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < 3; j++)
      H[6 * a + 3 + j] = (double)Jm[9 * j + 3 * a] * mean[0] + (double)Jm[9 * j + 3 * a + 1] * mean[1] +
                         (double)Jm[9 * j + 3 * a + 2] * mean[2];
  }
  double Hz[18];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int c = 0; c < 6; c++)
      Hz[6 * a + c] = v.LV[3 * a] * H[c] + v.LV[3 * a + 1] * H[6 + c] + v.LV[3 * a + 2] * H[12 + c];
  // dz = L U^T (mean2 - mu1)   (:335-337)
  const double dm[3] = {mean[0] - v.mu[0], mean[1] - v.mu[1], mean[2] - v.mu[2]};
  double dz[3];
#pragma unroll
  for (int a = 0; a < 3; a++) dz[a] = v.LV[3 * a] * dm[0] + v.LV[3 * a + 1] * dm[1] + v.LV[3 * a + 2] * dm[2];
  double WH[18], Wdz[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int c = 0; c < 6; c++)
      WH[6 * a + c] = W[3 * a] * Hz[c] + W[3 * a + 1] * Hz[6 + c] + W[3 * a + 2] * Hz[12 + c];
    Wdz[a] = W[3 * a] * dz[0] + W[3 * a + 1] * dz[1] + W[3 * a + 2] * dz[2];
  }
  int t = 0;
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int b = a; b < 6; b++) acc[t++] += Hz[a] * WH[b] + Hz[6 + a] * WH[6 + b] + Hz[12 + a] * WH[12 + b];
#pragma unroll
  for (int a = 0; a < 6; a++) acc[21 + a] += Hz[a] * Wdz[0] + Hz[6 + a] * Wdz[1] + Hz[12 + a] * Wdz[2];
  acc[27] += 1.0;
}

__device__ __forceinline__ void vox_contrib(const Chunk& ck, int pair, int cell, int iter, const float* Jm,
                                            double acc[NRED]) {
  if (vox_gate(ck, pair, cell, iter)) vox_algebra(ck, pair, cell, iter, Jm, acc);
}

// ----------------------------------------------------------------------------------------------
// K6: one thread per pair: Q = pinv(H^T W H), pred_stds, checkCondition, dx, X += dx (src/icet.cpp:410-433,
// :443-492) and the transform / get_H trigonometry of the next iteration.  tot = the 28 sums over the voxels.
// ----------------------------------------------------------------------------------------------
// Odometry chaining (odometry.cpp:82 `X0 << X[0], ...`): the registration of pair k+1 starts from the solution of
// pair k.  Called by the thread that has just written X / TR / J of `pair` in the LAST iteration, before that
// iteration is published.
__device__ __forceinline__ void chain_seed_next(const Chunk& ck, int pair, const float* Xn) {
  if (!(ck.flags & ICET_B200_FLAG_CHAIN_X0) || pair + 1 >= ck.npairs) return;
  const int q = pair + 1;
  for (int k = 0; k < 6; k++) { ck.X[q * 6 + k] = Xn[k]; ck.res[q].X[k] = Xn[k]; }
  for (int k = 0; k < 12; k++) {
    const float v = ck.TR[(size_t)pair * 12 + k];
    ck.TR[(size_t)q * 12 + k] = v;
    ck.TRprev[(size_t)q * 12 + k] = v;
  }
  for (int k = 0; k < 27; k++) ck.J[(size_t)q * 27 + k] = ck.J[(size_t)pair * 27 + k];
}

__device__ __noinline__ void solve_pair(const Chunk& ck, int pair, int iter, const double* tot) {
  float* X = ck.X + pair * 6;
  double A[36], b[6];
  {
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int c = a; c < 6; c++) {
        A[a * 6 + c] = tot[t];
        A[c * 6 + a] = tot[t];
        t++;
      }
#pragma unroll
    for (int a = 0; a < 6; a++) b[a] = tot[21 + a];
  }
  icet_b200_result* R = ck.res + pair;
  double Q[36], dx[6], stds[6];
  int dropped = 0, status = 0;
  double cond_out = 0.0;
  bool fast = false;
  if (!(ck.flags & ICET_B200_FLAG_FULL_EIG) && icet::chol_inv6(A, Q)) {
    double trA = 0.0, trQ = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) { trA += A[k * 6 + k]; trQ += Q[k * 6 + k]; }
    // cond <= trace(A) * trace(A^-1); comfortably below the 1e6 cutoff => no axis is dropped and
    // pinv == inverse, so dx = A^-1 b  (src/icet.cpp:410-433 with an empty while-loop at :469)
    if (trA * trQ < 0.999e6) {
      fast = true;
      cond_out = -(trA * trQ);
    }
  }
  if (fast) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 6; j++) s += Q[k * 6 + j] * b[j];
      dx[k] = s;
      stds[k] = sqrt(fabs(Q[k * 6 + k]));
    }
  } else {
    double ev[6], U[36];
    icet::jacobi6(A, ev, U);
    // noise_mat = pinv(H^T W H) (:410-411) from the eigen-decomposition, with the rank rule of the COD
    // (pivot > FLT_EPSILON * 6 * largest pivot) applied to the spectrum
    {
      double lm = 0.0;
      for (int k = 0; k < 6; k++) lm = fmax(lm, fabs(ev[k]));
      const double thr = (double)FLT_EPSILON * 6.0 * lm;
      for (int i = 0; i < 36; i++) Q[i] = 0.0;
      for (int k = 0; k < 6; k++) {
        if (!(fabs(ev[k]) > thr)) continue;
        const double il = 1.0 / ev[k];
        for (int i = 0; i < 6; i++)
          for (int j = 0; j < 6; j++) Q[i * 6 + j] += U[i * 6 + k] * U[j * 6 + k] * il;
      }
    }
    for (int k = 0; k < 6; k++) stds[k] = sqrt(fabs(Q[k * 6 + k]));
    const double cutoff = 1e6;
    double condition = ev[5] / ev[0];
    cond_out = condition;
    int eyecount = 1;
    while (fabs(condition) > cutoff) {  // checkCondition :469-486
      if (eyecount > 5) { status = ICET_B200_COND_OVERFLOW; break; }
      for (int k = 0; k < 6; k++) stds[k] += U[k * 6 + (eyecount - 1)];  // :479
      dropped++;
      condition = ev[5] / ev[eyecount];
      eyecount++;
    }
    // dx = pinv(L2 lam U2^T) L2 U2^T b = sum over kept k of u_k (u_k . b) / lam_k, with the rank
    // rule of the COD applied to the kept spectrum (:427-430)
    double lmax = 0.0;
    for (int k = dropped; k < 6; k++) lmax = fmax(lmax, fabs(ev[k]));
    const double tiny = (double)FLT_EPSILON * (double)(6 - dropped) * lmax;
    for (int k = 0; k < 6; k++) dx[k] = 0.0;
    for (int k = dropped; k < 6; k++) {
      if (!(fabs(ev[k]) > tiny)) continue;
      double ub = 0.0;
      for (int j = 0; j < 6; j++) ub += U[j * 6 + k] * b[j];
      for (int j = 0; j < 6; j++) dx[j] += U[j * 6 + k] * (ub / ev[k]);
    }
  }
  float Xn[6];
  for (int k = 0; k < 6; k++) Xn[k] = (float)((double)__ldcg(X + k) + dx[k]);  // X += dx (:433), X is fp32
  for (int k = 0; k < 6; k++) X[k] = Xn[k];
  {  // trigonometry of the next iteration: utils::R (src/icet.cpp:375-376) and get_H (:507-527)
    float* TR = ck.TR + (size_t)pair * 12;
    if (iter == ck.runlen - 1)
      for (int k = 0; k < 12; k++) ck.TRprev[(size_t)pair * 12 + k] = __ldcg(TR + k);
    TR[0] = Xn[0]; TR[1] = Xn[1]; TR[2] = Xn[2];
    icet::rotR(Xn[3], Xn[4], Xn[5], TR + 3);
    icet::getH_J(Xn[3], Xn[4], Xn[5], ck.J + (size_t)pair * 27);
    if (iter == ck.runlen - 1) chain_seed_next(ck, pair, Xn);
  }
  if (ck.dump_on) {
    for (int k = 0; k < 6; k++) { ck.dump.Xit[iter * 6 + k] = Xn[k]; ck.dump.HTWdz[iter * 6 + k] = (float)b[k]; }
    for (int k = 0; k < 36; k++) ck.dump.HTWH[iter * 36 + k] = (float)A[k];
  }
  if (iter == ck.runlen - 1) {
    for (int k = 0; k < 6; k++) { R->X[k] = Xn[k]; R->pred_stds[k] = (float)stds[k]; }
    for (int k = 0; k < 36; k++) R->Q[k] = (float)Q[k];
    R->n_used = (int)(tot[27] + 0.5);
    R->n_dropped = dropped;
    R->cond = (float)cond_out;
  }
  if (status) R->status = status;
}

// Warp-parallel form of the common case of solve_pair: Gauss-Jordan elimination of [A | I | b] (13 columns, one
// per lane, no pivoting: A = H^T W H is symmetric positive definite whenever this path is valid).  Returns (warp
// uniform) false when a pivot is not positive or the bound trace(A) trace(A^-1) cannot prove cond <= 1e6; the caller
// then runs solve_pair (eigen-decomposition + the reference's truncation loop) on one thread.
__device__ __forceinline__ bool solve_pair_warp(const Chunk& ck, int pair, int iter, const double* tot) {
  const int lane = threadIdx.x & 31;
  // requested now, used after the elimination: the current X and (last iteration) the transform it started from
  const float x_old = lane < 6 ? __ldcg(ck.X + pair * 6 + lane) : 0.f;
  const float tr_old = (lane < 12 && iter == ck.runlen - 1) ? __ldcg(ck.TR + (size_t)pair * 12 + lane) : 0.f;
  double col[6];
  {
    // lane j < 6: column j of A; lane 6 + j: column j of I; lane 12: b
    const int j = lane < 6 ? lane : 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      const double a = tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];  // packed upper triangle
      col[i] = lane < 6 ? a : (lane < 12 ? (lane - 6 == i ? 1.0 : 0.0) : (lane == 12 ? tot[21 + i] : 0.0));
    }
  }
  double trA = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++) trA += tot[i * 6 - i * (i - 1) / 2];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double piv = __shfl_sync(FULL, col[k], k);
    ok = ok && (piv > 0.0);
    const double ip = 1.0 / piv;
    col[k] *= ip;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      if (i == k) continue;
      const double f = __shfl_sync(FULL, col[i], k);
      col[i] -= f * col[k];
    }
  }
  double trQ = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) trQ += __shfl_sync(FULL, col[k], 6 + k);
  // cond <= trace(A) * trace(A^-1); comfortably below the 1e6 cutoff => checkCondition drops nothing and
  // pinv == inverse, so dx = A^-1 b  (src/icet.cpp:410-433 with an empty while-loop at :469)
  if (!(ok && trA * trQ < 0.999e6)) return false;
  icet_b200_result* R = ck.res + pair;
  const bool last = iter == ck.runlen - 1;
  // ---- the update, warp-wide (one lane doing all of it serially was the longest stretch of the iteration)
  // X += dx (:433), X is fp32: lane k < 6 owns component k (dx_k lives in lane 12's col[k])
  double dxk = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const double v = __shfl_sync(FULL, col[k], 12);
    dxk = (lane == k) ? v : dxk;
  }
  const float xn = (float)((double)x_old + dxk);
  // sines / cosines of the three angles on lanes 0..2, then every lane forms all of R(X) and the get_H matrices
  // (~100 flops, same expressions as utils::R / get_H) and stores the entries it owns
  float sv, cv;
  sincosf(__shfl_sync(FULL, xn, 3 + (lane % 3)), &sv, &cv);
  const float sph = __shfl_sync(FULL, sv, 0), cph = __shfl_sync(FULL, cv, 0);
  const float sth = __shfl_sync(FULL, sv, 1), cth = __shfl_sync(FULL, cv, 1);
  const float sps = __shfl_sync(FULL, sv, 2), cps = __shfl_sync(FULL, cv, 2);
  float Rm[9], Jm[27];
  icet::rotR_sc(sph, cph, sth, cth, sps, cps, Rm);
  icet::getH_J_sc(sph, cph, sth, cth, sps, cps, Jm);
  const float t_k = __shfl_sync(FULL, xn, lane < 3 ? lane : 0);
  float trv = t_k, jv = 0.f;
#pragma unroll
  for (int e = 0; e < 9; e++) trv = (lane == 3 + e) ? Rm[e] : trv;
#pragma unroll
  for (int e = 0; e < 27; e++) jv = (lane == e) ? Jm[e] : jv;
  const bool seed = last && (ck.flags & ICET_B200_FLAG_CHAIN_X0) && pair + 1 < ck.npairs;  // odometry.cpp:82
  if (lane < 6) {
    ck.X[pair * 6 + lane] = xn;
    if (last) R->X[lane] = xn;
    if (seed) { ck.X[(pair + 1) * 6 + lane] = xn; ck.res[pair + 1].X[lane] = xn; }
    if (ck.dump_on) { ck.dump.Xit[iter * 6 + lane] = xn; ck.dump.HTWdz[iter * 6 + lane] = (float)tot[21 + lane]; }
  }
  if (lane < 12) {
    if (last) ck.TRprev[(size_t)pair * 12 + lane] = tr_old;  // the transform the LAST iteration used (`points2`)
    ck.TR[(size_t)pair * 12 + lane] = trv;
    if (seed) { ck.TR[(size_t)(pair + 1) * 12 + lane] = trv; ck.TRprev[(size_t)(pair + 1) * 12 + lane] = trv; }
  }
  if (lane < 27) {
    ck.J[(size_t)pair * 27 + lane] = jv;
    if (seed) ck.J[(size_t)(pair + 1) * 27 + lane] = jv;
  }
  if (last && lane == 12) {
    R->n_used = (int)(tot[27] + 0.5);
    R->n_dropped = 0;
    R->cond = (float)(-(trA * trQ));
  }
  if (ck.dump_on && lane < 6) {
    for (int i = 0; i < 6; i++) {
      const int lo = i < lane ? i : lane, hi = i < lane ? lane : i;
      ck.dump.HTWH[iter * 36 + i * 6 + lane] = (float)tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    }
  }
  if (last && lane >= 6 && lane < 12) {
    const int j = lane - 6;
#pragma unroll
    for (int i = 0; i < 6; i++) R->Q[i * 6 + j] = (float)col[i];
    // pred_stds = sqrt|diag noise_mat| (:414-417)
    double d = col[0];
#pragma unroll
    for (int i = 1; i < 6; i++) d = (j == i) ? col[i] : d;
    R->pred_stds[j] = (float)sqrt(fabs(d));
  }
  return true;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot)                                                                         \
  do {                                                                                   \
    if (ck.dump_on && (threadIdx.x & 31) == 0) ck.dump.tl[(size_t)iter * 16 + (slot)] = gtime(); \
  } while (0)

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *p >= need.  A protocol failure must not hang the GPU: after ~2 s the wait gives up, records what it
// was waiting for in ck.dbg and marks the pair (results of the chunk are then invalid; the host reports an error).
// The polling itself uses RELAXED loads: an acquire load is followed by an invalidation of the SM's whole L1
// (CCTL.IVALL), and thousands of spinning warps would keep every L1 of the GPU empty for the warps that do the work.
// One acquire load after the condition has been seen orders the reads that follow.
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __noinline__ void loop_wait(const Chunk& ck, const int* p, int need, int kind, int pair, int iter, unsigned ticket) {
  if (ld_acquire(p) >= need) return;
  const unsigned long long t0 = gtime();
  unsigned spins = 0;
  for (;;) {
    __nanosleep(40);
    const int seen = ld_relaxed(p);
    if (seen >= need) {
      ld_acquire(p);
      return;
    }
    if ((++spins & 1023u) == 0 && gtime() - t0 > 2000000000ull) {
      if ((threadIdx.x & 31) == 0 && atomicCAS(ck.dbg, 0, 1) == 0) {
        ck.dbg[1] = kind; ck.dbg[2] = pair; ck.dbg[3] = iter; ck.dbg[4] = seen; ck.dbg[5] = need; ck.dbg[6] = (int)ticket;
        ck.res[pair].status = ICET_B200_LOOP_TIMEOUT;
      }
      return;
    }
  }
}

__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// warp tiles of a pair that are counted in tiles_done per iteration: those that hold points, at least tile 0
__device__ __forceinline__ int loop_tiles_of(int n2c, int tile_points) { return max(1, (n2c + tile_points - 1) / tile_points); }

// One vox task of k_loop (see there): the fitCells2 algebra of 32 consecutive cells, and, for the task that
// arrives last, the end of the iteration of the pair.
__device__ __noinline__ void vox_task(const Chunk& ck, int iter, int pair, int grp, int tpt, int vt, double* w_tot,
                                      float* w_J) {
  const int lane = threadIdx.x & 31;
  const int cell = grp * 32 + lane;
  const bool active = cell < ck.ncell && (ck.rec[(size_t)pair * ck.ncell + cell].flags & F_ACTIVE2) != 0;
  const bool any = __any_sync(FULL, active);
  if (any) {
    loop_wait(ck, reinterpret_cast<const int*>(ck.tiles_done + pair), (iter + 1) * loop_tiles_of(__ldg(ck.n2c + pair), tpt),
              1, pair, iter, 0u);
    if (lane < 27) w_J[lane] = __ldcg(ck.J + (size_t)pair * 27 + lane);
    __syncwarp();
    TL(1);
    double acc[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) acc[k] = 0.0;
    if (cell < ck.ncell && vox_gate(ck, pair, cell, iter)) vox_algebra(ck, pair, cell, iter, w_J, acc);
    TL(8);
    if (__any_sync(FULL, acc[27] != 0.0)) {
      double* out = ck.part + ((size_t)pair * vt + grp) * NRED;
      const double tot = warp_sum_transposed(acc, lane);
      if (lane < NRED) out[lane] = tot;
      TL(9);
      // (the fence below, executed by every lane before vox_done is bumped, also covers these stores and the mask bit)
      if (lane == 0) {
        red_or(ck.vmask + (size_t)pair * ((vt + 31) / 32) + (grp >> 5), 1u << (grp & 31));
      }
    }
    TL(2);
  } else if (ck.dump_on && cell < ck.ncell) {
    vox_gate(ck, pair, cell, iter);  // records the "inactive" markers
  }
  __threadfence();  // every lane: partial sums, cleared accumulators, the group's vmask bit
  __syncwarp();
  unsigned prev = 0;
  if (lane == 0) {
    prev = atom_add(ck.vox_done + (size_t)pair * ck.runlen + iter, 1u);
  }
  prev = __shfl_sync(FULL, prev, 0);
  if (prev + 1u == (unsigned)vt) {
    // -------------------------------------------------------------- end of the iteration of this pair
    // Iterations of a pair are strictly ordered: groups without an active voxel do not wait for the tiles, so when
    // NO group of the pair has one (degenerate inputs) nothing else would keep iteration k+1 from being closed
    // before iteration k.
    if (iter > 0) loop_wait(ck, ck.iter_done + pair, iter, 2, pair, iter, 0u);
    else if ((ck.flags & ICET_B200_FLAG_CHAIN_X0) && pair > 0)
      loop_wait(ck, ck.iter_done + pair - 1, ck.runlen, 3, pair, iter, 0u);  // X of this pair comes from pair - 1
    __threadfence();
    TL(0);
    double tot = 0.0;
    const int nw = (vt + 31) / 32;
    for (int w = 0; w < nw; w++) {
      unsigned* mp = ck.vmask + (size_t)pair * nw + w;
      unsigned m = __ldcg(mp);
      if (lane == 0 && m) *mp = 0u;
      const double* pp = ck.part + ((size_t)pair * vt + (size_t)w * 32) * NRED + (lane < NRED ? lane : 0);
      while (m) {  // group order; four loads in flight
        int g[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          g[u] = m ? __ffs(m) - 1 : -1;
          m = m ? (m & (m - 1)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = (g[u] >= 0) ? __ldcg(pp + (size_t)g[u] * NRED) : 0.0;
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (g[u] >= 0) tot += v[u];
      }
    }
    __syncwarp();
    if (lane < NRED) w_tot[lane] = tot;
    __syncwarp();
    TL(3);
    bool done = false;
    if (!(ck.flags & ICET_B200_FLAG_FULL_EIG)) done = solve_pair_warp(ck, pair, iter, w_tot);
    if (!done && lane == 0) solve_pair(ck, pair, iter, w_tot);
    __threadfence();  // every lane: X, TR, J, result fields
    __syncwarp();
    TL(4);
    if (lane == 0) st_release(ck.iter_done + pair, iter + 1);
    TL(5);
  }
}

// ----------------------------------------------------------------------------------------------
// The Gauss-Newton loop of every pair of the chunk in ONE persistent launch (fitScan2 x runlen,
// src/icet.cpp:47, :372-436): nothing returns to the host between iterations.
//
// The worker is the WARP.  Work is a stream of tasks, ordered
//     for it in 0..runlen:  [ (it, pair, tile) for every pair, tile ]  then  [ (it, pair, vox group) for every pair, group ]
// and warps draw tickets from one counter.
//   * tile task: one warp tile of scan 2 through pass_warp_tile (transform ... integer accumulation), then
//     tiles_done[pair] += 1.  Needs X of iteration it: waits until iter_done[pair] >= it.
//   * vox task: the fitCells2 algebra of 32 consecutive cells (one per lane), a fixed-order warp reduction of the
//     28 sums into part[pair][group][28].  Needs all tiles of (pair, it): waits on tiles_done[pair].  The vox task
//     that arrives last (vox_done[pair][it]) adds the partials in group order, solves the 6x6 system on the warp
//     (solve_pair_warp), writes the next transform and publishes iter_done[pair] = it + 1.
// Every task only ever waits for tasks with SMALLER tickets, which are held by warps that are already running, so
// the scheme cannot deadlock whatever the number of resident warps.  With many pairs in flight nobody waits (the
// solve of one pair overlaps the tiles of the others); with one pair the waits ARE the latency-critical path and
// the vox groups / tiles of the pair spread over the whole GPU.
// ----------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(PASS_THREADS, 3) k_loop(const Chunk ck, int tiles, int vt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tab = reinterpret_cast<float*>(smem_raw + PASS_WARPS * pass_wslots(K) * 16);
  {
    const int ntab = pass_tab_floats(ck.nT, ck.nP);
    for (int k = threadIdx.x; k < ntab; k += PASS_THREADS) tab[k] = __ldg(ck.binrec + k);
  }
  __syncthreads();  // the only block-wide barrier: from here on warps are independent workers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int4* went = reinterpret_cast<int4*>(smem_raw) + warp * pass_wslots(K);
  static_assert(pass_wslots(K) * 16 >= 512, "warp scratch too small");
  double* w_tot = reinterpret_cast<double*>(went);                          // [28]   (vox tasks only)
  float* w_J = reinterpret_cast<float*>(reinterpret_cast<char*>(went) + 256);  // [27]
  const unsigned ntile = (unsigned)ck.npairs * (unsigned)tiles;
  const unsigned per_iter = ntile + (unsigned)ck.npairs * (unsigned)vt;
  const unsigned total = per_iter * (unsigned)ck.runlen;
  const bool chain = (ck.flags & ICET_B200_FLAG_CHAIN_X0) != 0;
  unsigned t = 0;
  if (lane == 0) t = atom_add(ck.ticket, 1u);
  t = __shfl_sync(FULL, t, 0);
  while (t < total) {
    // Throughput shape: the next ticket is drawn now, its round trip to L2 hides behind this task.  Latency shape
    // (small tiles, more resident warps than tasks per iteration): one ticket per warp -- a warp that holds two
    // tiles of the same iteration starts the second one a whole tile late, while idle warps could have taken it.
    unsigned t_next = 0;
    if (K > PASS_K_SMALL && lane == 0) t_next = atom_add(ck.ticket, 1u);
    // ticket -> task.  Independent pairs: iteration-major (see above).  Chained pairs (ICET_B200_FLAG_CHAIN_X0): pair-
    // major, i.e. all iterations of pair k before any task of pair k + 1, whose first tiles wait for the last solve
    // of pair k (it seeds X / TR / J of pair k + 1) -- still only waits on smaller tickets.
    int iter, pair_t, sub;
    bool is_tile;
    if (chain) {
      const unsigned per_it1 = (unsigned)(tiles + vt);
      const unsigned per_pair = per_it1 * (unsigned)ck.runlen;
      pair_t = (int)(t / per_pair);
      const unsigned r1 = t - (unsigned)pair_t * per_pair;
      iter = (int)(r1 / per_it1);
      const unsigned r2 = r1 - (unsigned)iter * per_it1;
      is_tile = r2 < (unsigned)tiles;
      sub = is_tile ? (int)r2 : (int)(r2 - (unsigned)tiles);
    } else {
      iter = (int)(t / per_iter);
      const unsigned rem = t - (unsigned)iter * per_iter;
      is_tile = rem < ntile;
      if (is_tile) {
        pair_t = (int)(rem / (unsigned)tiles);
        sub = (int)(rem - (unsigned)pair_t * (unsigned)tiles);
      } else {
        pair_t = (int)((rem - ntile) / (unsigned)vt);
        sub = (int)((rem - ntile) - (unsigned)pair_t * (unsigned)vt);
      }
    }
    if (is_tile) {
      // ------------------------------------------------------------------ tile task
      const int pair = pair_t;
      const int tile = sub;
      const int n = __ldg(ck.n2c + pair);
      const int w0 = tile * 32 * K;
      if (w0 < n || tile == 0) {
        if (iter > 0) loop_wait(ck, ck.iter_done + pair, iter, 0, pair, iter, t);
        else if (chain && pair > 0) loop_wait(ck, ck.iter_done + pair - 1, ck.runlen, 4, pair, iter, t);
        float tr[12];
        {
          const float4* tp = reinterpret_cast<const float4*>(ck.TR + (size_t)pair * 12);
          const float4 a = __ldcg(tp), b = __ldcg(tp + 1), c = __ldcg(tp + 2);
          tr[0] = a.x; tr[1] = a.y; tr[2] = a.z; tr[3] = a.w; tr[4] = b.x; tr[5] = b.y; tr[6] = b.z; tr[7] = b.w;
          tr[8] = c.x; tr[9] = c.y; tr[10] = c.z; tr[11] = c.w;
        }
        const CellRec* recs = ck.rec + (size_t)pair * ck.ncell;
        unsigned long long* accp = ck.acc + (size_t)pair * ck.ncell * NQ;
        if (tile == 0) TL(6);
        if (ck.dump_on && iter == 3 && tile < 2048 && lane == 0) ck.dump.tl[(size_t)ck.runlen * 16 + 2 * tile] = gtime();
        pass_warp_tile<true, K, 2, (K <= 4 ? K : 1)>(ck, went, tab, recs, tr, ck.pog + (size_t)pair * 3 * ck.n2max, (size_t)ck.n2max, n, w0,
                                accp, (ck.dump_on && iter == 3 && tile < 2048) ? ck.dump.tl + (size_t)ck.runlen * 16 + 4096 + tile : nullptr);
        if (tile == 0 && lane == 0)
          pass_dropped_returns(ck, reinterpret_cast<const float4*>(tab), reinterpret_cast<const float4*>(tab) + ck.nT + 2,
                               recs, tr, accp, __ldg(ck.nz2 + pair));
        if (tile == 0) TL(7);
        if (ck.dump_on && iter == 3 && tile < 2048 && lane == 0) ck.dump.tl[(size_t)ck.runlen * 16 + 2 * tile + 1] = gtime();
        __threadfence();  // every lane: its accumulator updates are visible before the tile is counted
        __syncwarp();
        if (lane == 0) red_add(ck.tiles_done + pair, 1u);
      }  // tiles beyond the compacted point count of the pair are not counted (see loop_tiles_of)
    } else {
      // ------------------------------------------------------------------ vox task
      vox_task(ck, iter, pair_t, sub, 32 * K, vt, w_tot, w_J);
    }
    if (K <= PASS_K_SMALL && lane == 0) t_next = atom_add(ck.ticket, 1u);
    t = __shfl_sync(FULL, t_next, 0);
  }
}

// Legacy split form of the loop (ICET_B200_FLAG_UNFUSED_LOOP): k_pass<true>, then these two, per iteration.
__global__ void __launch_bounds__(VOX_THREADS) k_vox2(const Chunk ck, int iter) {
  const int pair = blockIdx.y;
  const int cell = blockIdx.x * VOX_THREADS + threadIdx.x;
  __shared__ double s_red[NRED];
  double acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; k++) acc[k] = 0.0;
  if (cell < ck.ncell) vox_contrib(ck, pair, cell, iter, ck.J + (size_t)pair * 27, acc);
  // fixed-order reduction: xor butterfly inside each warp, then warp 1 + warp 0
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool any = __syncthreads_or(acc[27] != 0.0);
  double* out = ck.part + ((size_t)pair * gridDim.x + blockIdx.x) * NRED;
  if (!any) {  // block-uniform: nothing to add
    if (threadIdx.x < NRED) out[threadIdx.x] = 0.0;
    return;
  }
  const double tot = warp_sum_transposed(acc, lane);
  if (wid == 1 && lane < NRED) s_red[lane] = tot;
  __syncthreads();
  if (wid == 0 && lane < NRED) out[lane] = tot + s_red[lane];
}

__global__ void __launch_bounds__(32) k_solve6(const Chunk ck, int iter, int nblk) {
  const int pair = blockIdx.x;
  const int lane = threadIdx.x;
  __shared__ double s_tot[NRED];
  if (lane < NRED) {
    double mine = 0.0;
    const double* pp = ck.part + (size_t)pair * nblk * NRED + lane;
    for (int b = 0; b < nblk; b++) mine += pp[(size_t)b * NRED];
    s_tot[lane] = mine;
  }
  __syncwarp();
  // (one thread: this kernel is bound by instruction fetch of once-executed code, the warp-parallel solve of
  // k_loop is not faster here)
  if (lane == 0) solve_pair(ck, pair, iter, s_tot);
}

// public member `points2` of the reference: scan 2 as transformed by the last iteration
// ((points2_OG + t) * R with the X that iteration STARTED from, src/icet.cpp:375-378), pair 0 of the chunk
__global__ void k_points2(const Chunk ck, int n2, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const PairDesc d = ck.desc[0];
  float x = d.s2[i], y = d.s2[d.ld2 + i], z = d.s2[2 * (size_t)d.ld2 + i];
  float r, th, ph;
  icet::c2s(x, y, z, r, th, ph);   // points2_OG (prepScan2, src/icet.cpp:263-275), recomputed: the workspace
  icet::s2c(r, th, ph, x, y, z);   // copy is compacted
  icet::transform(x, y, z, ck.TRprev, ck.TRprev + 3, x, y, z);
  out[i] = x;
  out[n2 + i] = y;
  out[2 * (size_t)n2 + i] = z;
}

// spherical coordinates + cell index of a cloud (parity-test entry point)
__global__ void k_sph_bins(const float* s, int n, int ld, int nT, int nP, icet::BinTable bth, icet::BinTable bph,
                           float* sph, int32_t* cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r, th, ph;
  icet::c2s(s[i], s[ld + i], s[2 * (size_t)ld + i], r, th, ph);
  int bt, bp;
  bt = icet::bin_lookup(th, bth, 2 * M_PI);
  bp = icet::bin_lookup(ph, bph, M_PI);
  sph[i] = r; sph[n + i] = th; sph[2 * (size_t)n + i] = ph;
  cell[i] = nT * bp + bt;
}

// synthetic scans
__global__ void k_synth(uint64_t seed, int first_scan, int nscans, int rings, int azim, const synth::Pose* poses,
                        float* out) {
  const int npts = rings * azim;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)nscans * npts) return;
  const int s = (int)(gid / npts), i = (int)(gid % npts);
  const int ring = i / azim, az = i % azim;
  float x, y, z;
  synth::ray(seed, first_scan + s, poses[s], ring, rings, az, azim, x, y, z);
  float* o = out + (size_t)s * 3 * npts;
  o[i] = x; o[npts + i] = y; o[2 * (size_t)npts + i] = z;
}

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
constexpr int ICET_NSLOT = 4;  // staging slots of the host-buffer pipeline
constexpr int ICET_NLANE = 4;  // compute lanes: consecutive chunks rotate over up to four streams (each with its own
                               // workspace) so that the latency-bound ends of one chunk's kernels overlap the other's

struct icet_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t lanes[ICET_NLANE] = {};  // compute lanes 1.. (lane 0 is `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join[ICET_NLANE] = {};
  cudaEvent_t ev_aux[2] = {};  // single-pair chunks: prepScan2 runs beside the scan-1 kernels on lane 1
  int nlanes_default = 4;
  int nlanes = 4;
  cudaEvent_t ev_copy[ICET_NSLOT] = {};
  cudaEvent_t ev_done[ICET_NSLOT] = {};
  int chunk_pairs = 256;
  int host_chunk = 64;  // pairs per chunk of the host-buffer pipeline (upload of chunk k+1 || registration of chunk k)
  int64_t launches = 0;
  int dump_on = 0;
  int sm_count = 148;
  int pass_smem_set = 0;  // dynamic shared memory the pass kernels are currently allowed
  int* loop_dbg[ICET_NLANE] = {};  // watchdog record of the last k_loop launch per lane
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
  char last_ck[640];
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
size_t carve_chunk(void* base, int P, int ncell, int n1max, int n2max, int runlen, Chunk& ck, size_t* zero_bytes,
                   bool shipped = false) {
  const int vt = (ncell + 31) / 32;
  Carve c(base);
  ck.cnt1 = c.take<int32_t>((size_t)P * ncell);
  ck.cntz = c.take<int32_t>((size_t)P * ncell);
  ck.cursor = c.take<int32_t>((size_t)P * ncell);
  ck.acc = c.take<unsigned long long>((size_t)P * ncell * NQ);
  ck.n2c = c.take<int32_t>((size_t)P);
  ck.nz2 = c.take<int32_t>((size_t)P);
  ck.ticket = c.take<unsigned>(1);
  ck.tiles_done = c.take<unsigned>((size_t)P);
  ck.iter_done = c.take<int>((size_t)P);
  ck.vox_done = c.take<unsigned>((size_t)P * std::max(1, runlen));
  ck.vmask = c.take<unsigned>((size_t)P * ((vt + 31) / 32));
  ck.dbg = c.take<int>(8);
  c.off = (c.off + 255) & ~(size_t)255;
  if (zero_bytes) *zero_bytes = c.off;
  ck.off = c.take<int32_t>((size_t)P * ncell);
  ck.work = c.take<int32_t>((size_t)P * ncell);
  ck.nwork = c.take<int32_t>((size_t)P);
  ck.nbig = c.take<int32_t>((size_t)P);
  ck.rec = c.take<CellRec>((size_t)P * ncell);
  ck.vox = c.take<Vox1>((size_t)P * ncell);
  ck.cellid1 = c.take<int32_t>((size_t)P * n1max);
  ck.r1 = c.take<float>((size_t)P * n1max);
  ck.th1 = c.take<float>((size_t)P * n1max);
  ck.ph1 = c.take<float>((size_t)P * n1max);
  ck.rbuf = c.take<float>((size_t)P * n1max);
  ck.kbuf = shipped ? c.take<unsigned long long>((size_t)P * n1max) : nullptr;
  ck.pos1 = shipped ? c.take<int32_t>((size_t)P * n1max) : nullptr;
  ck.pog = c.take<float>((size_t)P * 3 * n2max);
  ck.X = c.take<float>((size_t)P * 6);
  ck.TR = c.take<float>((size_t)P * 12);
  ck.TRprev = c.take<float>((size_t)P * 12);
  ck.J = c.take<float>((size_t)P * 27);
  ck.part = c.take<double>((size_t)P * vt * 28);  // per vox group (k_loop) / per 64-voxel block (split loop)
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
  std::vector<float> e(((nbase + 3) & ~(size_t)3) + (size_t)4 * (nT + nP + 4));
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
  const int nT = p->bins_theta, nP = p->bins_phi, ncell = nT * nP;
  int rc = ensure_edges(ctx, nT, nP);
  if (rc) return rc;
  Chunk ck;
  memset(&ck, 0, sizeof(ck));
  size_t zero_bytes = 0;
  const bool shipped = (p->flags & ICET_B200_FLAG_SHIPPED_ORDER) != 0;
  if (shipped && P != 1)
    return fail(ICET_B200_E_INVALID, "ICET_B200_FLAG_SHIPPED_ORDER is a single-pair validation mode");
  size_t need = carve_chunk(nullptr, P, ncell, n1max, n2max, p->runlen, ck, &zero_bytes, shipped);
  rc = ctx->ws[lane].ensure(need);
  if (rc) return rc;
  carve_chunk(ctx->ws[lane].p, P, ncell, n1max, n2max, p->runlen, ck, &zero_bytes, shipped);
  ck.desc = d_desc;
  ck.npairs = P; ck.ncell = ncell; ck.nT = nT; ck.nP = nP; ck.n = p->n; ck.runlen = p->runlen;
  ck.flags = p->flags; ck.thresh = p->thresh; ck.buff = p->buff;
  ck.n1max = n1max; ck.n2max = n2max;
  fill_tables(ctx, nT, nP, &ck.azE, &ck.elE, &ck.bth, &ck.bph, &ck.binrec);
  ck.x0 = d_x0;
  ck.res = d_res;
  ck.dump_on = dump ? 1 : 0;
  if (dump) ck.dump = ctx->dump_ptrs;
  cudaStream_t st = lane == 0 ? ctx->stream : ctx->lanes[lane];
  CK(cudaMemsetAsync(ctx->ws[lane].p, 0, zero_bytes, st));
  const dim3 g1((n1max + 255) / 256, P), g2((n2max + 255) / 256, P);
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
  // Latency shape (one pair): prepScan2 does not depend on the scan-1 kernels, so it runs beside them on lane 1.
  const bool prep_aside = P == 1 && n2max > 0 && lane == 0 && !ctx->profile_on && ctx->lanes[1] != nullptr;
  if (prep_aside) {
    CK(cudaEventRecord(ctx->ev_aux[0], st));  // after the workspace has been cleared and the descriptor uploaded
    CK(cudaStreamWaitEvent(ctx->lanes[1], ctx->ev_aux[0], 0));
    k_prep2<<<g2, 256, 0, ctx->lanes[1]>>>(ck);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev_aux[1], ctx->lanes[1]));
  }
  if (n1max > 0) LAUNCH(0, k_scan1_bin<<<g1, 256, 0, st>>>(ck));
  LAUNCH(1, k_cell_scan<<<P, 256, 0, st>>>(ck));
  if (n1max > 0) {
    if (shipped) {
      // The row order the reference ends up with (src/icet.cpp:72-83), reproduced on the host from the ranges the
      // device computed: the same index sort by range (std::sort; the reference's std::execution::par falls back
      // to it without TBB) and the same swap loop, which is NOT a valid permutation application.
      std::vector<float> hr((size_t)n1max);
      CK(cudaMemcpyAsync(hr.data(), ck.r1, (size_t)n1max * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<int> index((size_t)n1max), orig((size_t)n1max), pos((size_t)n1max);
      for (int i = 0; i < n1max; i++) index[i] = orig[i] = i;
      std::sort(index.begin(), index.end(), [&](int a, int b) { return hr[a] < hr[b]; });
      for (int i = 0; i < n1max; i++) {
        if (index[i] != i) {
          const int j = index[i];
          std::swap(orig[i], orig[j]);    // points1Spherical.row(i).swap(points1Spherical.row(index[i]))
          std::swap(index[i], index[j]);  // std::swap(index[i], index[index[i]])
        }
      }
      for (int i = 0; i < n1max; i++) pos[orig[i]] = i;
      CK(cudaMemcpyAsync(ck.pos1, pos.data(), (size_t)n1max * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));
      LAUNCH(2, k_off_shipped<<<P, 256, 0, st>>>(ck));
      LAUNCH(2, k_scatter_shipped<<<g1, 256, 0, st>>>(ck));
      LAUNCH(3, k_cluster_shipped<<<dim3(std::max(1, std::min(ncell, 1024)), P), 128, 0, st>>>(ck));
    } else {
    LAUNCH(2, k_scatter<<<g1, 256, 0, st>>>(ck));
    // one warp per listed cell; enough CTAs to cover a typical work list (~25 % of the cells) in one pass
    int gx = std::max(1, std::min((ncell + CLUSTER_WARPS - 1) / CLUSTER_WARPS,
                                  std::max(32, (ctx->sm_count * 16 + P - 1) / P)));
    LAUNCH(3, k_cluster<<<dim3(gx, P), CLUSTER_WARPS * 32, 0, st>>>(ck));
    }
    if (small_batch) {  // latency shape: 128 points per warp
      const int tile_s = pass_tile_points(PASS_K_SMALL);
      LAUNCH(4, k_pass<false, PASS_K_SMALL, 3, 1, PASS_K_SMALL><<<dim3((n1max + tile_s - 1) / tile_s, P), PASS_THREADS, psm2, st>>>(ck));
    } else {
      LAUNCH(4, k_pass<false><<<gp1, PASS_THREADS, psm, st>>>(ck));
    }
  }
  LAUNCH(5, k_fit1<<<dim3((ncell + 127) / 128, P), 128, 0, st>>>(ck));
  if (prep_aside) CK(cudaStreamWaitEvent(st, ctx->ev_aux[1], 0));
  else if (n2max > 0) LAUNCH(6, k_prep2<<<g2, 256, 0, st>>>(ck));
  const bool use_loop = chain || (p->flags & ICET_B200_FLAG_PERSISTENT_LOOP) ||
                        (!(p->flags & ICET_B200_FLAG_UNFUSED_LOOP) && P <= ICET_LOOP_MAX_PAIRS);
  if (!use_loop) {
    for (int it = 0; it < p->runlen; it++) {
      if (n2max > 0) LAUNCH(7, k_pass<true><<<gp2, PASS_THREADS, psm, st>>>(ck));
      LAUNCH(8, k_vox2<<<dim3(nblk, P), VOX_THREADS, 0, st>>>(ck, it));
      LAUNCH(9, k_solve6<<<P, 32, 0, st>>>(ck, it, nblk));
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
  static_assert(sizeof(Chunk) <= 640, "Chunk too large for last_ck");
  ctx->last_valid = (P == 1);
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

#include "callers.cuh"

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int icet_b200_version(void) { return ICET_B200_VERSION; }
const char* icet_b200_last_error(void) { return g_err.c_str(); }

int icet_b200_create(int device, icet_b200_ctx** out) {
  if (!out) return fail(ICET_B200_E_INVALID, "ctx out-pointer is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(ICET_B200_E_NODEVICE, std::string("no CUDA device available (") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                          "); icet_b200 has no CPU fallback");
  }
  if (device < 0) CK(cudaGetDevice(&device));
  if (device >= ndev) return fail(ICET_B200_E_INVALID, "device index out of range");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(ICET_B200_E_NODEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                          std::to_string(prop.minor) + "; this library is built for sm_100a only");
  CK(cudaSetDevice(device));
  icet_b200_ctx* c = new icet_b200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int l = 1; l < ICET_NLANE; l++) {
    CK(cudaStreamCreateWithFlags(&c->lanes[l], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&c->ev_aux[k], cudaEventDisableTiming));
  for (int i = 0; i < ICET_NSLOT; i++) {
    CK(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  }
  *out = c;
  return 0;
}

int icet_b200_destroy(icet_b200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->copy_stream);
  for (int l = 1; l < ICET_NLANE; l++)
    if (c->lanes[l]) cudaStreamSynchronize(c->lanes[l]);
  for (int i = 0; i < ICET_NLANE; i++) c->ws[i].release();
  c->edges.release(); c->resbuf.release(); c->dumpbuf.release(); c->posebuf.release();
  c->rawbuf[0].release(); c->rawbuf[1].release(); c->planebuf.release();
  for (int i = 0; i < ICET_NSLOT; i++) {
    c->stage[i].release(); c->descbuf[i].release(); c->x0buf[i].release();
    if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  }
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (int l = 1; l < ICET_NLANE; l++) {
    if (c->lanes[l]) cudaStreamDestroy(c->lanes[l]);
    if (c->ev_join[l]) cudaEventDestroy(c->ev_join[l]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int k = 0; k < 2; k++)
    if (c->ev_aux[k]) cudaEventDestroy(c->ev_aux[k]);
  delete c;
  return 0;
}

int icet_b200_set_stream(icet_b200_ctx* c, void* stream) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)stream;
  c->own_stream = false;
  return 0;
}

int icet_b200_set_chunk(icet_b200_ctx* c, int32_t m) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  if (m < 0) return fail(ICET_B200_E_INVALID, "chunk must be >= 0");
  c->chunk_pairs = m == 0 ? 256 : std::min(m, 65535);
  return 0;
}

int icet_b200_set_host_chunk(icet_b200_ctx* c, int32_t m) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  if (m < 0) return fail(ICET_B200_E_INVALID, "chunk must be >= 0");
  c->host_chunk = m == 0 ? 64 : std::min(m, 65535);
  return 0;
}

int icet_b200_synchronize(icet_b200_ctx* c) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  for (int l = 1; l < ICET_NLANE; l++) CK(cudaStreamSynchronize(c->lanes[l]));
  return 0;
}

int icet_b200_set_lanes(icet_b200_ctx* c, int32_t n) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  if (n < 0 || n > ICET_NLANE) return fail(ICET_B200_E_INVALID, "lanes must be 0 (default) .. 4");
  c->nlanes = n == 0 ? c->nlanes_default : n;
  return 0;
}

int64_t icet_b200_kernel_launches(icet_b200_ctx* c) { return c ? c->launches : 0; }

static const char* const k_names[ICET_B200_NKERNELS] = {"k_scan1_bin", "k_cell_scan", "k_scatter", "k_cluster",
                                                        "k_pass<scan1>", "k_fit1", "k_prep2", "k_pass<scan2>",
                                                        "k_vox2", "k_solve6", "k_loop"};
const char* icet_b200_kernel_name(int id) { return (id >= 0 && id < ICET_B200_NKERNELS) ? k_names[id] : ""; }

static int prof_collect(icet_b200_ctx* c) {
  CK(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < c->prof_id.size(); i++) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
    c->prof_ms[c->prof_id[i]] += ms;
    c->prof_n[c->prof_id[i]]++;
  }
  c->prof_id.clear();
  c->prof_used = 0;
  return 0;
}

int icet_b200_set_profile(icet_b200_ctx* c, int32_t enable) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  int rc = prof_collect(c);
  if (rc) return rc;
  if (enable && !c->profile_on)
    for (int k = 0; k < ICET_B200_NKERNELS; k++) { c->prof_ms[k] = 0.0; c->prof_n[k] = 0; }
  c->profile_on = enable ? 1 : 0;
  return 0;
}

int icet_b200_get_profile(icet_b200_ctx* c, double* ms, int64_t* launches) {
  if (!c || !ms || !launches) return fail(ICET_B200_E_INVALID, "NULL argument");
  CK(cudaSetDevice(c->device));
  int rc = prof_collect(c);
  if (rc) return rc;
  for (int k = 0; k < ICET_B200_NKERNELS; k++) { ms[k] = c->prof_ms[k]; launches[k] = c->prof_n[k]; }
  return 0;
}

int icet_b200_set_dump(icet_b200_ctx* c, int32_t enable) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  c->dump_on = enable ? 1 : 0;
  if (!enable) c->dump_valid = false;
  return 0;
}

// -- device-resident batch --------------------------------------------------------------------
static int batch_device_impl(icet_b200_ctx* c, const icet_b200_params* p, int32_t npairs, const PairDesc* h_desc,
                             const float* d_x0, icet_b200_result* d_out, bool dump, const PairDesc* d_desc_all = nullptr,
                             int nmax_dev = 0) {
  // h_desc lives in host memory; descriptors are uploaded per chunk through the pinned bounce buffer.
  // d_desc_all (callers layer): the descriptors are already on the device -- built there, with data-dependent
  // sizes no larger than nmax_dev -- and h_desc is not read.
  int rc = ensure_pinned(c, (size_t)std::min(npairs, c->chunk_pairs) * sizeof(PairDesc) * ICET_NSLOT + 4096);
  if (rc) return rc;
  // consecutive chunks alternate between the two compute lanes (own stream + workspace each); lane 1 starts after
  // everything already queued on the caller's stream and the caller's stream resumes after lane 1
  const bool chain = (p->flags & ICET_B200_FLAG_CHAIN_X0) != 0;  // chunks depend on each other: one lane, in order
  // Chunk size: the configured bound, but small enough that every lane gets a chunk (not below 64 pairs): with all
  // lanes busy the latency-bound kernels at the end of each iteration of one chunk hide behind the pass kernels of
  // the others (measured at 512 pairs: 2 x 256 -> 66.9 k pairs/s, 4 x 128 -> 69.1 k).
  const int chunk_pairs = (c->nlanes > 1 && !dump && !chain)
                              ? std::min(c->chunk_pairs, std::max(64, (npairs + c->nlanes - 1) / c->nlanes))
                              : c->chunk_pairs;
  const int nchunks = (npairs + chunk_pairs - 1) / chunk_pairs;
  const int nl = (c->nlanes > 1 && nchunks > 1 && !dump && !chain) ? std::min(c->nlanes, nchunks) : 1;
  if (nl > 1) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (int l = 1; l < nl; l++) CK(cudaStreamWaitEvent(c->lanes[l], c->ev_fork, 0));
  }
  static_assert(ICET_NLANE <= ICET_NSLOT, "one descriptor slot per lane");
  int chunk_no = 0;
  for (int base = 0; base < npairs; base += chunk_pairs, chunk_no++) {
    const int P = std::min(chunk_pairs, npairs - base);
    const int lane = chunk_no % nl;
    const int slot = lane;  // descriptors of a lane are reused in stream order
    cudaStream_t st = lane == 0 ? c->stream : c->lanes[lane];
    int n1max = nmax_dev, n2max = nmax_dev;
    const PairDesc* d_desc = d_desc_all ? d_desc_all + base : nullptr;
    if (!d_desc) {
      for (int i = 0; i < P; i++) {
        n1max = std::max(n1max, h_desc[base + i].n1);
        n2max = std::max(n2max, h_desc[base + i].n2);
      }
      rc = c->descbuf[slot].ensure((size_t)P * sizeof(PairDesc));
      if (rc) return rc;
      // the pinned half `slot` may still be in flight from two chunks ago
      CK(cudaEventSynchronize(c->ev_done[slot]));
      PairDesc* hp = (PairDesc*)c->pinned + (size_t)slot * std::min(npairs, c->chunk_pairs);
      memcpy(hp, h_desc + base, (size_t)P * sizeof(PairDesc));
      CK(cudaMemcpyAsync(c->descbuf[slot].p, hp, (size_t)P * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(c->ev_done[slot], st));
      d_desc = (const PairDesc*)c->descbuf[slot].p;
    }
    const float* x0c = d_x0 ? d_x0 + (chain ? 0 : (size_t)base * 6) : nullptr;
    if (chain && base > 0) x0c = d_out[base - 1].X;  // stream order: the previous chunk has finished by then
    rc = run_chunk(c, p, P, d_desc, n1max, n2max, x0c, d_out + base, dump, lane);
    if (rc) return rc;
  }
  for (int l = 1; l < nl; l++) {
    CK(cudaEventRecord(c->ev_join[l], c->lanes[l]));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join[l], 0));
  }
  return 0;
}

int icet_b200_register_batch_device(icet_b200_ctx* c, const icet_b200_params* p, int32_t npairs,
                                    const float* const* scan1, const int32_t* n1, const float* const* scan2,
                                    const int32_t* n2, const float* x0, icet_b200_result* out) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  int rc = validate(p);
  if (rc) return rc;
  if (npairs < 0 || (npairs > 0 && (!scan1 || !scan2 || !n1 || !n2 || !out)))
    return fail(ICET_B200_E_INVALID, "NULL argument");
  if (npairs == 0) return 0;
  CK(cudaSetDevice(c->device));
  std::vector<PairDesc> d(npairs);
  for (int i = 0; i < npairs; i++) {
    if (n1[i] < 0 || n2[i] < 0 || (n1[i] > 0 && !scan1[i]) || (n2[i] > 0 && !scan2[i]))
      return fail(ICET_B200_E_INVALID, "bad scan pointer / size in pair " + std::to_string(i));
    d[i] = PairDesc{scan1[i], scan2[i], n1[i], n1[i], n2[i], n2[i]};
  }
  return batch_device_impl(c, p, npairs, d.data(), x0, out, false);
}

int icet_b200_register_sequence_device(icet_b200_ctx* c, const icet_b200_params* p, int32_t nscans,
                                       const float* scans, int32_t n, icet_b200_result* out) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  int rc = validate(p);
  if (rc) return rc;
  if (nscans < 1 || n < 0 || !scans || (nscans > 1 && !out)) return fail(ICET_B200_E_INVALID, "bad argument");
  if (nscans == 1) return 0;
  CK(cudaSetDevice(c->device));
  std::vector<PairDesc> d(nscans - 1);
  for (int i = 0; i < nscans - 1; i++)
    d[i] = PairDesc{scans + (size_t)i * 3 * n, scans + (size_t)(i + 1) * 3 * n, n, n, n, n};
  return batch_device_impl(c, p, nscans - 1, d.data(), nullptr, out, false);
}

// -- host-buffer entry points ------------------------------------------------------------------
int icet_b200_register_batch(icet_b200_ctx* c, const icet_b200_params* p, int32_t npairs,
                             const float* const* scan1, const int32_t* n1, const float* const* scan2,
                             const int32_t* n2, const float* x0, icet_b200_result* out) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  int rc = validate(p);
  if (rc) return rc;
  if (npairs < 0 || (npairs > 0 && (!scan1 || !scan2 || !n1 || !n2 || !out)))
    return fail(ICET_B200_E_INVALID, "NULL argument");
  if (npairs == 0) return 0;
  for (int i = 0; i < npairs; i++)
    if (n1[i] < 0 || n2[i] < 0 || (n1[i] > 0 && !scan1[i]) || (n2[i] > 0 && !scan2[i]))
      return fail(ICET_B200_E_INVALID, "bad scan pointer / size in pair " + std::to_string(i));
  CK(cudaSetDevice(c->device));
  rc = c->resbuf.ensure((size_t)npairs * sizeof(icet_b200_result));
  if (rc) return rc;
  icet_b200_result* d_res = (icet_b200_result*)c->resbuf.p;
  // Upload/compute pipeline: the batch is cut into chunks; chunk k+1 is uploaded (copy stream) while chunk k is
  // registered (compute stream).  The link is the bottleneck at 64-channel size (1.5 MB per scan over PCIe against
  // ~25 us of GPU time per pair), so the chunks are small (host_chunk pairs) and ramp up from / down to an eighth of
  // that: the first upload and the last registration are the only parts that nothing overlaps.
  const int CH = std::max(1, std::min(c->chunk_pairs, c->host_chunk));
  std::vector<int> sizes;
  {
    int left = npairs;
    std::vector<int> head, tail;
    for (int s = std::max(1, CH / 8); s < CH && left > 2 * CH; s *= 2) {
      head.push_back(s);
      tail.push_back(s);
      left -= 2 * s;
    }
    sizes = head;
    while (left > 0) {
      const int s = std::min(CH, left);
      sizes.push_back(s);
      left -= s;
    }
    for (size_t k = tail.size(); k-- > 0;) sizes.push_back(tail[k]);
  }
  const size_t pin_slot = (size_t)CH * (sizeof(PairDesc) + 6 * sizeof(float)) + 2048;
  rc = ensure_pinned(c, pin_slot * ICET_NSLOT + 4096);
  if (rc) return rc;
  auto pad = [](size_t f) { return (f + 63) & ~(size_t)63; };
  int base = 0;
  const float* prev_scan2_dev = nullptr;  // device copy of the previous chunk's last scan 2
  const bool chain = (p->flags & ICET_B200_FLAG_CHAIN_X0) != 0;  // chunks depend on each other: one lane, in order
  // (the host pipeline is bound by the link, two lanes are plenty)
  const bool two = c->nlanes > 1 && sizes.size() > 1 && !(c->dump_on && npairs == 1) && !chain;
  if (two) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(c->lanes[1], c->ev_fork, 0));
  }
  for (size_t k = 0; k < sizes.size(); base += sizes[k], k++) {
    const int P = sizes[k];
    const int slot = (int)(k % ICET_NSLOT);
    // plan the staging area: a scan shared by consecutive pairs (scan2[g-1] == scan1[g]) is uploaded once, also
    // across the chunk boundary (the previous chunk's staging slot is still intact, see the waits below)
    std::vector<PairDesc> d(P);
    std::vector<size_t> off1(P), off2(P);
    std::vector<char> up1(P);
    size_t floats = 0;
    int n1max = 0, n2max = 0;
    for (int i = 0; i < P; i++) {
      const int g = base + i;
      const bool shared = g > 0 && scan1[g] == scan2[g - 1] && n1[g] == n2[g - 1] && (i > 0 || prev_scan2_dev);
      up1[i] = !shared;
      if (shared) {
        off1[i] = (i > 0) ? off2[i - 1] : (size_t)-1;
      } else {
        off1[i] = floats;
        floats += pad((size_t)3 * n1[g]);
      }
      off2[i] = floats;
      floats += pad((size_t)3 * n2[g]);
      n1max = std::max(n1max, n1[g]);
      n2max = std::max(n2max, n2[g]);
    }
    // Slot reuse: this slot last held chunk k - NSLOT; chunk k - NSLOT + 1 may still read its last scan (shared
    // across the boundary), so wait for THAT chunk's registration before overwriting.
    CK(cudaEventSynchronize(c->ev_done[slot]));
    CK(cudaEventSynchronize(c->ev_done[(slot + 1) % ICET_NSLOT]));
    rc = c->stage[slot].ensure(floats * sizeof(float));
    if (rc) return rc;
    rc = c->descbuf[slot].ensure((size_t)CH * sizeof(PairDesc));
    if (rc) return rc;
    rc = c->x0buf[slot].ensure((size_t)CH * 6 * sizeof(float));
    if (rc) return rc;
    float* sbase = (float*)c->stage[slot].p;
    // uploads; runs that are contiguous on both sides are merged into one copy
    const float* run_src = nullptr;
    float* run_dst = nullptr;
    size_t run_floats = 0;
    auto flush = [&]() -> int {
      if (run_floats)
        CK(cudaMemcpyAsync(run_dst, run_src, run_floats * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
      run_floats = 0;
      return 0;
    };
    auto upload = [&](const float* src, float* dst, size_t nf) -> int {
      if (nf == 0) return 0;
      if (run_floats && src == run_src + run_floats && dst == run_dst + run_floats) {
        run_floats += nf;
        return 0;
      }
      int r = flush();
      if (r) return r;
      run_src = src;
      run_dst = dst;
      run_floats = nf;
      return 0;
    };
    for (int i = 0; i < P; i++) {
      const int g = base + i;
      if (up1[i]) {
        rc = upload(scan1[g], sbase + off1[i], (size_t)3 * n1[g]);
        if (rc) return rc;
      }
      rc = upload(scan2[g], sbase + off2[i], (size_t)3 * n2[g]);
      if (rc) return rc;
      const float* s1p = up1[i] ? sbase + off1[i] : (i > 0 ? sbase + off2[i - 1] : prev_scan2_dev);
      d[i] = PairDesc{s1p, sbase + off2[i], n1[g], n1[g], n2[g], n2[g]};
    }
    rc = flush();
    if (rc) return rc;
    prev_scan2_dev = sbase + off2[P - 1];
    char* hp = (char*)c->pinned + (size_t)slot * pin_slot;
    memcpy(hp, d.data(), (size_t)P * sizeof(PairDesc));
    CK(cudaMemcpyAsync(c->descbuf[slot].p, hp, (size_t)P * sizeof(PairDesc), cudaMemcpyHostToDevice, c->copy_stream));
    const float* d_x0 = nullptr;
    if (chain && base > 0) {
      d_x0 = d_res[base - 1].X;  // the previous chunk's last solution (same stream, already ordered)
    } else if (x0) {
      const size_t nx = chain ? 6 : (size_t)P * 6;  // chained: one seed, that of pair 0
      float* hx = (float*)(hp + (size_t)CH * sizeof(PairDesc));
      memcpy(hx, x0 + (size_t)base * 6, nx * sizeof(float));
      CK(cudaMemcpyAsync(c->x0buf[slot].p, hx, nx * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
      d_x0 = (const float*)c->x0buf[slot].p;
    }
    CK(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    const int lane = two ? (int)(k & 1) : 0;
    cudaStream_t lst = lane == 0 ? c->stream : c->lanes[1];
    CK(cudaStreamWaitEvent(lst, c->ev_copy[slot], 0));
    const bool dump = c->dump_on && npairs == 1;
    if (dump) {
      rc = ensure_dump(c, p);
      if (rc) return rc;
    }
    rc = run_chunk(c, p, P, (const PairDesc*)c->descbuf[slot].p, n1max, n2max, d_x0, d_res + base, dump, lane);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev_done[slot], lst));
    if (dump) {
      c->dump_params = *p;
      c->dump_valid = true;
    }
    if (npairs == 1) {
      c->last_n2 = n2[0];
      c->last_runlen = p->runlen;
    }
  }
  if (two) {
    CK(cudaEventRecord(c->ev_join[1], c->lanes[1]));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join[1], 0));
  }
  CK(cudaMemcpyAsync(out, d_res, (size_t)npairs * sizeof(icet_b200_result), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_loop_watchdog(c);
}

int icet_b200_register(icet_b200_ctx* c, const icet_b200_params* p, const float* scan1, int32_t n1, int32_t ld1,
                       const float* scan2, int32_t n2, int32_t ld2, const float x0[6], icet_b200_result* out) {
  if (!c) return fail(ICET_B200_E_INVALID, "ctx is NULL");
  if (!out) return fail(ICET_B200_E_INVALID, "out is NULL");
  if (n1 < 0 || n2 < 0 || ld1 < n1 || ld2 < n2) return fail(ICET_B200_E_INVALID, "bad n / ld");
  if (ld1 == n1 && ld2 == n2) return icet_b200_register_batch(c, p, 1, &scan1, &n1, &scan2, &n2, x0, out);
  // general leading dimension: pack the three planes on the host side of the copy
  std::vector<float> a((size_t)3 * n1), b((size_t)3 * n2);
  for (int k = 0; k < 3; k++) {
    if (n1) memcpy(a.data() + (size_t)k * n1, scan1 + (size_t)k * ld1, (size_t)n1 * sizeof(float));
    if (n2) memcpy(b.data() + (size_t)k * n2, scan2 + (size_t)k * ld2, (size_t)n2 * sizeof(float));
  }
  const float* pa = a.data();
  const float* pb = b.data();
  return icet_b200_register_batch(c, p, 1, &pa, &n1, &pb, &n2, x0, out);
}

int icet_b200_get_dump(icet_b200_ctx* c, icet_b200_voxel_dump* o) {
  if (!c || !o) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (!c->dump_valid) return fail(ICET_B200_E_INVALID, "no dump recorded: call icet_b200_set_dump(ctx,1) before icet_b200_register");
  CK(cudaSetDevice(c->device));
  const icet_b200_params& p = c->dump_params;
  const size_t ncell = (size_t)p.bins_phi * p.bins_theta, rl = (size_t)p.runlen;
  const Dump& d = c->dump_ptrs;
  // workspace-resident pieces: cnt1 and the cell records of pair 0
  Chunk ck;
  memset(&ck, 0, sizeof(ck));
  size_t zb;
  // NOTE: the workspace layout of the last (single-pair) chunk
  cudaStream_t st = c->stream;
  CK(cudaStreamSynchronize(st));
  // n1max/n2max only affect arrays behind rec/cnt1, which are carved first
  carve_chunk(c->ws[0].p, 1, (int)ncell, 0, 0, p.runlen, ck, &zb);
  if (o->cnt1) CK(cudaMemcpy(o->cnt1, ck.cnt1, ncell * 4, cudaMemcpyDeviceToHost));
  if (o->bounds) {
    std::vector<CellRec> rec(ncell);
    CK(cudaMemcpy(rec.data(), ck.rec, ncell * sizeof(CellRec), cudaMemcpyDeviceToHost));
    const int nT = p.bins_theta, nP = p.bins_phi;
    for (size_t cidx = 0; cidx < ncell; cidx++) {
      const int t = (int)(cidx % nT), q = (int)(cidx / nT);
      float* b = o->bounds + cidx * 6;
      b[0] = (static_cast<float>(t) / nT) * (2 * M_PI);
      b[1] = (static_cast<float>(t + 1) / nT) * (2 * M_PI);
      b[2] = (static_cast<float>(q) / nP) * (M_PI);
      b[3] = (static_cast<float>(q + 1) / nP) * (M_PI);
      b[4] = rec[cidx].inner;
      b[5] = rec[cidx].outer;
    }
  }
#define CP(dst, src, bytes) if (o->dst) CK(cudaMemcpy(o->dst, d.src, (bytes), cudaMemcpyDeviceToHost))
  CP(nin1, nin1, ncell * 4); CP(has1, has1, ncell); CP(mu1, mu1, ncell * 12); CP(sigma1, sigma1, ncell * 36);
  CP(evec1, evec1, ncell * 36); CP(eval1, eval1, ncell * 12); CP(lmask, lmask, ncell * 3);
  CP(cnt2, cnt2, rl * ncell * 4); CP(nin2, nin2, rl * ncell * 4); CP(used2, used2, rl * ncell);
  CP(mu2, mu2, rl * ncell * 12); CP(sigma2, sigma2, rl * ncell * 36); CP(Xit, Xit, rl * 24);
  CP(HTWH, HTWH, rl * 144); CP(HTWdz, HTWdz, rl * 24);
#undef CP
  return 0;
}

int icet_b200_debug_timeline(icet_b200_ctx* c, uint64_t* out, int32_t runlen) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (!c->dump_valid || runlen != c->dump_params.runlen) return fail(ICET_B200_E_INVALID, "no dump recorded");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out, c->dump_ptrs.tl, ((size_t)runlen * 16 + 6144) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

int icet_b200_get_points2(icet_b200_ctx* c, float* out, int32_t n2) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (!c->last_valid) return fail(ICET_B200_E_INVALID, "no single-pair registration to take points2 from");
  if (n2 != c->last_n2) return fail(ICET_B200_E_INVALID, "n2 does not match the last registration");
  if (n2 == 0) return 0;
  CK(cudaSetDevice(c->device));
  Chunk ck;
  memcpy(&ck, c->last_ck, sizeof(Chunk));
  // the staging buffers are idle here: use one for the device-side output
  CK(cudaStreamSynchronize(c->stream));
  int rc = c->resbuf.ensure((size_t)3 * n2 * sizeof(float) + 1024);
  if (rc) return rc;
  float* d_out = (float*)c->resbuf.p;
  k_points2<<<(n2 + 255) / 256, 256, 0, c->stream>>>(ck, n2, d_out);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d_out, (size_t)3 * n2 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int icet_b200_spherical_bins(icet_b200_ctx* c, const icet_b200_params* p, const float* scan, int32_t n, int32_t ld,
                             float* sph, int32_t* cell) {
  if (!c || !scan || !sph || !cell) return fail(ICET_B200_E_INVALID, "NULL argument");
  int rc = validate(p);
  if (rc) return rc;
  if (n <= 0 || ld < n) return fail(ICET_B200_E_INVALID, "bad n / ld");
  CK(cudaSetDevice(c->device));
  rc = c->stage[0].ensure(((size_t)3 * ld + 3 * (size_t)n + n) * 4);
  if (rc) return rc;
  CK(cudaEventSynchronize(c->ev_done[0]));
  float* d_s = (float*)c->stage[0].p;
  float* d_sph = d_s + (size_t)3 * ld;
  int32_t* d_cell = (int32_t*)(d_sph + (size_t)3 * n);
  CK(cudaMemcpyAsync(d_s, scan, (size_t)3 * ld * 4, cudaMemcpyHostToDevice, c->stream));
  rc = ensure_edges(c, p->bins_theta, p->bins_phi);
  if (rc) return rc;
  const float *azE, *elE;
  icet::BinTable bth, bph;
  fill_tables(c, p->bins_theta, p->bins_phi, &azE, &elE, &bth, &bph);
  k_sph_bins<<<(n + 255) / 256, 256, 0, c->stream>>>(d_s, n, ld, p->bins_theta, p->bins_phi, bth, bph, d_sph, d_cell);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(sph, d_sph, (size_t)3 * n * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(cell, d_cell, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int icet_b200_synth_scans_device(icet_b200_ctx* c, uint64_t seed, int32_t first_scan, int32_t nscans, int32_t rings,
                                 int32_t azim, float* out) {
  if (!c || !out) return fail(ICET_B200_E_INVALID, "NULL argument");
  if (first_scan < 0 || nscans < 1 || rings < 1 || azim < 1) return fail(ICET_B200_E_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  // poses of scans first_scan .. first_scan+nscans-1 (composition of the per-step motions)
  std::vector<synth::Pose> poses(nscans);
  synth::Pose P;
  synth::pose_identity(P);
  for (int k = 0; k < first_scan + nscans; k++) {
    if (k >= first_scan) poses[k - first_scan] = P;
    double d[6];
    synth::drive_step(seed, k, P, d);
    synth::advance(P, d);
  }
  int rc = c->posebuf.ensure(poses.size() * sizeof(synth::Pose));
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpyAsync(c->posebuf.p, poses.data(), poses.size() * sizeof(synth::Pose), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const size_t total = (size_t)nscans * rings * azim;
  k_synth<<<(unsigned)((total + 127) / 128), 128, 0, c->stream>>>(seed, first_scan, nscans, rings, azim,
                                                                  (const synth::Pose*)c->posebuf.p, out);
  c->launches++;
  CK(cudaGetLastError());
  return 0;
}

#include "callers_abi.inl"

}  // extern "C"

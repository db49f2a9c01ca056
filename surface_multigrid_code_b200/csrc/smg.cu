// smg.cu -- the C ABI of libsmg.so (include/smg.h): handle, device residency of the
// multigrid hierarchy, min_quad_with_fixed_mg_precompute / _solve and the mg_VCycle
// operators, driven by the sm_100a kernels of kernels.cu.
//
// Reference behaviour replaced (file:line under the reference tree):
//   src/min_quad_with_fixed_mg.cpp:3-51, :137-257   precompute  -> smg_precompute
//   src/min_quad_with_fixed_mg.cpp:80-135, :288-361 solve       -> smg_solve[_device]
//   src/mg_VCycle.cpp:3-59                           mg_VCycle   -> smg_vcycle / vcycle_device
//   src/mg_VCycle.cpp:62-92, :113-201                A/restrict/prolong/relax/coarseSolve
//
// Data layout in HBM (per level l, n_l rows, everything FP64 / int32):
//   * rows are renumbered phase-major (smoother colour or wavefront level first, then
//     a breadth-first locality rank, then descending row length inside sigma windows);
//     all level vectors (b, u, r; n_l x k column-major) live in that numbering;
//   * A_l   : SELL-32 (col, val = column-i-as-row-i, valT = true row i), plus its CSC
//             value array in reference order (Galerkin input / parity read-back);
//   * P_l   : SELL-32 by fine row (columns in level l's numbering), PT_l: SELL-32 by
//             coarse row (columns in level l-1's numbering);
//   * coarsest level: dense symmetric inverse (n_c x n_c) for the direct solve.
// There is no CPU compute path in this file: without a CUDA device every compute
// entry point fails with SMG_E_CUDA / SMG_E_STATE.
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <nvtx3/nvToolsExt.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/smg.h"
#include "kernels.hpp"
#include "patch.hpp"
#include "plan.hpp"

namespace {

using smg::Csc;
using smg::SellDev;

// NVTX range (header-only NVTX 3: a no-op unless a profiler is attached) around the host-visible
// phases of the entry points and, when the V-cycle is not replayed from a graph, around every level
struct NvtxRange {
  bool open = true;
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  void end() {
    if (open) nvtxRangePop();
    open = false;
  }
  ~NvtxRange() { end(); }  // (early error returns close their ranges too)
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---- device buffers ----------------------------------------------------------
// Stream-ordered allocation (cudaMallocAsync / cudaFreeAsync) on the stream of the handle
// the calling thread is working on: cudaMalloc / cudaFree synchronise the whole device, and
// a rank that is behind must never wait for the kernels of a rank that is ahead and already
// spinning in a halo exchange (ranks may share a device).
thread_local cudaStream_t tls_stream = nullptr;
thread_local bool tls_async = false;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      if (tls_async) cudaFreeAsync(p, tls_stream);
      else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e = tls_async ? cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), tls_stream)
                              : cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
    if (e == cudaSuccess) n = count;
    else p = nullptr;
    return e;
  }
  cudaError_t reserve(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t st) {
    cudaError_t e = alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
  }
};

struct SellBufs {
  DevBuf<int> slice_ptr, col, src;
  DevBuf<double> val, valT;
  int nrows = 0, nslices = 0, max_chunk = 0, max_width = 0, max_chunk32 = 0, max_chunk16 = 0,
      max_chunk32s = 0;
  int64_t padded = 0;
  SellDev view() const {
    SellDev d;
    d.nrows = nrows;
    d.nslices = nslices;
    d.max_chunk = max_chunk;
    d.max_width = max_width;
    d.max_chunk32 = max_chunk32;
    d.max_chunk16 = max_chunk16;
    d.max_chunk32s = max_chunk32s;
    d.slice_ptr = slice_ptr.p;
    d.col = col.p;
    d.val = val.p;
    d.valT = valT.p ? valT.p : val.p;
    return d;
  }
};

// one exchange pattern (plan.hpp::Exchange) as seen by this rank
struct ExchDev {
  DevBuf<smg::XchgPeer> peers;
  std::vector<DevBuf<int>> lists;
  bool any = false;  // some pair of ranks moves data (otherwise the exchange is skipped by all)
  int ctas = 1;      // CTAs per peer: one per 4096 values of the longest list
};

struct LevelDev {
  int n = 0;
  // CSC of mg[l].A in reference order (pattern static, values numeric)
  DevBuf<int> a_colptr, a_rowidx, a_col, tmap, diag_pos;
  DevBuf<double> a_val;
  SellBufs sellA;
  DevBuf<double> diag;  // permuted numbering
  DevBuf<int> perm;     // new -> old
  std::vector<int> phase_ptr;
  // transfer operators between level l-1 (fine) and l (coarse), l >= 1
  SellBufs sellP, sellPT;
  DevBuf<int> p_colptr, p_rowidx, pt_colptr, pt_rowidx;  // CSC of P and of PT
  DevBuf<double> p_val, pt_val;
  DevBuf<int> t_colptr, t_rowidx, t_col;  // T1 = PT * A_{l-1}
  DevBuf<double> t_val;
  // work vectors (permuted numbering), n x kcap column-major, ld = n
  DevBuf<double> b, u, r;
  // communication-avoiding patch smoother (patch.hpp): a relax call of this level in one
  // launch; u2 is the second buffer of u (a patch launch reads one and writes the other)
  struct PatchBufs {
    DevBuf<unsigned char> blob;
    DevBuf<long long> off;
    DevBuf<int> fill_dst, fill_src;
    int64_t n_fill = 0;
    smg::PatchDev view;
    int iters = -1;
  } patch_down, patch_up;
  bool patched = false;
  DevBuf<double> u2;
  // rows this rank smooths, one range per phase (all rows of the phase unless the level is
  // row-partitioned), and the rows it applies operators to
  std::vector<std::pair<int, int>> gs_ranges;
  int layout = smg::LAYOUT_PLAIN;
  int own_b = 0, own_e = 0;                          // PARTITIONED: this rank's rows
  std::vector<std::pair<int, int>> restrict_ranges;  // rows of PT this rank computes
  ExchDev x_halo_u, x_halo_r, x_halo_pu, x_gather;
  std::vector<ExchDev> x_halo_u_phase;
};

struct HostGroup;

// multi-GPU context: the comm buffer of every rank is mapped into every other rank
struct DistCtx {
  int rank = 0, world = 1;
  bool connected = false;
  int exact = 0;  // halo exchange after every colour (1: same result as one GPU), sweep (0), relax call (2)
  size_t comm_bytes = 0, flags_bytes = 0, slot_doubles = 0;
  char* comm = nullptr;
  std::vector<char*> peer;    // world entries, peer[rank] == comm
  std::vector<char> opened;   // peer[q] came from cudaIpcOpenMemHandle
  DevBuf<int> ctrl;
  DevBuf<double> normv;
  ExchDev x_norm;
  ExchDev x_slices;  // rows [slice_b(s), slice_b(s+1)) of the caller's n-vectors: rank s -> everybody (sharded H2D)
  bool slices_ok = false;
  int64_t exchanges = 0;
  bool shared_device = false;  // some other rank uses this device too (tests): late PDL trigger
  std::shared_ptr<HostGroup> host_group;  // non-null: host-synchronised exchanges (see HostGroup)
  bool failed = false;  // an exchange timed out: results are garbage, every later call fails
};
constexpr size_t kFlagStride = 128;  // header region of the comm buffer: one line per rank (reserved)

// Host rendezvous of ranks that live in one process AND share a device (the test layout on a
// single-GPU box).  Such ranks share hardware queues, copy engines, the allocator and the
// module loader of one context, and a kernel that spins for a peer can starve exactly the
// work the peer needs to issue.  They therefore never wait on the device: every exchange is
// "push kernel, stream sync, host barrier, receive kernel" (no CUDA graphs in this mode).
// One process per GPU, the production layout, uses the single fused kernel that waits on the
// device.
struct HostGroup {
  std::mutex m;
  std::condition_variable cv;
  int size = 0, count = 0;
  long long gen = 0;
  bool broken = false;
  bool wait() {  // false: a rank did not arrive within 60 s (or the group is broken)
    std::unique_lock<std::mutex> lk(m);
    if (broken) return false;
    const long long g = gen;
    if (++count == size) {
      count = 0;
      gen++;
      cv.notify_all();
      return true;
    }
    if (!cv.wait_for(lk, std::chrono::seconds(60), [&] { return gen != g || broken; })) broken = true;
    if (broken) cv.notify_all();
    return !broken;
  }
};
std::mutex g_groups_mutex;
std::map<unsigned long long, std::weak_ptr<HostGroup>> g_groups;

struct DistBlob {  // what smg_dist_get_handle exports (smg_dist_handle_bytes() bytes)
  cudaIpcMemHandle_t ipc;
  long long pid;
  unsigned long long ptr;
  long long bytes;
  int device;
  int rank;
  char uuid[16];  // of the device: two ranks on one physical GPU are detected by it
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  int64_t first = 0;  // solve-loop graphs: kernels before the loop
};

}  // namespace

struct smg_handle {
  smg_options opt;
  bool plan_only = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  cublasHandle_t cublas = nullptr;
  std::string err;

  bool have_hierarchy = false, have_plan = false;
  std::vector<int> n_rows;
  std::vector<Csc> P_full;
  smg::Plan plan;
  std::vector<LevelDev> lv;
  int kcap = 0;

  // system gather / scatter
  DevBuf<double> a_in;  // caller's A values
  size_t a_nnz = 0;     // entries of the last precomputed A (a_in.n is a capacity, never shrinks)
  DevBuf<int> lhs_src, auk_src;
  DevBuf<int> g;                          // permuted unknown row -> caller index
  DevBuf<int> auk_ptr, auk_q, auk_pos;    // Auk by permuted row; auk_pos -> entry of Auk CSC
  DevBuf<double> auk_csc_val, auk_val;
  DevBuf<int> kidx, ksrc;
  int n_known_distinct = 0;

  // coarse direct solve
  DevBuf<double> ainv, ainv_tiles, coarse_scratch, linv;
  DevBuf<double> potrf_work;
  DevBuf<int> dev_info;

  // staging
  DevBuf<double> st_a, st_b, st_c, st_d;
  DevBuf<double> norm_scratch, norm_out;
  DevBuf<unsigned int> norm_counter;  // zero between launches
  double* h_norm = nullptr;  // pinned
  DevBuf<double> flush;      // L2 flush buffer for smg_time_kernel

  std::map<std::tuple<int, int, int, int>, GraphEntry> graphs;
  int patch_kcols = 1;  // right-hand-side columns the patch layouts are sized for
  // device-side solve loop: one graph per k (residual norm, test, conditional WHILE around the
  // V-cycle); loop_state < 0: not supported here (the host loop is used)
  std::map<int, GraphEntry> loop_graphs;
  DevBuf<smg::SolveCtl> loop_ctl;
  smg::SolveCtl* h_ctl = nullptr;  // pinned
  int loop_state = 0;
  int last_solve_on_device = 0;  // the last solve ran its loop on the device (smg_solve_on_device)
  int64_t launches = 0;
  double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // mean-curvature-flow assembly (smg_mcf_*)
  struct {
    bool ready = false;
    int nV = 0, nF = 0;
    double delta = 0.0;
    DevBuf<int> F, vf_ptr, vf_face;
    DevBuf<double> Lval, dblA, mass, U, rhs, z;
  } mcf;
  bool no_prefetch = false;
  bool async_alloc = false;
  DistCtx dist;
  int dist_levels = -1, dist_min_rows = 0;
};

namespace {

bool dist_on(const smg_handle* h) { return h->dist.world > 1; }

int fail(smg_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define SMG_CUDA(h, call)                                                          \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      return fail(h, SMG_E_CUDA,                                                   \
                  std::string(#call) + ": " + cudaGetErrorString(e__));            \
  } while (0)

#define SMG_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != SMG_OK) return rc__; \
  } while (0)

int check_launch(smg_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, SMG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return SMG_OK;
}

void drop_graphs(smg_handle* h) {
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
  for (auto& kv : h->loop_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->loop_graphs.clear();
}

int upload_sell(smg_handle* h, const smg::Sell& S, SellBufs* out, bool need_valT) {
  out->nrows = S.nrows;
  out->nslices = S.nslices;
  out->padded = S.padded();
  out->max_chunk = out->max_width = out->max_chunk32 = out->max_chunk16 = out->max_chunk32s = 0;
  for (int s0 = 0; s0 < S.nslices; s0++) {  // any 8 consecutive slices: one CTA's TMA chunk
    const int s1 = std::min(S.nslices, s0 + 8);
    out->max_chunk = std::max(out->max_chunk, S.slice_ptr[s1] - S.slice_ptr[s0]);
    out->max_width = std::max(out->max_width, (S.slice_ptr[s0 + 1] - S.slice_ptr[s0]) / 32);
    out->max_chunk16 = std::max(out->max_chunk16, S.slice_ptr[std::min(S.nslices, s0 + 16)] - S.slice_ptr[s0]);
    out->max_chunk32s = std::max(out->max_chunk32s, S.slice_ptr[std::min(S.nslices, s0 + 32)] - S.slice_ptr[s0]);
  }
  // a row range may start at any slice (row-partitioned levels): any 32 consecutive slices
  out->max_chunk32 = out->max_chunk32s;
  SMG_CUDA(h, out->slice_ptr.upload(S.slice_ptr, h->stream));
  SMG_CUDA(h, out->col.upload(S.col, h->stream));
  SMG_CUDA(h, out->src.upload(S.src, h->stream));
  SMG_CUDA(h, out->val.alloc(static_cast<size_t>(S.padded())));
  if (need_valT) SMG_CUDA(h, out->valT.alloc(static_cast<size_t>(S.padded())));
  else out->valT.release();
  return SMG_OK;
}

// device-side view of one exchange pattern for this rank; peers in ascending rank order
int build_exchange(smg_handle* h, const smg::Exchange& X, ExchDev* out) {
  DistCtx& D = h->dist;
  const int W = D.world, me = D.rank;
  out->lists.clear();
  out->any = !X.empty();
  out->ctas = static_cast<int>(std::min<size_t>(32, std::max<size_t>(1, (X.max_count() + 4095) / 4096)));
  if (X.idx.size() != static_cast<size_t>(W) * W) {
    out->any = false;
    return SMG_OK;
  }
  if ((X.max_count() * smg::kMaxK + 1) * 2 > D.slot_doubles)  // 16 bytes per value, + the sync word
    return fail(h, SMG_E_INVALID,
                "halo exchange of " + std::to_string(X.max_count()) + " rows does not fit the comm buffer; "
                "raise smg_dist_init's comm_bytes");
  std::vector<smg::XchgPeer> peers;
  for (int q = 0; q < W; q++) {
    if (q == me) continue;
    const std::vector<int>& snd = X.idx[static_cast<size_t>(me) * W + q];
    const std::vector<int>& rcv = X.idx[static_cast<size_t>(q) * W + me];
    smg::XchgPeer pr;
    pr.n_send = static_cast<int>(snd.size());
    pr.n_recv = static_cast<int>(rcv.size());
    out->lists.emplace_back();
    SMG_CUDA(h, out->lists.back().upload(snd, h->stream));
    pr.send_idx = out->lists.back().p;
    out->lists.emplace_back();
    SMG_CUDA(h, out->lists.back().upload(rcv, h->stream));
    pr.recv_idx = out->lists.back().p;
    // slot s of a rank's staging area holds what rank s sent
    pr.remote_slot = reinterpret_cast<double*>(D.peer[q] + D.flags_bytes) + static_cast<size_t>(me) * D.slot_doubles;
    pr.local_slot = reinterpret_cast<const double*>(D.comm + D.flags_bytes) + static_cast<size_t>(q) * D.slot_doubles;
    peers.push_back(pr);
  }
  SMG_CUDA(h, out->peers.upload(peers, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));  // host vectors go out of scope
  return SMG_OK;
}

// ranges, exchange lists of every level for this rank (after upload of the matrices)
int upload_dist(smg_handle* h) {
  smg::Plan& pl = h->plan;
  DistCtx& D = h->dist;
  const int nlev = static_cast<int>(pl.lv.size());
  for (int l = 0; l < nlev; l++) {
    const smg::LevelPlan& P = pl.lv[l];
    LevelDev& L = h->lv[l];
    const std::vector<int>& pp = P.order.phase_ptr;
    const int np = P.n_phases, W = P.nparts;
    L.layout = P.layout;
    L.gs_ranges.clear();
    L.restrict_ranges.clear();
    L.own_b = 0;
    L.own_e = P.n;
    if (P.layout == smg::LAYOUT_PARTITIONED) {
      for (int p = 0; p < np; p++)
        L.gs_ranges.emplace_back(pp[P.group(D.rank, p)], pp[P.group(D.rank, p) + 1]);
      L.own_b = pp[static_cast<size_t>(D.rank) * np];
      L.own_e = pp[static_cast<size_t>(D.rank + 1) * np];
      L.restrict_ranges.emplace_back(L.own_b, L.own_e);
    } else if (P.layout == smg::LAYOUT_SPLIT) {
      for (int p = 0; p < np; p++) {
        L.gs_ranges.emplace_back(pp[static_cast<size_t>(p) * W], pp[static_cast<size_t>(p + 1) * W]);
        L.restrict_ranges.emplace_back(pp[P.group(D.rank, p)], pp[P.group(D.rank, p) + 1]);
      }
    } else {
      for (int p = 0; p < np; p++) L.gs_ranges.emplace_back(pp[p], pp[p + 1]);
      L.restrict_ranges.emplace_back(0, P.n);
    }
    if (!dist_on(h)) continue;
    if (P.nparts > 1) SMG_TRY(build_exchange(h, P.gather_all, &L.x_gather));
    if (P.layout == smg::LAYOUT_PARTITIONED) {
      SMG_TRY(build_exchange(h, P.halo_u, &L.x_halo_u));
      SMG_TRY(build_exchange(h, P.halo_r, &L.x_halo_r));
      SMG_TRY(build_exchange(h, P.halo_pu, &L.x_halo_pu));
      L.x_halo_u_phase.clear();
      L.x_halo_u_phase.resize(P.halo_u_phase.size());
      for (size_t p = 0; p < P.halo_u_phase.size(); p++)
        SMG_TRY(build_exchange(h, P.halo_u_phase[p], &L.x_halo_u_phase[p]));
    }
  }
  if (dist_on(h)) {  // partial sums of the residual norm: slot s of the row <- rank s
    smg::Exchange X;
    X.idx.assign(static_cast<size_t>(D.world) * D.world, {});
    for (int s = 0; s < D.world; s++)
      for (int d = 0; d < D.world; d++)
        if (s != d) X.idx[static_cast<size_t>(s) * D.world + d] = {s};
    SMG_TRY(build_exchange(h, X, &D.x_norm));
    // Sharded host-to-device copies of smg_solve: rank s copies only rows [n s / W, n (s + 1) / W)
    // of the caller's RHS / z0 over PCIe and the ranks complete each other's copies over
    // NVLink.  Optional: without room in the comm buffer every rank copies everything.
    D.slices_ok = false;
    const int n = h->plan.n, W = D.world;
    const size_t per = static_cast<size_t>((n + W - 1) / W);
    if (!D.host_group && (per * smg::kMaxK + 1) * 2 <= D.slot_doubles) {
      smg::Exchange S;
      S.idx.assign(static_cast<size_t>(W) * W, {});
      for (int src = 0; src < W; src++) {
        const int b = static_cast<int>(static_cast<int64_t>(n) * src / W), e = static_cast<int>(static_cast<int64_t>(n) * (src + 1) / W);
        std::vector<int> rows(static_cast<size_t>(e - b));
        std::iota(rows.begin(), rows.end(), b);
        for (int d = 0; d < W; d++)
          if (d != src) S.idx[static_cast<size_t>(src) * W + d] = rows;
      }
      SMG_TRY(build_exchange(h, S, &D.x_slices));
      D.slices_ok = D.x_slices.any;
    }
  }
  return SMG_OK;
}

int upload_patches(smg_handle* h, int k_cols);

int ensure_k(smg_handle* h, int k) {
  // the pinned residual scratch (h_norm: 1024 doubles, flags from [32], per-rank sums from
  // [64]) is sized for SMG_MAX_RHS columns and 64 ranks
  if (k > SMG_MAX_RHS)
    return fail(h, SMG_E_INVALID, "more than SMG_MAX_RHS right-hand-side columns");
  if (k <= h->kcap) return SMG_OK;
  drop_graphs(h);
  if (std::min(k, smg::kMaxK) > h->patch_kcols && !dist_on(h)) {  // (partitioned handles pre-size for kMaxK
    // columns but keep the one-column patch layout: wider solves fall back to the phase kernels there)
    // the patch layouts were sized for fewer columns per pass: lay them out again (smaller
    // patches where needed) and refill their matrix values
    bool any = false;
    for (auto& L : h->lv) any = any || L.patched;
    if (any) {
      SMG_TRY(upload_patches(h, std::min(k, smg::kMaxK)));
      for (auto& L : h->lv)
        if (L.patched)
          for (LevelDev::PatchBufs* B : {&L.patch_down, &L.patch_up}) {
            smg::launch_patch_fill(reinterpret_cast<double*>(B->blob.p), B->fill_dst.p, B->fill_src.p, L.a_val.p,
                                   B->n_fill, h->stream);
            h->launches++;
          }
      SMG_TRY(check_launch(h, "patch refill"));
    }
  }
  SMG_CUDA(h, h->coarse_scratch.reserve(smg::dense_sym_scratch_doubles(h->lv.back().n, k)));
  for (auto& L : h->lv) {
    const size_t cnt = static_cast<size_t>(L.n) * k;
    SMG_CUDA(h, L.b.alloc(cnt));
    SMG_CUDA(h, L.u.alloc(cnt));
    SMG_CUDA(h, L.r.alloc(cnt));
    if (L.patched) SMG_CUDA(h, L.u2.alloc(cnt));
  }
  h->kcap = k;
  return SMG_OK;
}

// ---- multi-GPU helpers ------------------------------------------------------------
// Ranks that share a device (tests) also share its copy engines, whose queues are FIFO: a
// copy enqueued behind a kernel that still spins in an exchange would block the copies of
// the very rank it waits for.  So wait for the stream before enqueuing a copy.
int drain_if_shared(smg_handle* h) {
  if (h->dist.shared_device) SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMG_OK;
}

// collective: every rank runs the same sequence of exchanges (kernels.hpp::XchgPeer)
void exchange(smg_handle* h, ExchDev& X, double* vec, int ld, int k) {
  if (!dist_on(h) || !X.any) return;
  DistCtx& D = h->dist;
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    double* v = vec + static_cast<size_t>(k0) * ld;
    const size_t stride = D.slot_doubles * D.world;
    if (D.host_group) {
      smg::launch_halo_exchange(X.peers.p, D.world - 1, X.ctas, v, ld, kk, stride, D.ctrl.p, true, 1, h->stream);
      cudaStreamSynchronize(h->stream);
      if (!D.host_group->wait()) D.failed = true;  // reported by check_exchange
      smg::launch_halo_exchange(X.peers.p, D.world - 1, X.ctas, v, ld, kk, stride, D.ctrl.p, true, 2, h->stream);
      h->launches += 2;
    } else {
      smg::launch_halo_exchange(X.peers.p, D.world - 1, X.ctas, v, ld, kk, stride, D.ctrl.p, D.shared_device,
                                3, h->stream);
      h->launches++;
    }
    D.exchanges++;
  }
}

// ---- device-side operators on the level work vectors ---------------------------
// On a row-partitioned level every operator touches this rank's rows only; the halo
// exchanges that make the result usable by the next operator are issued here too.
void relax_device(smg_handle* h, int l, int iters, const double* b, double* u, int k) {
  LevelDev& L = h->lv[l];
  const SellDev A = L.sellA.view();
  const int np = static_cast<int>(L.gs_ranges.size());
  const bool part = dist_on(h) && L.layout == smg::LAYOUT_PARTITIONED;
  int p_first = -1, p_last = -1;  // non-empty phases
  for (int p = 0; p < np; p++)
    if (L.gs_ranges[p].second > L.gs_ranges[p].first) {
      if (p_first < 0) p_first = p;
      p_last = p;
    }
  if (iters <= 0 || (p_first < 0 && !part)) return;
  smg::GsFlow flow;
  // the columns of a block of right-hand sides are independent: one complete relax call
  // (its own epoch range) per group of kMaxK columns
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    double* uu = u + static_cast<size_t>(k0) * L.n;
    for (int it = 0; it < iters; it++) {
      for (int p = 0; p < np; p++) {
        const int ps = L.gs_ranges[p].first, pe = L.gs_ranges[p].second;
        if (pe > ps) {
          // next non-empty phase of this call (its matrix chunk is prefetched into L2)
          flow.pf_slice0 = -1;
          if (!(it == iters - 1 && p == p_last) && !h->no_prefetch) {
            int q = p;
            do q = (q + 1) % np; while (L.gs_ranges[q].second <= L.gs_ranges[q].first);
            flow.pf_slice0 = L.gs_ranges[q].first >> 5;
            flow.pf_slice_end = (L.gs_ranges[q].second + 31) >> 5;
          }
          smg::launch_gs_phase(A, L.diag.p, b + static_cast<size_t>(k0) * L.n, uu, L.n, kk, ps, pe,
                               flow, h->stream);
          h->launches++;
        }
        // exact mode: the rows of this colour reach the other ranks before the next colour
        if (part && h->dist.exact == 1) exchange(h, L.x_halo_u_phase[p], uu, L.n, kk);
      }
      // hybrid modes: Gauss-Seidel inside a rank; across ranks one halo exchange per sweep
      // (0) or per relax call (2: the sweeps of a call see the neighbours' rows as they were
      // when the call started)
      if (part && (h->dist.exact == 0 || (h->dist.exact == 2 && it == iters - 1)))
        exchange(h, L.x_halo_u, uu, L.n, kk);
    }
  }
}

// the matrix restricted to the rows this rank applies on level l
SellDev own_rows(const smg_handle* h, const LevelDev& L, const SellDev& M) {
  if (dist_on(h) && L.layout == smg::LAYOUT_PARTITIONED) return M.rows(L.own_b, L.own_e);
  return M;
}

void residual_device(smg_handle* h, int l, const double* b, const double* u, double* r, int k) {
  LevelDev& L = h->lv[l];
  if (L.n <= 0) return;
  const SellDev A = own_rows(h, L, L.sellA.view());
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    const size_t o = static_cast<size_t>(k0) * L.n;
    smg::launch_residual(A, b + o, u + o, r + o, L.n, kk, h->stream);
    h->launches++;
  }
}

void apply_A_device(smg_handle* h, int l, const double* u, double* y, int k) {
  LevelDev& L = h->lv[l];
  if (L.n <= 0) return;
  const SellDev A = own_rows(h, L, L.sellA.view());
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    const size_t o = static_cast<size_t>(k0) * L.n;
    smg::launch_spmv(A, true, u + o, L.n, y + o, L.n, kk, h->stream);
    h->launches++;
  }
}

// x on level l (fine), y on level l+1; zero != nullptr: also zero[...] = 0 (same shape as y).
// Partitioned fine level: x must carry valid halo_r values; every rank computes the rows of
// its part, and a replicated coarse level is completed with an all-gather.
void restrict_device(smg_handle* h, int l, const double* x, double* y, int k, double* zero = nullptr) {
  LevelDev& F = h->lv[l];
  LevelDev& C = h->lv[l + 1];
  if (C.n <= 0) return;
  const bool part = dist_on(h) && F.layout == smg::LAYOUT_PARTITIONED;
  if (part && zero) {  // the halo rows of the coarse guess are zero too
    smg::launch_fill(zero, 0.0, static_cast<int64_t>(C.n) * k, h->stream);
    h->launches++;
  }
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    const double* xx = x + static_cast<size_t>(k0) * F.n;
    const size_t oc = static_cast<size_t>(k0) * C.n;
    if (part) {
      // one launch for up to kMaxRanges ranges (one per colour on the split level)
      const auto& rgs = C.restrict_ranges;
      for (size_t r0 = 0; r0 < rgs.size(); r0 += smg::kMaxRanges) {
        smg::RowRanges rr;
        for (size_t r = r0; r < std::min(rgs.size(), r0 + smg::kMaxRanges); r++)
          if (rgs[r].second > rgs[r].first) {
            rr.rb[rr.n] = rgs[r].first;
            rr.re[rr.n] = rgs[r].second;
            rr.n++;
          }
        if (rr.n == 0) continue;
        smg::launch_spmv_ranges(C.sellPT.view(), rr, xx, F.n, y + oc, C.n, kk, h->stream);
        h->launches++;
      }
      continue;
    }
    if (zero) smg::launch_spmv_zero(C.sellPT.view(), xx, F.n, y + oc, zero + oc, C.n, kk, h->stream);
    else smg::launch_spmv(C.sellPT.view(), false, xx, F.n, y + oc, C.n, kk, h->stream);
    h->launches++;
  }
  if (part && C.layout == smg::LAYOUT_SPLIT) exchange(h, C.x_gather, y, C.n, k);
}

// y (level l) = P x (level l+1)   /   u (level l) += P x
void prolong_device(smg_handle* h, int l, double* x, double* y, int k, bool add) {
  LevelDev& F = h->lv[l];
  LevelDev& C = h->lv[l + 1];
  if (F.n <= 0) return;
  const bool part = dist_on(h) && F.layout == smg::LAYOUT_PARTITIONED;
  if (part && C.layout == smg::LAYOUT_PARTITIONED) exchange(h, C.x_halo_pu, x, C.n, k);
  const SellDev P = own_rows(h, F, C.sellP.view());
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    const double* xx = x + static_cast<size_t>(k0) * C.n;
    double* yy = y + static_cast<size_t>(k0) * F.n;
    if (add) smg::launch_prolong_add(P, xx, C.n, yy, F.n, kk, h->stream);
    else smg::launch_spmv(P, false, xx, C.n, yy, F.n, kk, h->stream);
    h->launches++;
  }
  if (part && add) exchange(h, F.x_halo_u, y, F.n, k);  // the corrected halo rows of the neighbours
}

void coarse_solve_device(smg_handle* h, const double* b, double* u, int k) {
  LevelDev& L = h->lv.back();
  if (L.n <= 0) return;
  for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
    const int kk = std::min(smg::kMaxK, k - k0);
    const size_t o = static_cast<size_t>(k0) * L.n;
    smg::launch_dense_sym_add(h->ainv_tiles.p, b + o, u + o, h->coarse_scratch.p, L.n, kk, h->stream);
    h->launches += 2;
  }
}

bool use_patches(const smg_handle* h, int l, int pre, int post, int k);

// mg_VCycle (src/mg_VCycle.cpp:3-59) unrolled over levels; operates on the resident
// work vectors lv[l].b / .u of levels l >= lv0.
void vcycle_device(smg_handle* h, int lv0, int pre, int post, int k) {
  const int last = static_cast<int>(h->lv.size()) - 1;
  char label[32];
  for (int l = lv0; l < last; l++) {
    LevelDev& L = h->lv[l];
    LevelDev& C = h->lv[l + 1];
    std::snprintf(label, sizeof(label), "L%d down", l);
    smg::trace_label(label);
    NvtxRange nvtx_level(label);
    if (use_patches(h, l, pre, post, k)) {  // :36-47 in one launch; u moves to the second buffer
      for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
        const int kk = std::min(smg::kMaxK, k - k0);
        const size_t o = static_cast<size_t>(k0) * L.n, oc = static_cast<size_t>(k0) * C.n;
        // what the V-cycle launches next on a patched level: the down leg of the coarser
        // level, or (deepest patched level) this level's up leg after the coarse solve
        const smg::PatchDev* next = l + 1 < last && use_patches(h, l + 1, pre, post, k) ? &C.patch_down.view
                                                                                         : &L.patch_up.view;
        smg::launch_patch(L.patch_down.view, smg::PATCH_DOWN, L.u.p + o, L.u2.p + o, L.b.p + o, L.n, nullptr,
                          C.b.p + oc, C.u.p + oc, C.n, kk, next, h->stream);
        h->launches++;
      }
      continue;
    }
    relax_device(h, l, pre, L.b.p, L.u.p, k);              // :36
    residual_device(h, l, L.b.p, L.u.p, L.r.p, k);         // :41-42
    if (dist_on(h) && L.layout == smg::LAYOUT_PARTITIONED) exchange(h, L.x_halo_r, L.r.p, L.n, k);
    restrict_device(h, l, L.r.p, C.b.p, k, C.u.p);         // :44 and uc = 0 (:46-47), fused
  }
  std::snprintf(label, sizeof(label), "L%d", last);
  smg::trace_label(label);
  coarse_solve_device(h, h->lv[last].b.p, h->lv[last].u.p, k);  // :28-33
  for (int l = last - 1; l >= lv0; l--) {
    LevelDev& L = h->lv[l];
    LevelDev& C = h->lv[l + 1];
    std::snprintf(label, sizeof(label), "L%d up", l);
    smg::trace_label(label);
    NvtxRange nvtx_level(label);
    if (use_patches(h, l, pre, post, k)) {  // :52-56 in one launch; u returns to the first buffer
      for (int k0 = 0; k0 < k; k0 += smg::kMaxK) {
        const int kk = std::min(smg::kMaxK, k - k0);
        const size_t o = static_cast<size_t>(k0) * L.n, oc = static_cast<size_t>(k0) * C.n;
        const smg::PatchDev* next = l > lv0 && use_patches(h, l - 1, pre, post, k) ? &h->lv[l - 1].patch_up.view : nullptr;
        smg::launch_patch(L.patch_up.view, smg::PATCH_UP, L.u2.p + o, L.u.p + o, L.b.p + o, L.n, C.u.p + oc, nullptr,
                          nullptr, C.n, kk, next, h->stream);
        h->launches++;
      }
      continue;
    }
    prolong_device(h, l, C.u.p, L.u.p, k, true);  // :52-53
    relax_device(h, l, post, L.b.p, L.u.p, k);    // :56
  }
}

int vcycle_run(smg_handle* h, int lv0, int pre, int post, int k) {
  if (!h->opt.use_graph) {
    vcycle_device(h, lv0, pre, post, k);
    return check_launch(h, "vcycle");
  }
  const auto key = std::make_tuple(lv0, pre, post, k);
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    GraphEntry ge;
    cudaGraph_t graph = nullptr;
    const int64_t before = h->launches;
    SMG_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    vcycle_device(h, lv0, pre, post, k);
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    ge.launches = h->launches - before;
    h->launches = before;
    if (e != cudaSuccess) return fail(h, SMG_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(h, SMG_E_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    it = h->graphs.emplace(key, ge).first;
  }
  SMG_CUDA(h, cudaGraphLaunch(it->second.exec, h->stream));
  h->launches += it->second.launches;
  return SMG_OK;
}

int residual_norm_device(smg_handle* h, int l, const double* b, const double* u, int k,
                         double* out_host) {
  LevelDev& L = h->lv[l];
  DistCtx& D = h->dist;
  const bool part = dist_on(h) && L.layout == smg::LAYOUT_PARTITIONED;
  const int nb = smg::residual_norm_blocks(L.n);
  const int nchunks = (k + smg::kMaxK - 1) / smg::kMaxK;
  SMG_CUDA(h, h->norm_scratch.reserve(static_cast<size_t>(nb)));
  SMG_CUDA(h, h->norm_out.reserve(static_cast<size_t>(std::max(nchunks, 1))));
  if (L.n <= 0) {
    *out_host = 0.0;
    return SMG_OK;
  }
  if (part) SMG_CUDA(h, D.normv.reserve(static_cast<size_t>(D.world) * std::max(nchunks, 16)));
  const SellDev A = own_rows(h, L, L.sellA.view());
  for (int c = 0; c < nchunks; c++) {
    const int k0 = c * smg::kMaxK;
    const int kk = std::min(smg::kMaxK, k - k0);
    const size_t o = static_cast<size_t>(k0) * L.n;
    // partitioned: the sum over this rank's rows lands in slot `rank` of the chunk's row
    double* out = part ? D.normv.p + static_cast<size_t>(c) * D.world + D.rank : h->norm_out.p + c;
    smg::launch_residual_norm2(A, b + o, u + o, L.n, kk, h->norm_scratch.p, h->norm_counter.p, out, h->stream);
    h->launches += 1;
    if (part) exchange(h, D.x_norm, D.normv.p + static_cast<size_t>(c) * D.world, D.world, 1);
  }
  if (part) {
    // every rank adds the same partial sums in rank order: identical residuals everywhere
    int* xflag = reinterpret_cast<int*>(h->h_norm + 48);
    SMG_TRY(drain_if_shared(h));
    SMG_CUDA(h, cudaMemcpyAsync(h->h_norm + 64, D.normv.p, sizeof(double) * nchunks * D.world,
                                cudaMemcpyDeviceToHost, h->stream));
    SMG_CUDA(h, cudaMemcpyAsync(xflag, D.ctrl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    SMG_CUDA(h, cudaStreamSynchronize(h->stream));
    if (*xflag) {
      D.failed = true;
      return fail(h, SMG_E_INTERNAL, "halo exchange: a wait for a peer rank timed out");
    }
    double ss = 0.0;
    for (int c = 0; c < nchunks; c++)
      for (int q = 0; q < D.world; q++) ss += h->h_norm[64 + static_cast<size_t>(c) * D.world + q];
    *out_host = std::sqrt(ss);
    return SMG_OK;
  }
  SMG_CUDA(h, cudaMemcpyAsync(h->h_norm, h->norm_out.p, sizeof(double) * nchunks,
                              cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  double ss = 0.0;
  for (int c = 0; c < nchunks; c++) ss += h->h_norm[c];
  *out_host = std::sqrt(ss);
  return SMG_OK;
}

// ---- device-side solve loop ----------------------------------------------------------
// min_quad_with_fixed_mg.cpp:330-347 as ONE CUDA graph: [residual norm, test] followed by a
// conditional WHILE node whose body is [V-cycle, residual norm, test].  The test kernel
// (solve_decide_kernel) records the residual, and keeps the loop going while the residual is
// finite and not below the tolerance and fewer than max_iter measurements exist -- the
// reference's sequence of measurements and cycles exactly, including its last, unmeasured
// cycle.  Used for single-GPU handles with graphs enabled; partitioned handles sum their
// residual on the host (identical on every rank) and keep the host loop.
// partitioned level 0: every rank sums its own rows into slot `rank` of the chunk's row of
// normv, the exchange makes every slot valid on every rank, and the test kernel adds the slots
// in rank order: identical residuals, hence identical control flow, on all ranks
void enqueue_norm(smg_handle* h, int k) {
  LevelDev& L0 = h->lv[0];
  DistCtx& D = h->dist;
  const bool part = dist_on(h) && L0.layout == smg::LAYOUT_PARTITIONED;
  const SellDev A = own_rows(h, L0, L0.sellA.view());
  const int nchunks = (k + smg::kMaxK - 1) / smg::kMaxK;
  for (int c = 0; c < nchunks; c++) {
    const int k0 = c * smg::kMaxK, kk = std::min(smg::kMaxK, k - k0);
    const size_t o = static_cast<size_t>(k0) * L0.n;
    double* out = part ? D.normv.p + static_cast<size_t>(c) * D.world + D.rank : h->norm_out.p + c;
    smg::launch_residual_norm2(A, L0.b.p + o, L0.u.p + o, L0.n, kk, h->norm_scratch.p, h->norm_counter.p, out,
                               h->stream);
    h->launches++;
    if (part) exchange(h, D.x_norm, D.normv.p + static_cast<size_t>(c) * D.world, D.world, 1);
  }
}

int build_loop_graph(smg_handle* h, int k, GraphEntry* out) {
  LevelDev& L0 = h->lv[0];
  const int nchunks = (k + smg::kMaxK - 1) / smg::kMaxK;
  SMG_CUDA(h, h->norm_scratch.reserve(static_cast<size_t>(smg::residual_norm_blocks(L0.n))));
  SMG_CUDA(h, h->norm_out.reserve(static_cast<size_t>(std::max(nchunks, 1))));
  const bool part = dist_on(h) && L0.layout == smg::LAYOUT_PARTITIONED;
  if (part) SMG_CUDA(h, h->dist.normv.reserve(static_cast<size_t>(h->dist.world) * std::max(nchunks, 16)));
  // what the test kernel sums: the chunks' sums, or (partitioned) every rank's sum of every chunk
  const double* norm2 = part ? h->dist.normv.p : h->norm_out.p;
  const int n_norm2 = part ? nchunks * h->dist.world : nchunks;
  if (!h->loop_ctl.p) SMG_CUDA(h, h->loop_ctl.alloc(1));
  if (!h->h_ctl) SMG_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_ctl), sizeof(smg::SolveCtl)));
  cudaGraph_t g = nullptr;
  cudaGraphExec_t exec = nullptr;
  const int64_t before = h->launches;
  auto bail = [&](const char* what, cudaError_t e) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(h->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
      cudaGraph_t junk = nullptr;
      cudaStreamEndCapture(h->stream, &junk);
    }
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    h->launches = before;
    h->loop_state = -1;  // never tried again on this handle: the host loop is used
    h->err = std::string("device-side solve loop unavailable (") + what + ": " + cudaGetErrorString(e) + ")";
    return SMG_E_UNSUPPORTED;
  };
  cudaError_t e = cudaGraphCreate(&g, 0);
  if (e != cudaSuccess) return bail("graph create", e);
  cudaGraphConditionalHandle cond;
  if ((e = cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault)) != cudaSuccess)
    return bail("conditional handle", e);
  // part 1: first measurement + test
  if ((e = cudaStreamBeginCaptureToGraph(h->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) !=
      cudaSuccess)
    return bail("capture", e);
  const int64_t l_first = h->launches;
  enqueue_norm(h, k);
  smg::launch_solve_decide(cond, h->loop_ctl.p, norm2, n_norm2, 0, h->stream);
  const int64_t first_launches = h->launches - l_first + 1;
  std::vector<cudaGraphNode_t> deps;
  {
    cudaStreamCaptureStatus st;
    const cudaGraphNode_t* d = nullptr;
    size_t nd = 0;
    if ((e = cudaStreamGetCaptureInfo_v2(h->stream, &st, nullptr, nullptr, &d, &nd)) != cudaSuccess)
      return bail("capture info", e);
    deps.assign(d, d + nd);
  }
  cudaGraph_t same = nullptr;
  if ((e = cudaStreamEndCapture(h->stream, &same)) != cudaSuccess) return bail("end capture", e);
  // part 2: WHILE node
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = cond;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t wnode;
  if ((e = cudaGraphAddNode(&wnode, g, deps.data(), deps.size(), &np)) != cudaSuccess) return bail("while node", e);
  cudaGraph_t body = np.conditional.phGraph_out[0];
  if ((e = cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) !=
      cudaSuccess)
    return bail("body capture", e);
  const int64_t l0 = h->launches;
  vcycle_device(h, 0, h->opt.pre_relax, h->opt.post_relax, k);
  enqueue_norm(h, k);
  smg::launch_solve_decide(cond, h->loop_ctl.p, norm2, n_norm2, 1, h->stream);
  out->launches = h->launches - l0 + 1;  // kernels per loop iteration
  out->first = first_launches;           // kernels of the first measurement
  h->launches = before;
  if ((e = cudaStreamEndCapture(h->stream, &same)) != cudaSuccess) return bail("end body capture", e);
  if ((e = cudaGraphInstantiate(&exec, g, 0)) != cudaSuccess) return bail("instantiate", e);
  cudaGraphDestroy(g);
  out->exec = exec;
  return SMG_OK;
}

// runs the whole loop on the device; SMG_E_UNSUPPORTED: use the host loop instead
int solve_loop_device(smg_handle* h, int k, double tol, int max_iter, double* r_his, int* nh, double* residual) {
  // (ranks that share a device synchronise their exchanges on the host: no graphs there)
  if (h->loop_state < 0 || !h->opt.use_graph || (dist_on(h) && h->dist.host_group) || max_iter < 1 ||
      max_iter > smg::kSolveCtlHis)
    return SMG_E_UNSUPPORTED;
  auto it = h->loop_graphs.find(k);
  if (it == h->loop_graphs.end()) {
    GraphEntry ge;
    const int rc = build_loop_graph(h, k, &ge);
    if (rc != SMG_OK) return rc;
    it = h->loop_graphs.emplace(k, ge).first;
  }
  smg::SolveCtl* hc = h->h_ctl;
  hc->tol = tol;
  hc->max_iter = max_iter;
  hc->n_his = 0;
  hc->nonfinite = 0;
  hc->pad = 0;
  const size_t head = offsetof(smg::SolveCtl, r_his);
  SMG_CUDA(h, cudaMemcpyAsync(h->loop_ctl.p, hc, head, cudaMemcpyHostToDevice, h->stream));
  SMG_CUDA(h, cudaGraphLaunch(it->second.exec, h->stream));
  SMG_CUDA(h, cudaMemcpyAsync(hc, h->loop_ctl.p, head + sizeof(double) * max_iter, cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  const int n = std::max(0, std::min(hc->n_his, max_iter));
  for (int i = 0; i < n; i++) r_his[i] = hc->r_his[i];
  *nh = n;
  *residual = n > 0 ? hc->r_his[n - 1] : 0.0;
  // cycles run: one after every measurement that did not end the loop
  const bool ended_by_test = n > 0 && (!std::isfinite(*residual) || *residual < tol);
  const int cycles = ended_by_test ? n - 1 : n;
  h->launches += it->second.first + static_cast<int64_t>(cycles) * it->second.launches;
  return SMG_OK;
}

// ---- numeric part of precompute (device) -----------------------------------------
int numeric_setup(smg_handle* h) {
  smg::Plan& pl = h->plan;
  const int nlev = static_cast<int>(h->lv.size());
  cudaStream_t st = h->stream;
  // SMG_PRECOMPUTE_TIMING=1: stage times on stderr (synchronises after every stage)
  static const bool timing = std::getenv("SMG_PRECOMPUTE_TIMING") != nullptr;
  double t_stage = now_ms();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const double t = now_ms();
    std::fprintf(stderr, "numeric setup: %-34s %8.3f ms\n", what, t - t_stage);
    t_stage = t;
  };
  // LHS = A(unknown,unknown), Auk = A(unknown,known): value gathers (cpp:167,170)
  LevelDev& L0 = h->lv[0];
  smg::launch_gather_values(h->a_in.p, h->lhs_src.p, L0.a_val.p, pl.LHS.nnz(), st);
  h->launches++;
  if (pl.has_fixed && pl.Auk.nnz() > 0) {
    smg::launch_gather_values(h->a_in.p, h->auk_src.p, h->auk_csc_val.p, pl.Auk.nnz(), st);
    smg::launch_gather_values(h->auk_csc_val.p, h->auk_pos.p, h->auk_val.p, pl.Auk.nnz(), st);
    h->launches += 2;
  }
  lap("value gathers");
  // Galerkin products A_l = (PT * A_{l-1}) * P (cpp:25, :227)
  for (int l = 1; l < nlev; l++) {
    LevelDev& F = h->lv[l - 1];
    LevelDev& C = h->lv[l];
    const smg::LevelPlan& P = pl.lv[l];
    smg::launch_galerkin_t1(P.T1.nnz(), C.t_rowidx.p, C.t_col.p, F.a_colptr.p, F.a_rowidx.p,
                            F.a_val.p, C.pt_colptr.p, C.pt_rowidx.p, C.pt_val.p, C.t_val.p, st);
    smg::launch_galerkin_ac(P.A.nnz(), C.a_rowidx.p, C.a_col.p, C.p_colptr.p, C.p_rowidx.p,
                            C.p_val.p, C.t_colptr.p, C.t_rowidx.p, C.t_val.p, C.a_val.p, st);
    h->launches += 2;
  }
  lap("Galerkin products");
  // coarsest diagonal shift (cpp:31-36, :237-242)
  LevelDev& Lc = h->lv[nlev - 1];
  smg::launch_shift_diag(Lc.a_val.p, Lc.diag_pos.p, Lc.n, 1e-12, st);
  h->launches++;
  // SELL values + A_diag (cpp:38-41, :244-246)
  for (int l = 0; l < nlev; l++) {
    LevelDev& L = h->lv[l];
    smg::launch_fill_sell(L.a_val.p, L.sellA.src.p, L.tmap.p, L.sellA.val.p, L.sellA.valT.p,
                          L.sellA.padded, st);
    smg::launch_extract_diag(L.a_val.p, L.diag_pos.p, L.perm.p, L.diag.p, L.n, st);
    h->launches += 2;
    if (L.patched)
      for (LevelDev::PatchBufs* B : {&L.patch_down, &L.patch_up}) {
        smg::launch_patch_fill(reinterpret_cast<double*>(B->blob.p), B->fill_dst.p, B->fill_src.p, L.a_val.p,
                               B->n_fill, st);
        h->launches++;
      }
  }
  SMG_TRY(check_launch(h, "numeric setup"));
  lap("SELL / diagonal / patch fills");
  // coarse factorisation (cpp:46-48, :253-254): dense Cholesky, explicit inverse
  const int nc = Lc.n;
  if (nc > SMG_MAX_COARSE_ROWS)
    return fail(h, SMG_E_UNSUPPORTED,
                "coarsest level has " + std::to_string(nc) + " rows: the dense inverse supports at most " +
                    std::to_string(SMG_MAX_COARSE_ROWS) + " (add a coarser level)");
  if (nc > 0) {
    SMG_CUDA(h, h->ainv.reserve(static_cast<size_t>(nc) * nc));
    SMG_CUDA(h, cudaMemsetAsync(h->ainv.p, 0, sizeof(double) * nc * nc, st));
    // natural (reference) numbering -> permuted numbering of the coarsest level
    std::vector<int>& iperm = pl.lv[nlev - 1].order.iperm;
    DevBuf<int> d_iperm;
    SMG_CUDA(h, d_iperm.upload(iperm, st));
    smg::launch_csc_to_dense(pl.lv[nlev - 1].A.nnz(), Lc.a_rowidx.p, Lc.a_col.p, Lc.a_val.p,
                             d_iperm.p, h->ainv.p, nc, st);
    h->launches++;
    int lwork1 = 0;
    if (cusolverDnDpotrf_bufferSize(h->cusolver, CUBLAS_FILL_MODE_LOWER, nc, h->ainv.p, nc, &lwork1) !=
        CUSOLVER_STATUS_SUCCESS)
      return fail(h, SMG_E_CUSOLVER, "cusolver bufferSize failed");
    SMG_CUDA(h, h->potrf_work.reserve(static_cast<size_t>(std::max(lwork1, 1))));
    SMG_CUDA(h, h->dev_info.reserve(2));
    if (cusolverDnDpotrf(h->cusolver, CUBLAS_FILL_MODE_LOWER, nc, h->ainv.p, nc,
                         h->potrf_work.p, lwork1, h->dev_info.p) != CUSOLVER_STATUS_SUCCESS)
      return fail(h, SMG_E_CUSOLVER, "cusolverDnDpotrf failed");
    lap("coarse: dense assembly + potrf");
    // explicit inverse from the Cholesky factor, A^-1 = L^-T L^-1, as two level-3 BLAS calls
    // (X = L^-1 by a triangular solve against the identity, then X^T X): cusolverDnDpotri does
    // the same arithmetic at a fraction of the speed (15 ms at n = 4098 on B200), and this runs
    // once per time step of the mean-curvature flow
    SMG_CUDA(h, h->linv.reserve(static_cast<size_t>(nc) * nc));
    SMG_CUDA(h, cudaMemsetAsync(h->linv.p, 0, sizeof(double) * nc * nc, st));
    smg::launch_set_identity(h->linv.p, nc, st);
    h->launches++;
    const double one = 1.0, zero = 0.0;
    if (cublasDtrsm(h->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, nc, nc, &one,
                    h->ainv.p, nc, h->linv.p, nc) != CUBLAS_STATUS_SUCCESS ||
        cublasDsyrk(h->cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, nc, nc, &one, h->linv.p, nc, &zero, h->ainv.p,
                    nc) != CUBLAS_STATUS_SUCCESS)
      return fail(h, SMG_E_CUSOLVER, "cublas trsm / syrk (coarse inverse) failed");
    int info1 = 0;
    SMG_CUDA(h, cudaMemcpyAsync(&info1, h->dev_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SMG_CUDA(h, cudaStreamSynchronize(st));
    if (info1 != 0)
      return fail(h, SMG_E_CUSOLVER,
                  "coarsest matrix is not positive definite (potrf info " + std::to_string(info1) + ")");
    lap("coarse: inverse (trsm + syrk)");
    // keep only the packed lower tiles (half the bytes, contiguous 32 KB blocks)
    SMG_CUDA(h, h->ainv_tiles.reserve(smg::dense_sym_tiles_doubles(nc)));
    smg::launch_pack_sym_tiles(h->ainv.p, h->ainv_tiles.p, nc, st);
    h->launches++;
    SMG_TRY(check_launch(h, "coarse inverse"));
  }
  SMG_CUDA(h, cudaStreamSynchronize(st));
  lap("coarse: tile packing");
  return SMG_OK;
}

// Patch smoother: levels small enough to be latency-bound (multicolour mode).  A level is
// patched when both of its launches (down: pre-smoothing + residual + restriction; up:
// prolongation + post-smoothing) can be laid out; otherwise it keeps one kernel per phase.
int upload_patches(smg_handle* h, int k_cols) {
  smg::Plan& pl = h->plan;
  h->patch_kcols = k_cols;
  for (auto& L : h->lv) L.patched = false;
  const int nlev = static_cast<int>(pl.lv.size());
  cudaStream_t st = h->stream;
  if (h->opt.smoother != SMG_SMOOTHER_MULTICOLOUR || h->opt.patch_rows < 0) return SMG_OK;
  int max_rows = 70000;
  if (const char* e = std::getenv("SMG_PATCH_MAX_ROWS")) max_rows = std::atoi(e);
  int dev_smem = 0, nsm = 0;
  cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  if (dev_smem <= 0 || nsm <= 0) return SMG_OK;
  // One patch per SM (measured on B200, profiles/r2_patch_stages.md: a patch launch costs
  // ~9 us almost independently of the patch size -- dependent chains of shared-memory loads,
  // FP64 adds, a division and a barrier per colour phase -- so more, smaller patches only
  // add redundant halo work and, past one wave, a second round of that latency).
  int budget = dev_smem - 1024;
  if (const char* e = std::getenv("SMG_PATCH_SMEM_KB")) budget = std::min(dev_smem - 1024, std::atoi(e) * 1024);
  for (int l = 0; l + 1 < nlev; l++) {
    const smg::LevelPlan& P = pl.lv[l];
    LevelDev& L = h->lv[l];
    if (P.n <= 0 || P.n > max_rows || P.layout == smg::LAYOUT_PARTITIONED) continue;
    // One patch per SM (more when the patches do not fit shared memory otherwise).
    const bool fixed = h->opt.patch_rows > 0;
    int target = fixed ? h->opt.patch_rows : std::max(96, (P.n + nsm - 1) / nsm);
    if (const char* e = std::getenv("SMG_PATCH_COUNT")) target = (P.n + std::max(1, std::atoi(e)) - 1) / std::max(1, std::atoi(e));
    smg::PatchSet down, up;
    std::string why;
    if (!smg::build_patches(pl, l, smg::PATCH_DOWN, h->opt.pre_relax, target, budget, k_cols, &down, &why) ||
        !smg::build_patches(pl, l, smg::PATCH_UP, h->opt.post_relax, target, budget, k_cols, &up, &why))
      continue;
    if (!fixed && !std::getenv("SMG_PATCH_FORCE")) {
      // Patch only where it pays.  Measured on B200 (profiles/r2_patch_stages.md): a wave of
      // patches costs ~6 us plus ~0.6 us per colour phase and per 256 rows a patch updates in it
      // (more for the wide rows of decimated-mesh Galerkin operators: ~sqrt(mean width / 6));
      // the phase-by-phase kernels cost ~2.7 us per dependent launch.  Levels with many colours
      // grow halos that need several waves of small patches (hilbert_cube's 13 790-row level:
      // 11 colours, 8 + 4 waves, 290 us against 110 us phase by phase): those stay phase by phase.
      auto cost = [&](const smg::PatchSet& ps) {
        const double waves = std::ceil(static_cast<double>(ps.n_patches) / nsm);
        const double w = static_cast<double>(ps.sum_entries) / std::max<int64_t>(1, ps.sum_b) / 6.0;  // mean row width
        const double wide = w <= 1.0 ? 1.0 : std::sqrt(w);
        return waves * (6.0 + 0.6 * ps.max_passes * wide * std::max(1, k_cols / 2 + k_cols % 2));
      };
      const int C = P.n_phases;
      const double phase_cost = 2.7 * ((h->opt.pre_relax + h->opt.post_relax) * C + 3);
      if (cost(down) + cost(up) > 0.85 * phase_cost) continue;
    }
    if (std::getenv("SMG_PATCH_VERIFY")) {
      for (const smg::PatchSet* ps : {&down, &up}) {
        const std::string err = smg::verify_patches(pl, *ps);
        if (!err.empty()) return fail(h, SMG_E_INTERNAL, "patch plan of level " + std::to_string(l) + ": " + err);
      }
    }
    auto put = [&](const smg::PatchSet& ps, LevelDev::PatchBufs& B) -> int {
      SMG_CUDA(h, B.blob.upload(ps.blob, st));
      SMG_CUDA(h, B.off.upload(ps.off, st));
      SMG_CUDA(h, B.fill_dst.upload(ps.fill_dst, st));
      SMG_CUDA(h, B.fill_src.upload(ps.fill_src, st));
      B.n_fill = static_cast<int64_t>(ps.fill_dst.size());
      B.iters = ps.iters;
      B.view.n_patches = ps.n_patches;
      B.view.blob = B.blob.p;
      B.view.off = B.off.p;
      B.view.max_blob_bytes = ps.max_blob_bytes;
      B.view.max_vec_doubles = ps.max_vec_doubles;
      B.view.max_active = ps.max_active;
      return SMG_OK;
    };
    SMG_TRY(put(down, L.patch_down));
    SMG_TRY(put(up, L.patch_up));
    SMG_CUDA(h, cudaStreamSynchronize(st));  // the host vectors above go out of scope
    L.patched = true;
  }
  return SMG_OK;
}

// level l runs its relax calls as patch launches for this V-cycle shape
bool use_patches(const smg_handle* h, int l, int pre, int post, int k) {
  const LevelDev& L = h->lv[l];
  if (!L.patched || pre != L.patch_down.iters || post != L.patch_up.iters) return false;
  const int kk = std::min(k, smg::kMaxK);
  int dev_smem = 0;
  cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  return static_cast<int>(smg::patch_smem_bytes(L.patch_down.view, kk)) + 64 <= dev_smem &&
         static_cast<int>(smg::patch_smem_bytes(L.patch_up.view, kk)) + 64 <= dev_smem;
}

int upload_plan(smg_handle* h) {
  smg::Plan& pl = h->plan;
  const int nlev = static_cast<int>(pl.lv.size());
  cudaStream_t st = h->stream;
  drop_graphs(h);
  h->lv.clear();
  h->lv.resize(nlev);
  h->kcap = 0;
  for (int l = 0; l < nlev; l++) {
    smg::LevelPlan& P = pl.lv[l];
    LevelDev& L = h->lv[l];
    L.n = P.n;
    SMG_CUDA(h, L.a_colptr.upload(P.A.colptr, st));
    SMG_CUDA(h, L.a_rowidx.upload(P.A.rowidx, st));
    SMG_CUDA(h, L.a_col.upload(P.a_col, st));
    SMG_CUDA(h, L.tmap.upload(P.tmap, st));
    SMG_CUDA(h, L.diag_pos.upload(P.diag_pos, st));
    SMG_CUDA(h, L.a_val.alloc(static_cast<size_t>(P.A.nnz())));
    SMG_TRY(upload_sell(h, P.sellA, &L.sellA, true));
    SMG_CUDA(h, L.diag.alloc(static_cast<size_t>(P.n)));
    SMG_CUDA(h, L.perm.upload(P.order.perm, st));
    L.phase_ptr = P.order.phase_ptr;
    if (l >= 1) {
      SMG_TRY(upload_sell(h, P.sellP, &L.sellP, false));
      SMG_TRY(upload_sell(h, P.sellPT, &L.sellPT, false));
      SMG_CUDA(h, L.p_colptr.upload(P.P.colptr, st));
      SMG_CUDA(h, L.p_rowidx.upload(P.P.rowidx, st));
      SMG_CUDA(h, L.p_val.upload(P.P.val, st));
      SMG_CUDA(h, L.pt_colptr.upload(P.PT.colptr, st));
      SMG_CUDA(h, L.pt_rowidx.upload(P.PT.rowidx, st));
      SMG_CUDA(h, L.pt_val.upload(P.PT.val, st));
      SMG_CUDA(h, L.t_colptr.upload(P.T1.colptr, st));
      SMG_CUDA(h, L.t_rowidx.upload(P.T1.rowidx, st));
      SMG_CUDA(h, L.t_col.upload(P.t1_col, st));
      SMG_CUDA(h, L.t_val.alloc(static_cast<size_t>(P.T1.nnz())));
      // values of the transfer operators are static: fill the SELL arrays now.
      // sellP was built from PT's CSC, sellPT from P's CSC (plan.cpp).
      smg::launch_fill_sell(L.pt_val.p, L.sellP.src.p, nullptr, L.sellP.val.p, nullptr,
                            L.sellP.padded, st);
      smg::launch_fill_sell(L.p_val.p, L.sellPT.src.p, nullptr, L.sellPT.val.p, nullptr,
                            L.sellPT.padded, st);
      h->launches += 2;
    }
  }
  SMG_TRY(upload_patches(h, 1));
  SMG_CUDA(h, h->lhs_src.upload(pl.lhs_src, st));
  // permuted unknown row -> caller index
  const std::vector<int>& perm0 = pl.lv[0].order.perm;
  const int nu = pl.lv[0].n;
  std::vector<int> g(nu);
  for (int r = 0; r < nu; r++) g[r] = pl.unknown[perm0[r]];
  SMG_CUDA(h, h->g.upload(g, st));
  h->auk_ptr.release();
  h->n_known_distinct = 0;
  if (pl.has_fixed) {
    const Csc& K = pl.Auk;
    const int nk = K.cols;
    SMG_CUDA(h, h->auk_src.upload(pl.auk_src, st));
    SMG_CUDA(h, h->auk_csc_val.alloc(static_cast<size_t>(K.nnz())));
    SMG_CUDA(h, h->auk_val.alloc(static_cast<size_t>(K.nnz())));
    // Auk by row (ascending known column inside a row), rows in permuted order
    std::vector<int> rp(static_cast<size_t>(nu) + 1, 0), rq(K.nnz()), rpos(K.nnz());
    {
      std::vector<int> cnt(static_cast<size_t>(nu) + 1, 0);
      for (int p = 0; p < K.nnz(); p++) cnt[K.rowidx[p] + 1]++;
      for (int i = 0; i < nu; i++) cnt[i + 1] += cnt[i];
      std::vector<int> q0(K.nnz()), pos0(K.nnz());
      std::vector<int> nx(cnt.begin(), cnt.end() - 1);
      for (int c = 0; c < nk; c++)
        for (int p = K.colptr[c]; p < K.colptr[c + 1]; p++) {
          const int d = nx[K.rowidx[p]]++;
          q0[d] = c;
          pos0[d] = p;
        }
      int o = 0;
      for (int r = 0; r < nu; r++) {
        const int row = perm0[r];
        rp[r] = o;
        for (int t = cnt[row]; t < cnt[row + 1]; t++, o++) {
          rq[o] = q0[t];
          rpos[o] = pos0[t];
        }
      }
      rp[nu] = o;
    }
    SMG_CUDA(h, h->auk_ptr.upload(rp, st));
    SMG_CUDA(h, h->auk_q.upload(rq, st));
    SMG_CUDA(h, h->auk_pos.upload(rpos, st));
    // z(known) = known_val via igl::slice_into (cpp:355): sequential writes, the last
    // occurrence of a repeated index wins
    std::vector<int> last(static_cast<size_t>(pl.n), -1);
    for (int i = 0; i < nk; i++) last[pl.known[i]] = i;
    std::vector<int> kidx, ksrc;
    for (int i = 0; i < pl.n; i++)
      if (last[i] >= 0) {
        kidx.push_back(i);
        ksrc.push_back(last[i]);
      }
    h->n_known_distinct = static_cast<int>(kidx.size());
    SMG_CUDA(h, h->kidx.upload(kidx, st));
    SMG_CUDA(h, h->ksrc.upload(ksrc, st));
  }
  SMG_TRY(check_launch(h, "upload plan"));
  SMG_CUDA(h, cudaStreamSynchronize(st));  // host vectors above go out of scope
  return upload_dist(h);
}

void drop_graphs_if_device(smg_handle* h) {
  if (!h->plan_only && h->device >= 0) {
    cudaSetDevice(h->device);
    tls_stream = h->stream;
    tls_async = h->async_alloc;
    drop_graphs(h);
  }
}

int check_ready(const smg_handle* h, bool need_device) {
  if (!h) return SMG_E_INVALID;
  if (!h->have_plan) return fail(const_cast<smg_handle*>(h), SMG_E_STATE, "smg_precompute has not been called");
  if (need_device && h->plan_only)
    return fail(const_cast<smg_handle*>(h), SMG_E_STATE, "plan-only handle (SMG_DEVICE_NONE) cannot compute");
  if (need_device && h->dist.failed)
    return fail(const_cast<smg_handle*>(h), SMG_E_INTERNAL, "a halo exchange timed out earlier on this handle");
  return SMG_OK;
}

int set_device(smg_handle* h) {
  if (h->plan_only) return SMG_OK;
  SMG_CUDA(h, cudaSetDevice(h->device));
  tls_stream = h->stream;
  tls_async = h->async_alloc;
  return SMG_OK;
}

// an exchange wait that timed out (a peer rank died or never made the matching call)
int check_exchange(smg_handle* h) {
  if (!dist_on(h)) return SMG_OK;
  if (h->dist.failed) return fail(h, SMG_E_INTERNAL, "halo exchange: a peer rank did not reach the host rendezvous");
  int* flag = reinterpret_cast<int*>(h->h_norm + 48);
  *flag = 0;
  SMG_TRY(drain_if_shared(h));
  SMG_CUDA(h, cudaMemcpyAsync(flag, h->dist.ctrl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*flag) {
    h->dist.failed = true;
    return fail(h, SMG_E_INTERNAL, "halo exchange: a wait for a peer rank timed out");
  }
  return SMG_OK;
}

// host (reference numbering, n x k) -> level work vector (permuted numbering)
int stage_in(smg_handle* h, int l, const double* host, DevBuf<double>& stage, double* dst, int k) {
  LevelDev& L = h->lv[l];
  const size_t cnt = static_cast<size_t>(L.n) * k;
  SMG_CUDA(h, stage.reserve(cnt));
  if (cnt == 0) return SMG_OK;
  SMG_CUDA(h, cudaMemcpyAsync(stage.p, host, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  smg::launch_permute_in(stage.p, L.perm.p, dst, L.n, k, h->stream);
  h->launches++;
  return SMG_OK;
}

int stage_out(smg_handle* h, int l, const double* src, DevBuf<double>& stage, double* host, int k) {
  LevelDev& L = h->lv[l];
  const size_t cnt = static_cast<size_t>(L.n) * k;
  SMG_CUDA(h, stage.reserve(cnt));
  if (cnt == 0) return SMG_OK;
  smg::launch_permute_out(src, L.perm.p, stage.p, L.n, k, h->stream);
  h->launches++;
  SMG_TRY(drain_if_shared(h));
  SMG_CUDA(h, cudaMemcpyAsync(host, stage.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  SMG_TRY(check_exchange(h));
  return check_launch(h, "stage_out");
}

int valid_level(smg_handle* h, int lv, bool need_coarser) {
  const int nlev = static_cast<int>(h->lv.size());
  if (lv < 0 || lv >= nlev || (need_coarser && lv + 1 >= nlev))
    return fail(h, SMG_E_INVALID, "level out of range");
  return SMG_OK;
}

// multi-GPU: a level vector whose rows were computed by their owners only -> complete on
// every rank (collective)
void complete_rows(smg_handle* h, int l, double* vec, int k) {
  LevelDev& L = h->lv[l];
  if (dist_on(h) && L.layout == smg::LAYOUT_PARTITIONED) exchange(h, L.x_gather, vec, L.n, k);
}

// min_quad_with_fixed_mg_solve on device pointers
int solve_core(smg_handle* h, const double* d_RHS, const double* d_kv, const double* d_z0, int k,
               double tol, int max_iter, double* d_z, double* r_his, int* n_his, int* converged) {
  smg::Plan& pl = h->plan;
  LevelDev& L0 = h->lv[0];
  const int nu = L0.n;
  const int nk = pl.has_fixed ? static_cast<int>(pl.known.size()) : 0;
  SMG_TRY(ensure_k(h, k));
  // z_unknown = z0(unknown); RHS_unknown = RHS(unknown) - Auk * known_val  (cpp:310-318)
  const bool with_auk = pl.has_fixed && nk > 0 && pl.Auk.nnz() > 0;
  if (with_auk && !d_kv) return fail(h, SMG_E_INVALID, "known_val is required");
  smg::launch_gather_system(d_RHS, d_z0, d_kv, pl.n, nk, h->g.p, with_auk ? h->auk_ptr.p : nullptr,
                            h->auk_q.p, h->auk_val.p, L0.b.p, L0.u.p, nu, k, h->stream);
  h->launches++;
  double residual = 0.0;
  int nh = 0;
  const int rc_loop = solve_loop_device(h, k, tol, max_iter, r_his, &nh, &residual);
  h->last_solve_on_device = rc_loop == SMG_OK ? 1 : 0;
  if (rc_loop == SMG_OK) {
    if (h->opt.verbose)
      for (int i = 0; i < nh; i++) std::printf("%.17g\n", r_his[i]);
    if (nh > 0 && !std::isfinite(residual)) {
      *n_his = nh;
      *converged = 0;
      return fail(h, SMG_E_NONFINITE, "residual is not finite");
    }
  } else if (rc_loop != SMG_E_UNSUPPORTED) {
    return rc_loop;
  } else {
    for (int iter = 0; iter < max_iter; iter++) {  // cpp:330-347 / :108-125
      SMG_TRY(residual_norm_device(h, 0, L0.b.p, L0.u.p, k, &residual));
      r_his[nh++] = residual;
      if (h->opt.verbose) std::printf("%.17g\n", residual);
      if (!std::isfinite(residual)) {
        *n_his = nh;
        *converged = 0;
        return fail(h, SMG_E_NONFINITE, "residual is not finite");
      }
      if (residual < tol) break;
      SMG_TRY(vcycle_run(h, 0, h->opt.pre_relax, h->opt.post_relax, k));
    }
  }
  if (h->opt.verbose) std::printf("residual norm: %.17g\n", residual);
  // z(unknown) = z_unknown ; z(known) = known_val   (cpp:353-355)
  complete_rows(h, 0, L0.u.p, k);  // multi-GPU: every rank returns the whole solution
  smg::launch_scatter_solution(L0.u.p, h->g.p, d_z, pl.n, nu, k, h->stream);
  h->launches++;
  if (pl.has_fixed && h->n_known_distinct > 0) {
    if (!d_kv) return fail(h, SMG_E_INVALID, "known_val is required");
    smg::launch_scatter_known(d_kv, h->kidx.p, h->ksrc.p, d_z, pl.n, nk, h->n_known_distinct, k,
                              h->stream);
    h->launches++;
  }
  SMG_TRY(check_launch(h, "solve"));
  SMG_TRY(check_exchange(h));
  *n_his = nh;
  *converged = residual > tol ? 0 : 1;  // cpp:357-360 (stale residual, by design)
  return SMG_OK;
}

Csc make_csc(int rows, int cols, const int* colptr, const int* rowidx, const double* val) {
  Csc m;
  m.rows = rows;
  m.cols = cols;
  m.colptr.assign(colptr, colptr + cols + 1);
  const int nnz = m.colptr[cols];
  m.rowidx.assign(rowidx, rowidx + nnz);
  if (val) m.val.assign(val, val + nnz);
  return m;
}

bool csc_valid(int rows, int cols, const int* colptr, const int* rowidx, bool sorted_required) {
  if (rows < 0 || cols < 0 || !colptr || colptr[0] != 0) return false;
  for (int j = 0; j < cols; j++)
    if (colptr[j + 1] < colptr[j]) return false;
  if (colptr[cols] > 0 && !rowidx) return false;
  for (int j = 0; j < cols; j++)
    for (int p = colptr[j]; p < colptr[j + 1]; p++) {
      if (rowidx[p] < 0 || rowidx[p] >= rows) return false;
      if (sorted_required && p > colptr[j] && rowidx[p] <= rowidx[p - 1]) return false;
    }
  return true;
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

void smg_default_options(smg_options* opt) {
  if (!opt) return;
  std::memset(opt, 0, sizeof(*opt));
  opt->pre_relax = 2;
  opt->post_relax = 2;
  opt->smoother = SMG_SMOOTHER_MULTICOLOUR;
  opt->device = SMG_DEVICE_CURRENT;
  opt->use_graph = 1;
  opt->verbose = 0;
  opt->locality_reorder = 1;
  opt->sigma = 256;
  opt->patch_rows = 0;  // automatic
}

int smg_version(void) { return SMG_VERSION; }

const char* smg_status_string(int status) {
  switch (status) {
    case SMG_OK: return "ok";
    case SMG_E_INVALID: return "invalid argument";
    case SMG_E_CUDA: return "CUDA error or no device";
    case SMG_E_NLEVELS: return "at least 2 multigrid levels are required";
    case SMG_E_NONFINITE: return "non-finite residual";
    case SMG_E_STATE: return "call order violated";
    case SMG_E_CUSOLVER: return "coarse factorisation failed";
    case SMG_E_NCCL: return "NCCL error";
    case SMG_E_NOT_SYMMETRIC: return "matrix pattern is not symmetric";
    case SMG_E_UNSUPPORTED: return "unsupported";
    case SMG_E_INTERNAL: return "internal error (device-side wait timed out)";
  }
  return "unknown status";
}

const char* smg_last_error(const smg_handle* h) { return h ? h->err.c_str() : "null handle"; }

int smg_create(smg_handle** out, const smg_options* opt) {
  if (!out) return SMG_E_INVALID;
  *out = nullptr;
  smg_handle* h = new smg_handle();
  if (opt) h->opt = *opt;
  else smg_default_options(&h->opt);
  if (h->opt.smoother != SMG_SMOOTHER_WAVEFRONT && h->opt.smoother != SMG_SMOOTHER_MULTICOLOUR) {
    delete h;
    return SMG_E_INVALID;
  }
  if (h->opt.device == SMG_DEVICE_NONE) {
    h->plan_only = true;
    *out = h;
    return SMG_OK;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    delete h;
    return SMG_E_CUDA;  // no CPU fallback, by design
  }
  int dev = h->opt.device;
  if (dev == SMG_DEVICE_CURRENT) {
    if (cudaGetDevice(&dev) != cudaSuccess) {
      delete h;
      return SMG_E_CUDA;
    }
  }
  if (dev < 0 || dev >= ndev || cudaSetDevice(dev) != cudaSuccess) {
    delete h;
    return SMG_E_CUDA;
  }
  h->device = dev;
  if (const char* e = std::getenv("SMG_NO_PDL")) smg::set_pdl_enabled(!(e[0] && e[0] != '0'));
  if (const char* e = std::getenv("SMG_GS_ROWS")) smg::set_gs_rows(std::atoi(e));
  if (const char* e = std::getenv("SMG_APPLY2_ROWS")) smg::set_apply2_rows(std::atoi(e));
  if (const char* e = std::getenv("SMG_MULTI_ROWS")) smg::set_multi_rows(std::atoi(e));
  if (const char* e = std::getenv("SMG_NO_PREFETCH")) h->no_prefetch = (e[0] && e[0] != '0');
  if (const char* e = std::getenv("SMG_PATCH_ROWS")) h->opt.patch_rows = std::atoi(e);
  if (const char* e = std::getenv("SMG_HOST_LOOP")) h->loop_state = (e[0] && e[0] != '0') ? -1 : 0;
  if (const char* e = std::getenv("SMG_NO_TMA")) smg::set_tma_enabled(!(e[0] && e[0] != '0'));
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&h->h_norm), 1024 * sizeof(double)) != cudaSuccess) {
    smg_destroy(h);
    return SMG_E_CUDA;
  }
  {  // stream-ordered allocation when the device has memory pools (see DevBuf)
    int pools = 0;
    cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, dev);
    if (const char* e = std::getenv("SMG_NO_ASYNC_ALLOC")) pools = pools && !(e[0] && e[0] != '0');
    if (pools) {
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;  // freed blocks stay in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        h->async_alloc = true;
      }
    }
    cudaGetLastError();
    tls_stream = h->stream;
    tls_async = h->async_alloc;
  }
  if (cusolverDnCreate(&h->cusolver) != CUSOLVER_STATUS_SUCCESS ||
      cusolverDnSetStream(h->cusolver, h->stream) != CUSOLVER_STATUS_SUCCESS ||
      cublasCreate(&h->cublas) != CUBLAS_STATUS_SUCCESS ||
      cublasSetStream(h->cublas, h->stream) != CUBLAS_STATUS_SUCCESS) {
    smg_destroy(h);
    return SMG_E_CUSOLVER;
  }
  if (h->norm_counter.alloc(4) != cudaSuccess ||
      cudaMemsetAsync(h->norm_counter.p, 0, 4 * sizeof(unsigned int), h->stream) != cudaSuccess) {
    smg_destroy(h);
    return SMG_E_CUDA;
  }
  *out = h;
  return SMG_OK;
}

void smg_destroy(smg_handle* h) {
  if (!h) return;
  if (!h->plan_only && h->device >= 0) {
    cudaSetDevice(h->device);
    tls_stream = h->stream;
    tls_async = h->async_alloc && h->stream;
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graphs(h);
    if (h->cusolver) cusolverDnDestroy(h->cusolver);
    if (h->cublas) cublasDestroy(h->cublas);
    if (h->h_norm) cudaFreeHost(h->h_norm);
    if (h->h_ctl) cudaFreeHost(h->h_ctl);
    h->loop_ctl.release();
    h->lv.clear();
    // remaining DevBufs are released by the destructor below, before the stream
    h->a_in.release(); h->lhs_src.release(); h->auk_src.release(); h->g.release();
    h->auk_ptr.release(); h->auk_q.release(); h->auk_pos.release();
    h->auk_csc_val.release(); h->auk_val.release(); h->kidx.release(); h->ksrc.release();
    h->ainv.release(); h->linv.release(); h->ainv_tiles.release(); h->coarse_scratch.release(); h->potrf_work.release(); h->dev_info.release();
    h->st_a.release(); h->st_b.release(); h->st_c.release(); h->st_d.release();
    h->norm_scratch.release(); h->norm_out.release(); h->norm_counter.release(); h->flush.release();
    h->mcf.F.release(); h->mcf.vf_ptr.release(); h->mcf.vf_face.release(); h->mcf.Lval.release();
    h->mcf.dblA.release(); h->mcf.mass.release(); h->mcf.U.release(); h->mcf.rhs.release(); h->mcf.z.release();
    DistCtx& D = h->dist;
    D.ctrl.release(); D.normv.release(); D.x_norm = ExchDev(); D.x_slices = ExchDev();
    for (size_t q = 0; q < D.peer.size(); q++)
      if (D.opened[q] && D.peer[q]) cudaIpcCloseMemHandle(D.peer[q]);
    if (h->stream) cudaStreamSynchronize(h->stream);  // the frees above are stream-ordered
    if (D.comm) cudaFree(D.comm);
    D.comm = nullptr;
    if (h->stream) cudaStreamDestroy(h->stream);
    tls_stream = nullptr;
    tls_async = false;  // anything the destructor below still frees goes through cudaFree
  }
  delete h;
}

int smg_set_hierarchy(smg_handle* h, int n_levels, const int* n_rows, const int* const* P_colptr,
                      const int* const* P_rowidx, const double* const* P_val) {
  if (!h) return SMG_E_INVALID;
  if (n_levels < 2) return fail(h, SMG_E_NLEVELS, "at least 2 multigrid levels are required");
  if (!n_rows || !P_colptr || !P_rowidx || !P_val) return fail(h, SMG_E_INVALID, "null argument");
  std::vector<Csc> P;
  for (int l = 1; l < n_levels; l++) {
    const int rows = n_rows[l - 1], cols = n_rows[l];
    if (!P_colptr[l - 1] || !P_val[l - 1] ||
        !csc_valid(rows, cols, P_colptr[l - 1], P_rowidx[l - 1], true))
      return fail(h, SMG_E_INVALID,
                  "prolongation " + std::to_string(l) + " is not a valid sorted CSC matrix");
    P.push_back(make_csc(rows, cols, P_colptr[l - 1], P_rowidx[l - 1], P_val[l - 1]));
  }
  h->n_rows.assign(n_rows, n_rows + n_levels);
  h->P_full = std::move(P);
  h->have_hierarchy = true;
  h->have_plan = false;
  h->mcf.ready = false;
  return SMG_OK;
}

int smg_precompute(smg_handle* h, int n, const int* A_colptr, const int* A_rowidx,
                   const double* A_val, const int* known, int n_known) {
  if (!h) return SMG_E_INVALID;
  if (!h->have_hierarchy) return fail(h, SMG_E_STATE, "smg_set_hierarchy has not been called");
  if (!A_colptr || !A_val || (n_known > 0 && !known)) return fail(h, SMG_E_INVALID, "null argument");
  if (n != h->n_rows[0]) return fail(h, SMG_E_INVALID, "A does not match the finest level size");
  if (!csc_valid(n, n, A_colptr, A_rowidx, true))
    return fail(h, SMG_E_INVALID, "A is not a valid sorted CSC matrix");
  SMG_TRY(set_device(h));
  h->have_plan = false;
  h->mcf.ready = false;  // sized for the previous matrix
  NvtxRange nvtx_all("smg_precompute");
  const double t0 = now_ms();
  NvtxRange nvtx_plan("smg_precompute: host index planning");
  Csc A = make_csc(n, n, A_colptr, A_rowidx, nullptr);
  smg::PlanOptions po;
  po.smoother = h->opt.smoother;
  po.locality_reorder = h->opt.locality_reorder;
  po.sigma = h->opt.sigma > 0 ? h->opt.sigma : 1;
  po.world = h->dist.world;
  po.dist_levels = h->dist_levels;
  if (h->dist_min_rows > 0) po.dist_min_rows = h->dist_min_rows;
  if (dist_on(h) && !h->plan_only && !h->dist.connected)
    return fail(h, SMG_E_STATE, "smg_dist_connect has not been called");
  const int rc = smg::build_plan(A, known, n_known, h->P_full, po, &h->plan);
  nvtx_plan.end();
  if (rc != SMG_OK) return fail(h, rc, h->plan.error);
  const double t1 = now_ms();
  h->timings[3] = t1 - t0;
  if (h->plan_only) {
    h->have_plan = true;
    return SMG_OK;
  }
  {
    NvtxRange r("smg_precompute: upload + patch layouts");
    SMG_TRY(upload_plan(h));
  }
  NvtxRange nvtx_num("smg_precompute: numeric (Galerkin, diagonals, coarse inverse)");
  const int nnz = A_colptr[n];
  h->a_nnz = static_cast<size_t>(nnz);
  SMG_CUDA(h, h->a_in.reserve(static_cast<size_t>(nnz)));
  SMG_CUDA(h, cudaMemcpyAsync(h->a_in.p, A_val, sizeof(double) * nnz, cudaMemcpyHostToDevice,
                              h->stream));
  SMG_TRY(numeric_setup(h));
  if (dist_on(h)) {
    // No allocation between collective calls (up to kMaxK right-hand sides): growing the
    // memory pool can synchronise the device, which a rank that shares its device with a
    // rank already spinning in an exchange must not do.
    h->have_plan = true;
    SMG_TRY(ensure_k(h, smg::kMaxK));
    const size_t cnt = static_cast<size_t>(h->plan.n) * smg::kMaxK;
    SMG_CUDA(h, h->st_a.reserve(cnt));
    SMG_CUDA(h, h->st_b.reserve(cnt));
    SMG_CUDA(h, h->st_c.reserve(cnt));
    SMG_CUDA(h, h->st_d.reserve(std::max<size_t>(h->plan.known.size(), 1) * smg::kMaxK));
    SMG_CUDA(h, h->norm_scratch.reserve(static_cast<size_t>(smg::residual_norm_blocks(h->lv[0].n))));
    SMG_CUDA(h, h->norm_out.reserve(16));
    SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->timings[4] = now_ms() - t1;
  h->have_plan = true;
  return SMG_OK;
}

int smg_update_values(smg_handle* h, const double* A_val) {
  SMG_TRY(check_ready(h, true));
  if (!A_val) return fail(h, SMG_E_INVALID, "null argument");
  SMG_TRY(set_device(h));
  NvtxRange nvtx_all("smg_update_values");
  const double t0 = now_ms();
  const size_t nnz = h->a_nnz;  // of the last smg_precompute, not the buffer capacity
  SMG_CUDA(h, cudaMemcpyAsync(h->a_in.p, A_val, sizeof(double) * nnz, cudaMemcpyHostToDevice,
                              h->stream));
  SMG_TRY(numeric_setup(h));
  h->timings[3] = 0.0;
  h->timings[4] = now_ms() - t0;
  return SMG_OK;
}

int smg_solve_device(smg_handle* h, const double* d_RHS, const double* d_known_val,
                     const double* d_z0, int k, double tol, int max_iter, double* d_z,
                     double* r_his, int* n_his, int* converged) {
  SMG_TRY(check_ready(h, true));
  if (!d_RHS || !d_z0 || !d_z || !r_his || !n_his || !converged || k < 1 || max_iter < 0)
    return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  const double t0 = now_ms();
  const int rc = solve_core(h, d_RHS, d_known_val, d_z0, k, tol, max_iter, d_z, r_his, n_his,
                            converged);
  if (rc == SMG_OK) SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->timings[0] = h->timings[2] = 0.0;
  h->timings[1] = now_ms() - t0;
  return rc;
}

int smg_solve(smg_handle* h, const double* RHS, const double* known_val, const double* z0, int k,
              double tol, int max_iter, double* z, double* r_his, int* n_his, int* converged) {
  SMG_TRY(check_ready(h, true));
  if (!RHS || !z0 || !z || !r_his || !n_his || !converged || k < 1 || max_iter < 0)
    return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  const smg::Plan& pl = h->plan;
  const size_t cnt = static_cast<size_t>(pl.n) * k;
  const size_t nk = pl.has_fixed ? pl.known.size() : 0;
  if (nk > 0 && !known_val) return fail(h, SMG_E_INVALID, "known_val is required");
  NvtxRange nvtx_all("smg_solve");
  const double t0 = now_ms();
  NvtxRange nvtx_h2d("smg_solve: H2D");
  SMG_CUDA(h, h->st_a.reserve(cnt));
  SMG_CUDA(h, h->st_b.reserve(cnt));
  SMG_CUDA(h, h->st_c.reserve(cnt));
  SMG_CUDA(h, h->st_d.reserve(nk * k));
  if (dist_on(h) && h->dist.slices_ok) {
    // every rank copies its slice of the rows (per column), the exchange completes the vectors
    const int W = h->dist.world, me = h->dist.rank;
    const size_t b = static_cast<size_t>(static_cast<int64_t>(pl.n) * me / W), e = static_cast<size_t>(static_cast<int64_t>(pl.n) * (me + 1) / W);
    for (int q = 0; q < k && e > b; q++) {
      const size_t o = static_cast<size_t>(q) * pl.n + b;
      SMG_CUDA(h, cudaMemcpyAsync(h->st_a.p + o, RHS + o, (e - b) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
      SMG_CUDA(h, cudaMemcpyAsync(h->st_b.p + o, z0 + o, (e - b) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    exchange(h, h->dist.x_slices, h->st_a.p, pl.n, k);
    exchange(h, h->dist.x_slices, h->st_b.p, pl.n, k);
  } else {
    SMG_CUDA(h, cudaMemcpyAsync(h->st_a.p, RHS, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    SMG_CUDA(h, cudaMemcpyAsync(h->st_b.p, z0, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  if (nk > 0)
    SMG_CUDA(h, cudaMemcpyAsync(h->st_d.p, known_val, nk * k * sizeof(double),
                                cudaMemcpyHostToDevice, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  nvtx_h2d.end();
  const double t1 = now_ms();
  NvtxRange nvtx_loop("smg_solve: solve loop");
  const int rc = solve_core(h, h->st_a.p, nk > 0 ? h->st_d.p : nullptr, h->st_b.p, k, tol,
                            max_iter, h->st_c.p, r_his, n_his, converged);
  if (rc != SMG_OK) return rc;
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  nvtx_loop.end();
  const double t2 = now_ms();
  NvtxRange nvtx_d2h("smg_solve: D2H");
  SMG_CUDA(h, cudaMemcpyAsync(z, h->st_c.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  const double t3 = now_ms();
  h->timings[0] = t1 - t0;
  h->timings[1] = t2 - t1;
  h->timings[2] = t3 - t2;
  return SMG_OK;
}

// ---- mean-curvature-flow step ------------------------------------------------------
int smg_mcf_setup(smg_handle* h, int nV, int nF, const int* F, const double* L_val, double delta) {
  SMG_TRY(check_ready(h, true));
  if (!F || !L_val || nV < 1 || nF < 0) return fail(h, SMG_E_INVALID, "bad argument");
  const smg::Plan& pl = h->plan;
  if (pl.has_fixed) return fail(h, SMG_E_STATE, "smg_mcf_* needs the precompute variant without fixed values");
  if (nV != pl.n) return fail(h, SMG_E_INVALID, "nV does not match the precomputed matrix");
  for (int i = 0; i < 3 * nF; i++)
    if (F[i] < 0 || F[i] >= nV) return fail(h, SMG_E_INVALID, "face index out of range");
  SMG_TRY(set_device(h));
  auto& m = h->mcf;
  m.ready = false;
  m.nV = nV;
  m.nF = nF;
  m.delta = delta;
  // incident faces of every vertex in the order setFromTriplets sums them: corner 0 (faces
  // ascending), corner 1, corner 2 (massmatrix_intrinsic.cpp:59-64)
  std::vector<int> ptr(static_cast<size_t>(nV) + 1, 0), faces(static_cast<size_t>(3) * nF);
  for (int i = 0; i < 3 * nF; i++) ptr[static_cast<size_t>(F[i]) + 1]++;
  for (int v = 0; v < nV; v++) ptr[static_cast<size_t>(v) + 1] += ptr[static_cast<size_t>(v)];
  {
    std::vector<int> nx(ptr.begin(), ptr.end() - 1);
    for (int c = 0; c < 3; c++)
      for (int f = 0; f < nF; f++) faces[static_cast<size_t>(nx[static_cast<size_t>(F[f + c * nF])]++)] = f;
  }
  SMG_CUDA(h, m.F.upload(std::vector<int>(F, F + static_cast<size_t>(3) * nF), h->stream));
  SMG_CUDA(h, m.vf_ptr.upload(ptr, h->stream));
  SMG_CUDA(h, m.vf_face.upload(faces, h->stream));
  SMG_CUDA(h, m.Lval.upload(std::vector<double>(L_val, L_val + pl.LHS.nnz()), h->stream));
  SMG_CUDA(h, m.dblA.alloc(static_cast<size_t>(std::max(nF, 1))));
  SMG_CUDA(h, m.mass.alloc(static_cast<size_t>(nV)));
  SMG_CUDA(h, m.U.alloc(static_cast<size_t>(nV) * 3));
  SMG_CUDA(h, m.rhs.alloc(static_cast<size_t>(nV) * 3));
  SMG_CUDA(h, m.z.alloc(static_cast<size_t>(nV) * 3));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  m.ready = true;
  return SMG_OK;
}

int smg_mcf_step(smg_handle* h, const double* U, double tol, int max_iter, double* U_out, double* r_his,
                 int* n_his, int* converged) {
  SMG_TRY(check_ready(h, true));
  auto& m = h->mcf;
  if (!m.ready) return fail(h, SMG_E_STATE, "smg_mcf_setup has not been called");
  if (!U || !U_out || !r_his || !n_his || !converged || max_iter < 0) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  NvtxRange nvtx_all("smg_mcf_step");
  const double t0 = now_ms();
  NvtxRange nvtx_pre("smg_mcf_step: assembly + numeric precompute");
  const size_t cnt = static_cast<size_t>(m.nV) * 3;
  SMG_CUDA(h, cudaMemcpyAsync(m.U.p, U, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  LevelDev& L0 = h->lv[0];
  // LHS = M(U) - delta * L straight into the device copy of the caller's matrix values
  // (free variant: LHS has the caller's pattern and order), RHS = M * U
  smg::launch_mcf_assemble(m.nV, m.nF, m.F.p, m.U.p, m.vf_ptr.p, m.vf_face.p, m.dblA.p, m.mass.p,
                           h->plan.LHS.nnz(), L0.a_rowidx.p, L0.a_col.p, m.delta, m.Lval.p, h->a_in.p, 3,
                           m.rhs.p, h->stream);
  h->launches += 4;
  SMG_TRY(check_launch(h, "mcf assemble"));
  SMG_TRY(numeric_setup(h));  // Galerkin products, diagonals, coarse factorisation: as smg_update_values
  nvtx_pre.end();
  const double t1 = now_ms();
  NvtxRange nvtx_loop("smg_mcf_step: solve loop");
  const int rc = solve_core(h, m.rhs.p, nullptr, m.U.p, 3, tol, max_iter, m.z.p, r_his, n_his, converged);
  if (rc != SMG_OK) return rc;
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  const double t2 = now_ms();
  SMG_TRY(drain_if_shared(h));
  SMG_CUDA(h, cudaMemcpyAsync(U_out, m.z.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->timings[0] = 0.0;
  h->timings[4] = t1 - t0;  // assembly + numeric precompute
  h->timings[1] = t2 - t1;
  h->timings[2] = now_ms() - t2;
  return SMG_OK;
}

// ---- mg_VCycle.h operators -------------------------------------------------------
int smg_vcycle(smg_handle* h, int lv, int pre, int post, const double* B, double* u, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!B || !u || k < 1 || pre < 0 || post < 0) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  LevelDev& L = h->lv[lv];
  SMG_TRY(stage_in(h, lv, B, h->st_a, L.b.p, k));
  SMG_TRY(stage_in(h, lv, u, h->st_b, L.u.p, k));
  SMG_TRY(vcycle_run(h, lv, pre, post, k));
  complete_rows(h, lv, L.u.p, k);
  SMG_TRY(check_exchange(h));
  return stage_out(h, lv, L.u.p, h->st_a, u, k);
}

int smg_relax(smg_handle* h, int lv, int iters, const double* B, double* u, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!B || !u || k < 1 || iters < 0) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  LevelDev& L = h->lv[lv];
  SMG_TRY(stage_in(h, lv, B, h->st_a, L.b.p, k));
  SMG_TRY(stage_in(h, lv, u, h->st_b, L.u.p, k));
  relax_device(h, lv, iters, L.b.p, L.u.p, k);
  complete_rows(h, lv, L.u.p, k);
  return stage_out(h, lv, L.u.p, h->st_a, u, k);
}

int smg_apply_A(smg_handle* h, int lv, const double* u, double* Au, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!u || !Au || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  LevelDev& L = h->lv[lv];
  SMG_TRY(stage_in(h, lv, u, h->st_a, L.u.p, k));
  apply_A_device(h, lv, L.u.p, L.r.p, k);
  complete_rows(h, lv, L.r.p, k);
  return stage_out(h, lv, L.r.p, h->st_a, Au, k);
}

int smg_residual(smg_handle* h, int lv, const double* B, const double* u, double* r, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!B || !u || !r || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  LevelDev& L = h->lv[lv];
  SMG_TRY(stage_in(h, lv, B, h->st_a, L.b.p, k));
  SMG_TRY(stage_in(h, lv, u, h->st_b, L.u.p, k));
  residual_device(h, lv, L.b.p, L.u.p, L.r.p, k);
  complete_rows(h, lv, L.r.p, k);
  return stage_out(h, lv, L.r.p, h->st_a, r, k);
}

int smg_residual_norm(smg_handle* h, int lv, const double* B, const double* u, int k,
                      double* norm) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!B || !u || !norm || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  LevelDev& L = h->lv[lv];
  SMG_TRY(stage_in(h, lv, B, h->st_a, L.b.p, k));
  SMG_TRY(stage_in(h, lv, u, h->st_b, L.u.p, k));
  return residual_norm_device(h, lv, L.b.p, L.u.p, k, norm);
}

int smg_restrict(smg_handle* h, int lv, const double* x, double* Rx, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, true));
  if (!x || !Rx || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  SMG_TRY(stage_in(h, lv, x, h->st_a, h->lv[lv].r.p, k));
  restrict_device(h, lv, h->lv[lv].r.p, h->lv[lv + 1].b.p, k);
  complete_rows(h, lv + 1, h->lv[lv + 1].b.p, k);
  return stage_out(h, lv + 1, h->lv[lv + 1].b.p, h->st_a, Rx, k);
}

int smg_prolong(smg_handle* h, int lv, const double* x, double* Px, int k) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, true));
  if (!x || !Px || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  SMG_TRY(stage_in(h, lv + 1, x, h->st_a, h->lv[lv + 1].u.p, k));
  prolong_device(h, lv, h->lv[lv + 1].u.p, h->lv[lv].r.p, k, false);
  complete_rows(h, lv, h->lv[lv].r.p, k);
  return stage_out(h, lv, h->lv[lv].r.p, h->st_a, Px, k);
}

int smg_coarse_solve(smg_handle* h, const double* B, double* u, int k) {
  SMG_TRY(check_ready(h, true));
  if (!B || !u || k < 1) return fail(h, SMG_E_INVALID, "bad argument");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  const int last = static_cast<int>(h->lv.size()) - 1;
  LevelDev& L = h->lv[last];
  SMG_TRY(stage_in(h, last, B, h->st_a, L.b.p, k));
  SMG_TRY(stage_in(h, last, u, h->st_b, L.u.p, k));
  coarse_solve_device(h, L.b.p, L.u.p, k);
  return stage_out(h, last, L.u.p, h->st_a, u, k);
}

// ---- multi-GPU -------------------------------------------------------------------------
int smg_dist_init(smg_handle* h, int rank, int world, size_t comm_bytes) {
  if (!h) return SMG_E_INVALID;
  if (world < 1 || world > smg::xchg_max_peers() || rank < 0 || rank >= world) return fail(h, SMG_E_INVALID, "bad rank / world");
  DistCtx& D = h->dist;
  if (D.comm || D.connected) return fail(h, SMG_E_STATE, "smg_dist_init was already called");
  D.rank = rank;
  D.world = world;
  h->have_plan = false;
  if (world == 1 || h->plan_only) return SMG_OK;
  SMG_TRY(set_device(h));
  if (comm_bytes == 0) comm_bytes = size_t(256) << 20;
  D.flags_bytes = ((static_cast<size_t>(world) * kFlagStride + 1023) / 1024) * 1024;
  if (comm_bytes < D.flags_bytes + static_cast<size_t>(world) * 2 * 4096)
    return fail(h, SMG_E_INVALID, "comm_bytes too small");
  D.slot_doubles = ((comm_bytes - D.flags_bytes) / sizeof(double) / (2 * static_cast<size_t>(world))) & ~size_t(31);
  D.comm_bytes = comm_bytes;
  smg::preload_kernels();
  if (const char* e = std::getenv("SMG_XCHG_TIMEOUT_MS")) smg::set_xchg_timeout_ms(std::atoll(e));
  SMG_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&D.comm), comm_bytes));
  SMG_CUDA(h, cudaMemsetAsync(D.comm, 0, comm_bytes, h->stream));
  SMG_CUDA(h, D.ctrl.upload(std::vector<int>(static_cast<size_t>(smg::xchg_ctrl_ints()), 0), h->stream));
  SMG_CUDA(h, D.normv.upload(std::vector<double>(static_cast<size_t>(world) * 16, 0.0), h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMG_OK;
}

int smg_dist_handle_bytes(void) { return static_cast<int>(sizeof(DistBlob)); }

int smg_dist_get_handle(smg_handle* h, void* blob) {
  if (!h || !blob) return SMG_E_INVALID;
  DistCtx& D = h->dist;
  if (!D.comm) return fail(h, SMG_E_STATE, "smg_dist_init (world > 1, with a device) has not been called");
  SMG_TRY(set_device(h));
  DistBlob b;
  std::memset(&b, 0, sizeof(b));
  SMG_CUDA(h, cudaIpcGetMemHandle(&b.ipc, D.comm));
  b.pid = static_cast<long long>(getpid());
  b.ptr = reinterpret_cast<unsigned long long>(D.comm);
  b.bytes = static_cast<long long>(D.comm_bytes);
  b.device = h->device;
  b.rank = D.rank;
  cudaDeviceProp prop;
  SMG_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
  std::memcpy(b.uuid, prop.uuid.bytes, 16);
  std::memcpy(blob, &b, sizeof(b));
  return SMG_OK;
}

int smg_dist_connect(smg_handle* h, const void* all_blobs) {
  if (!h || !all_blobs) return SMG_E_INVALID;
  DistCtx& D = h->dist;
  if (D.world == 1) return SMG_OK;
  if (!D.comm) return fail(h, SMG_E_STATE, "smg_dist_init has not been called");
  if (D.connected) return fail(h, SMG_E_STATE, "already connected");
  SMG_TRY(set_device(h));
  D.peer.assign(static_cast<size_t>(D.world), nullptr);
  D.opened.assign(static_cast<size_t>(D.world), 0);
  const long long me = static_cast<long long>(getpid());
  DistBlob mine;
  std::memcpy(&mine, static_cast<const char*>(all_blobs) + static_cast<size_t>(D.rank) * sizeof(DistBlob), sizeof(mine));
  for (int q = 0; q < D.world; q++) {
    DistBlob b;
    std::memcpy(&b, static_cast<const char*>(all_blobs) + static_cast<size_t>(q) * sizeof(DistBlob), sizeof(b));
    if (b.rank != q || b.bytes != static_cast<long long>(D.comm_bytes))
      return fail(h, SMG_E_INVALID, "blob " + std::to_string(q) + " does not match (rank order / comm_bytes)");
    if (q != D.rank && std::memcmp(b.uuid, mine.uuid, 16) == 0) D.shared_device = true;
    if (q == D.rank) {
      D.peer[q] = D.comm;
    } else if (b.pid == me) {  // a rank in this process: plain peer access
      if (b.device != h->device) {
        int can = 0;
        SMG_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, b.device));
        if (!can) return fail(h, SMG_E_UNSUPPORTED, "no peer access between the devices of two ranks");
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(h, SMG_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      D.peer[q] = reinterpret_cast<char*>(b.ptr);
    } else {
      void* p = nullptr;
      SMG_CUDA(h, cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
      D.peer[q] = static_cast<char*>(p);
      D.opened[q] = 1;
    }
  }
  // every rank in this process and some device shared: host-synchronised exchanges
  bool same_process = true;
  unsigned long long key = 0;
  for (int q = 0; q < D.world; q++) {
    DistBlob b;
    std::memcpy(&b, static_cast<const char*>(all_blobs) + static_cast<size_t>(q) * sizeof(DistBlob), sizeof(b));
    same_process = same_process && b.pid == me;
    if (q == 0) key = b.ptr;
  }
  if (same_process && D.shared_device) {
    std::lock_guard<std::mutex> lk(g_groups_mutex);
    std::shared_ptr<HostGroup> g = g_groups[key].lock();
    if (!g) {
      g = std::make_shared<HostGroup>();
      g->size = D.world;
      g_groups[key] = g;
    }
    D.host_group = g;
    h->opt.use_graph = 0;
  }
  D.connected = true;
  return SMG_OK;
}

// File rendezvous.  A file is {RvHeader, payload}; a reader accepts it only when the header's
// process is alive, so the leftovers of an earlier run with the same directory and tag (dead
// pids) are never taken for a peer's blob.  After a rank has read every blob it publishes an
// acknowledgement that names the (pid, seq) of the blob it answers to; once every peer has
// acknowledged, a rank removes its own blob (the small acknowledgement file stays; a later
// round replaces it, and (pid, seq) tells a reader which round it belongs to).
namespace {
struct RvHeader {
  unsigned long long magic;
  long long pid;
  long long seq;  // rendezvous calls made by this process so far
};
constexpr unsigned long long kRvMagic = 0x534d47525a563032ull;  // "SMGRZV02"
bool pid_alive(long long pid) {
  if (pid <= 0) return false;
  return ::kill(static_cast<pid_t>(pid), 0) == 0 || errno == EPERM;
}
bool rv_publish(const std::string& path, const RvHeader& hd, const void* payload, size_t bytes) {
  const std::string tmp = path + ".tmp." + std::to_string(hd.pid);
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f) return false;
  size_t w = std::fwrite(&hd, 1, sizeof(hd), f);
  if (bytes > 0) w += std::fwrite(payload, 1, bytes, f);
  if (std::fclose(f) != 0 || w != sizeof(hd) + bytes) return false;
  return std::rename(tmp.c_str(), path.c_str()) == 0;  // atomic within a file system
}
bool rv_read(const std::string& path, RvHeader* hd, void* payload, size_t bytes) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<char> buf(sizeof(RvHeader) + bytes + 1);
  const size_t r = std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if (r != sizeof(RvHeader) + bytes) return false;
  std::memcpy(hd, buf.data(), sizeof(RvHeader));
  if (hd->magic != kRvMagic || !pid_alive(hd->pid)) return false;  // stale: its writer is gone
  if (bytes > 0) std::memcpy(payload, buf.data() + sizeof(RvHeader), bytes);
  return true;
}
std::atomic<long long> g_rv_seq{0};
}  // namespace

int smg_rendezvous_files(const char* dir, const char* tag, int rank, int world, const void* mine,
                         size_t bytes, void* all, int timeout_ms) {
  if (!dir || !tag || !mine || !all || world < 1 || rank < 0 || rank >= world || bytes == 0)
    return SMG_E_INVALID;
  const std::string base = std::string(dir) + "/" + tag + ".";
  RvHeader me;
  me.magic = kRvMagic;
  me.pid = static_cast<long long>(getpid());
  me.seq = ++g_rv_seq;
  const std::string my_blob = base + std::to_string(rank), my_ack = my_blob + ".ack";
  if (!rv_publish(my_blob, me, mine, bytes)) return SMG_E_INVALID;
  const double t0 = now_ms();
  std::vector<RvHeader> hd(static_cast<size_t>(world));
  std::vector<char> have(static_cast<size_t>(world), 0);
  auto wait_all = [&](auto&& try_one) {
    int missing = world;
    std::fill(have.begin(), have.end(), 0);
    while (missing > 0) {
      for (int q = 0; q < world; q++)
        if (!have[q] && try_one(q)) {
          have[q] = 1;
          missing--;
        }
      if (missing > 0) {
        if (timeout_ms >= 0 && now_ms() - t0 > timeout_ms) return false;
        usleep(2000);
      }
    }
    return true;
  };
  // round 1: every rank's blob, written by a live process
  if (!wait_all([&](int q) {
        return rv_read(base + std::to_string(q), &hd[q], static_cast<char*>(all) + static_cast<size_t>(q) * bytes,
                       bytes);
      }))
    return SMG_E_INTERNAL;
  // round 2: acknowledgements.  An ack carries, per rank, the (pid, seq) of the blob its writer
  // read; a rank is done when every peer has read THIS round's blob of it.
  std::vector<long long> seen(static_cast<size_t>(world) * 2);
  for (int q = 0; q < world; q++) {
    seen[2 * q] = hd[q].pid;
    seen[2 * q + 1] = hd[q].seq;
  }
  if (!rv_publish(my_ack, me, seen.data(), seen.size() * sizeof(long long))) return SMG_E_INVALID;
  std::vector<long long> theirs(seen.size());
  if (!wait_all([&](int q) {
        RvHeader ah;
        if (!rv_read(base + std::to_string(q) + ".ack", &ah, theirs.data(), theirs.size() * sizeof(long long)))
          return false;
        return ah.pid == hd[q].pid && ah.seq == hd[q].seq && theirs[2 * rank] == me.pid &&
               theirs[2 * rank + 1] == me.seq;
      }))
    return SMG_E_INTERNAL;
  std::remove(my_blob.c_str());  // every peer has it; the ack stays until this rank's next round
  return SMG_OK;
}

int smg_dist_connect_files(smg_handle* h, const char* dir, const char* tag, int timeout_ms) {
  if (!h) return SMG_E_INVALID;
  if (h->dist.world == 1) return SMG_OK;
  DistBlob mine;
  SMG_TRY(smg_dist_get_handle(h, &mine));
  std::vector<DistBlob> all(static_cast<size_t>(h->dist.world));
  const int rc = smg_rendezvous_files(dir, tag, h->dist.rank, h->dist.world, &mine, sizeof(mine), all.data(),
                                      timeout_ms);
  if (rc != SMG_OK) return fail(h, rc, std::string("file rendezvous in ") + (dir ? dir : "(null)") + " failed");
  return smg_dist_connect(h, all.data());
}

int smg_dist_set_options(smg_handle* h, int exact, int dist_levels, int dist_min_rows) {
  if (!h) return SMG_E_INVALID;
  if (exact < 0 || exact > 2) return fail(h, SMG_E_INVALID, "halo mode must be 0, 1 or 2");
  h->dist.exact = exact;
  h->dist_levels = dist_levels;
  h->dist_min_rows = dist_min_rows;
  if (h->have_plan) drop_graphs_if_device(h);
  return SMG_OK;
}

int smg_dist_info(const smg_handle* h, int64_t* out) {
  if (!h || !out) return SMG_E_INVALID;
  out[0] = h->dist.rank;
  out[1] = h->dist.world;
  out[2] = h->have_plan ? h->plan.dist_levels : 0;
  out[3] = h->dist.exchanges;
  out[4] = h->dist.connected ? 1 : 0;
  out[5] = static_cast<int64_t>(h->dist.slot_doubles);
  out[6] = h->dist.exact;
  out[7] = 0;
  return SMG_OK;
}

int smg_dist_level_info(const smg_handle* h, int lv, int64_t* out) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !out) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  const int W = h->dist.world, me = h->dist.rank;
  for (int i = 0; i < 8; i++) out[i] = 0;
  out[0] = L.layout;
  out[1] = L.nparts;
  out[2] = L.n;
  out[7] = L.n;
  if (L.nparts > 1) {
    int64_t own = 0;
    for (int i = 0; i < L.n; i++) own += L.part[i] == me;
    out[2] = own;
  }
  if (L.layout == smg::LAYOUT_PARTITIONED) {
    const std::vector<int>& pp = L.order.phase_ptr;
    out[6] = pp[static_cast<size_t>(me) * L.n_phases];
    out[7] = pp[static_cast<size_t>(me + 1) * L.n_phases];
    for (int q = 0; q < W; q++) {
      if (q == me) continue;
      if (!L.halo_u.idx.empty()) {
        out[3] += static_cast<int64_t>(L.halo_u.idx[static_cast<size_t>(q) * W + me].size());
        out[4] += static_cast<int64_t>(L.halo_u.idx[static_cast<size_t>(me) * W + q].size());
      }
      if (!L.halo_r.idx.empty()) out[5] += static_cast<int64_t>(L.halo_r.idx[static_cast<size_t>(q) * W + me].size());
    }
  }
  return SMG_OK;
}

int smg_dist_get_part(const smg_handle* h, int lv, int* part_of_row) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !part_of_row) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  for (int i = 0; i < L.n; i++) part_of_row[i] = L.nparts > 1 ? L.part[i] : 0;
  return SMG_OK;
}

int smg_dist_get_exchange(const smg_handle* h, int lv, int which, int src, int dst, int* idx, int* n) {
  SMG_TRY(check_ready(h, false));
  const int W = h->dist.world;
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !n || src < 0 || dst < 0 || src >= W || dst >= W)
    return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  const smg::Exchange* X = nullptr;
  switch (which) {
    case SMG_X_HALO_U: X = &L.halo_u; break;
    case SMG_X_HALO_R: X = &L.halo_r; break;
    case SMG_X_HALO_PU: X = &L.halo_pu; break;
    case SMG_X_GATHER: X = &L.gather_all; break;
    default: return SMG_E_INVALID;
  }
  if (X->idx.size() != static_cast<size_t>(W) * W) {
    *n = 0;
    return SMG_OK;
  }
  const std::vector<int>& v = X->idx[static_cast<size_t>(src) * W + dst];
  *n = static_cast<int>(v.size());
  if (idx)
    for (size_t i = 0; i < v.size(); i++) idx[i] = L.order.perm[v[i]];
  return SMG_OK;
}

// ---- index / topology outputs ------------------------------------------------------
int smg_num_levels(const smg_handle* h) {
  if (!h) return -1;
  if (h->have_plan) return static_cast<int>(h->plan.lv.size());
  return h->have_hierarchy ? static_cast<int>(h->n_rows.size()) : -1;
}

int smg_level_rows(const smg_handle* h, int lv) {
  if (!h || !h->have_plan || lv < 0 || lv >= static_cast<int>(h->plan.lv.size())) return -1;
  return h->plan.lv[lv].n;
}

int smg_num_unknown(const smg_handle* h) {
  if (!h || !h->have_plan) return -1;
  return static_cast<int>(h->plan.unknown.size());
}

int smg_get_unknown(const smg_handle* h, int* unknown) {
  SMG_TRY(check_ready(h, false));
  if (!unknown) return SMG_E_INVALID;
  std::copy(h->plan.unknown.begin(), h->plan.unknown.end(), unknown);
  return SMG_OK;
}

int smg_get_keep(const smg_handle* h, int lv, int* keep, int* n_keep) {
  SMG_TRY(check_ready(h, false));
  if (lv < 1 || lv >= static_cast<int>(h->plan.lv.size()) || !n_keep) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  if (!L.pruned) {
    *n_keep = -1;
    return SMG_OK;
  }
  *n_keep = static_cast<int>(L.keep.size());
  if (keep) std::copy(L.keep.begin(), L.keep.end(), keep);
  return SMG_OK;
}

static const Csc* pick_matrix(const smg_handle* h, int lv, int which) {
  const smg::Plan& pl = h->plan;
  const int nlev = static_cast<int>(pl.lv.size());
  switch (which) {
    case SMG_MAT_A: return (lv >= 0 && lv < nlev) ? &pl.lv[lv].A : nullptr;
    case SMG_MAT_P: return (lv >= 1 && lv < nlev) ? &pl.lv[lv].P : nullptr;
    case SMG_MAT_PT: return (lv >= 1 && lv < nlev) ? &pl.lv[lv].PT : nullptr;
    case SMG_MAT_LHS: return &pl.LHS;
    case SMG_MAT_AUK: return pl.has_fixed ? &pl.Auk : nullptr;
  }
  return nullptr;
}

int smg_matrix_dims(const smg_handle* h, int lv, int which, int* rows, int* cols, int* nnz) {
  SMG_TRY(check_ready(h, false));
  const Csc* m = pick_matrix(h, lv, which);
  if (!m || !rows || !cols || !nnz) return SMG_E_INVALID;
  *rows = m->rows;
  *cols = m->cols;
  *nnz = m->nnz();
  return SMG_OK;
}

int smg_matrix_copy(smg_handle* h, int lv, int which, int* colptr, int* rowidx, double* val) {
  SMG_TRY(check_ready(h, false));
  const Csc* m = pick_matrix(h, lv, which);
  if (!m) return fail(h, SMG_E_INVALID, "no such matrix");
  if (colptr) std::copy(m->colptr.begin(), m->colptr.end(), colptr);
  if (rowidx) std::copy(m->rowidx.begin(), m->rowidx.end(), rowidx);
  if (!val || m->nnz() == 0) return SMG_OK;
  if (which == SMG_MAT_P || which == SMG_MAT_PT) {
    std::copy(m->val.begin(), m->val.end(), val);
    return SMG_OK;
  }
  if (h->plan_only) return fail(h, SMG_E_STATE, "plan-only handle holds no matrix values");
  SMG_TRY(set_device(h));
  const double* src = nullptr;
  if (which == SMG_MAT_A) src = h->lv[lv].a_val.p;
  else if (which == SMG_MAT_LHS) {
    // LHS values = caller's A gathered (mg[0].A additionally carries the 1e-12 shift
    // only when level 0 is the coarsest, which nlev >= 2 excludes)
    src = h->lv[0].a_val.p;
  } else src = h->auk_csc_val.p;
  SMG_CUDA(h, cudaMemcpyAsync(val, src, sizeof(double) * m->nnz(), cudaMemcpyDeviceToHost, h->stream));
  SMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return SMG_OK;
}

int smg_get_diag(smg_handle* h, int lv, double* diag) {
  SMG_TRY(check_ready(h, true));
  SMG_TRY(valid_level(h, lv, false));
  if (!diag) return SMG_E_INVALID;
  SMG_TRY(set_device(h));
  return stage_out(h, lv, h->lv[lv].diag.p, h->st_a, diag, 1);
}

int smg_get_phases(const smg_handle* h, int lv, int* n_phases, int* phase_of_row) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !n_phases) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  *n_phases = L.n_phases;
  if (phase_of_row) std::copy(L.phase.begin(), L.phase.end(), phase_of_row);
  return SMG_OK;
}

int smg_get_row_order(const smg_handle* h, int lv, int* perm, int* n_groups, int* group_ptr) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size())) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  if (perm) std::copy(L.order.perm.begin(), L.order.perm.end(), perm);
  if (n_groups) *n_groups = static_cast<int>(L.order.phase_ptr.size()) - 1;
  if (group_ptr) std::copy(L.order.phase_ptr.begin(), L.order.phase_ptr.end(), group_ptr);
  return SMG_OK;
}

int smg_level_padded_nnz(const smg_handle* h, int lv, int64_t* padded) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !padded) return SMG_E_INVALID;
  *padded = h->plan.lv[lv].sellA.padded();
  return SMG_OK;
}

int smg_level_stats(const smg_handle* h, int lv, int64_t* out) {
  SMG_TRY(check_ready(h, false));
  if (lv < 0 || lv >= static_cast<int>(h->plan.lv.size()) || !out) return SMG_E_INVALID;
  const smg::LevelPlan& L = h->plan.lv[lv];
  int64_t pnz = 0;
  for (double v : L.P.val) pnz += (v != 0.0);
  out[0] = L.n;
  out[1] = L.A.nnz();
  out[2] = L.Alive.nnz();
  out[3] = L.sellA.padded();
  out[4] = pnz;
  out[5] = L.sellP.padded();
  out[6] = L.sellPT.padded();
  out[7] = L.n_phases;
  return SMG_OK;
}

int smg_patch_plan(const smg_handle* h, int lv, int kind, int iters, int target_rows, int smem_limit,
                   int verify, int64_t* out) {
  SMG_TRY(check_ready(h, false));
  if (!out || (kind != smg::PATCH_DOWN && kind != smg::PATCH_UP)) return SMG_E_INVALID;
  smg::PatchSet ps;
  std::string why;
  if (!smg::build_patches(h->plan, lv, kind, iters, target_rows, smem_limit > 0 ? smem_limit : 226 * 1024, 1, &ps, &why))
    return fail(const_cast<smg_handle*>(h), SMG_E_UNSUPPORTED, "patch plan: " + why);
  if (verify) {
    const std::string err = smg::verify_patches(h->plan, ps);
    if (!err.empty()) return fail(const_cast<smg_handle*>(h), SMG_E_INTERNAL, "patch plan: " + err);
  }
  out[0] = ps.n_patches;
  out[1] = ps.sum_own;
  out[2] = ps.sum_loc;
  out[3] = ps.sum_b;
  out[4] = ps.sum_updates;
  out[5] = ps.max_blob_bytes;
  out[6] = ps.max_vec_doubles;
  out[7] = static_cast<int64_t>(ps.blob.size());
  out[8] = ps.max_passes;
  out[9] = ps.sum_entries;
  return SMG_OK;
}

int smg_solve_on_device(const smg_handle* h) { return h ? h->last_solve_on_device : 0; }

int smg_level_patched(const smg_handle* h, int lv) {
  if (!h || h->plan_only || lv < 0 || lv >= static_cast<int>(h->lv.size())) return 0;
  return h->lv[lv].patched ? h->lv[lv].patch_down.view.n_patches : 0;
}

// ---- measurement -----------------------------------------------------------------
int smg_time_kernel(smg_handle* h, int which, int lv, int k, int reps, int flush_l2,
                    float* ms_per_rep, int* launches_per_rep) {
  SMG_TRY(check_ready(h, true));
  if (!ms_per_rep || k < 1 || reps < 1) return fail(h, SMG_E_INVALID, "bad argument");
  const bool need_coarser = which == SMG_K_RESTRICT || which == SMG_K_PROLONG_ADD;
  if (which == SMG_K_COARSE_SOLVE) lv = static_cast<int>(h->lv.size()) - 1;
  if (which == SMG_K_VCYCLE || which == SMG_K_MG_ITERATION) lv = 0;
  SMG_TRY(valid_level(h, lv, need_coarser));
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  const size_t flush_cnt = size_t(1) << 25;  // 256 MB of doubles > 126 MB L2
  if (flush_l2) SMG_CUDA(h, h->flush.reserve(flush_cnt));
  std::vector<cudaEvent_t> ev(static_cast<size_t>(reps) * 2);
  for (auto& e : ev) SMG_CUDA(h, cudaEventCreate(&e));
  LevelDev& L = h->lv[lv];
  int64_t per_rep = 0;
  int rc = SMG_OK;
  for (int rep = -1; rep < reps && rc == SMG_OK; rep++) {  // rep -1 = warm-up
    if (flush_l2) smg::launch_fill(h->flush.p, 1.0, static_cast<int64_t>(flush_cnt), h->stream);
    const int64_t before = h->launches;
    if (rep >= 0) cudaEventRecord(ev[2 * rep], h->stream);
    switch (which) {
      case SMG_K_RESIDUAL: residual_device(h, lv, L.b.p, L.u.p, L.r.p, k); break;
      case SMG_K_RELAX_SWEEP: relax_device(h, lv, 1, L.b.p, L.u.p, k); break;
      case SMG_K_RELAX_PRE: relax_device(h, lv, h->opt.pre_relax, L.b.p, L.u.p, k); break;
      case SMG_K_RESTRICT: restrict_device(h, lv, L.r.p, h->lv[lv + 1].b.p, k); break;
      case SMG_K_PROLONG_ADD: prolong_device(h, lv, h->lv[lv + 1].u.p, L.u.p, k, true); break;
      case SMG_K_RESIDUAL_NORM: {
        // device part only (no host read-back inside the timed region)
        const int nb = smg::residual_norm_blocks(L.n);
        if (h->norm_scratch.reserve(static_cast<size_t>(nb)) != cudaSuccess ||
            h->norm_out.reserve(4) != cudaSuccess) {
          rc = fail(h, SMG_E_CUDA, "alloc");
          break;
        }
        smg::launch_residual_norm2(own_rows(h, L, L.sellA.view()), L.b.p, L.u.p, L.n,
                                   std::min(k, smg::kMaxK), h->norm_scratch.p, h->norm_counter.p, h->norm_out.p, h->stream);
        h->launches += 1;
        break;
      }
      case SMG_K_COARSE_SOLVE: coarse_solve_device(h, L.b.p, L.u.p, k); break;
      case SMG_K_VCYCLE: rc = vcycle_run(h, 0, h->opt.pre_relax, h->opt.post_relax, k); break;
      case SMG_K_MG_ITERATION: {
        double res = 0.0;
        rc = residual_norm_device(h, 0, L.b.p, L.u.p, k, &res);
        if (rc == SMG_OK) rc = vcycle_run(h, 0, h->opt.pre_relax, h->opt.post_relax, k);
        break;
      }
      default: rc = fail(h, SMG_E_INVALID, "unknown kernel id"); break;
    }
    if (rep >= 0) cudaEventRecord(ev[2 * rep + 1], h->stream);
    per_rep = h->launches - before;
  }
  cudaError_t e = cudaStreamSynchronize(h->stream);
  double total = 0.0;
  if (rc == SMG_OK && e == cudaSuccess)
    for (int rep = 0; rep < reps; rep++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[2 * rep], ev[2 * rep + 1]);
      total += ms;
    }
  for (auto& x : ev) cudaEventDestroy(x);
  if (rc != SMG_OK) return rc;
  if (e != cudaSuccess) return fail(h, SMG_E_CUDA, cudaGetErrorString(e));
  SMG_TRY(check_launch(h, "time_kernel"));
  *ms_per_rep = static_cast<float>(total / reps);
  if (launches_per_rep) *launches_per_rep = static_cast<int>(per_rep);
  return SMG_OK;
}

int smg_trace_iteration(smg_handle* h, int k, int max_events, char* names, int names_cap,
                        double* t0_us, double* t1_us, int* n_events) {
  SMG_TRY(check_ready(h, true));
  if (!names || !t0_us || !t1_us || !n_events || k < 1 || max_events < 1)
    return fail(h, SMG_E_INVALID, "bad argument");
  if (h->dist.host_group) return fail(h, SMG_E_UNSUPPORTED, "no CUDA graphs when ranks share a device");
  SMG_TRY(set_device(h));
  SMG_TRY(ensure_k(h, k));
  DevBuf<unsigned long long> buf;
  SMG_CUDA(h, buf.alloc(static_cast<size_t>(max_events) * 2));
  std::vector<unsigned long long> init(static_cast<size_t>(max_events) * 2);
  for (int i = 0; i < max_events; i++) {
    init[2 * i] = ~0ull;
    init[2 * i + 1] = 0ull;
  }
  LevelDev& L0 = h->lv[0];
  const int64_t launches_before = h->launches;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // the traced graph is built with the trace slots baked in and never cached
  SMG_CUDA(h, h->norm_scratch.reserve(static_cast<size_t>(smg::residual_norm_blocks(L0.n))));
  SMG_CUDA(h, h->norm_out.reserve(4));
  if (dist_on(h)) SMG_CUDA(h, h->dist.normv.reserve(static_cast<size_t>(h->dist.world) * 16));
  smg::trace_start(buf.p, max_events);
  smg::trace_label("norm");
  cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
  if (e == cudaSuccess) {
    const bool part = dist_on(h) && L0.layout == smg::LAYOUT_PARTITIONED;
    smg::launch_residual_norm2(own_rows(h, L0, L0.sellA.view()), L0.b.p, L0.u.p, L0.n,
                               std::min(k, smg::kMaxK), h->norm_scratch.p, h->norm_counter.p,
                               part ? h->dist.normv.p + h->dist.rank : h->norm_out.p, h->stream);
    if (part) exchange(h, h->dist.x_norm, h->dist.normv.p, h->dist.world, 1);
    vcycle_device(h, 0, h->opt.pre_relax, h->opt.post_relax, k);
    e = cudaStreamEndCapture(h->stream, &graph);
  }
  smg::trace_stop();
  h->launches = launches_before;
  if (e != cudaSuccess) return fail(h, SMG_E_CUDA, std::string("trace capture: ") + cudaGetErrorString(e));
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(h, SMG_E_CUDA, std::string("trace instantiate: ") + cudaGetErrorString(e));
  for (int rep = 0; rep < 3; rep++) {  // the last replay is the one reported
    cudaMemcpyAsync(buf.p, init.data(), init.size() * sizeof(unsigned long long),
                    cudaMemcpyHostToDevice, h->stream);
    cudaGraphLaunch(exec, h->stream);
  }
  std::vector<unsigned long long> out(init.size());
  drain_if_shared(h);
  cudaMemcpyAsync(out.data(), buf.p, out.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                  h->stream);
  e = cudaStreamSynchronize(h->stream);
  cudaGraphExecDestroy(exec);
  if (e != cudaSuccess) return fail(h, SMG_E_CUDA, cudaGetErrorString(e));
  const int n = smg::trace_count();
  unsigned long long origin = ~0ull;
  for (int i = 0; i < n; i++) origin = std::min(origin, out[2 * i]);
  std::string all;
  for (int i = 0; i < n; i++) {
    t0_us[i] = (out[2 * i] - origin) * 1e-3;
    t1_us[i] = (out[2 * i + 1] - origin) * 1e-3;
    all += smg::trace_name(i);
    all += '\n';
  }
  std::snprintf(names, static_cast<size_t>(names_cap), "%s", all.c_str());
  *n_events = n;
  return SMG_OK;
}

int64_t smg_launch_count(const smg_handle* h) { return h ? h->launches : 0; }

int smg_get_timings(const smg_handle* h, double* ms, int n) {
  if (!h || !ms || n < 0) return SMG_E_INVALID;
  for (int i = 0; i < n && i < 8; i++) ms[i] = h->timings[i];
  return SMG_OK;
}

void* smg_get_stream(const smg_handle* h) { return h ? static_cast<void*>(h->stream) : nullptr; }

}  // extern "C"

// kernels.cu -- hand-written sm_100a kernels of the surface-multigrid V-cycle.
//
// All hot-path matrices are stored SELL-32 (plan.hpp): a slice is 32 consecutive
// rows, stored column-major inside the slice, so the 32 lanes of a warp (one row
// each) read 32 consecutive doubles (256 B) + 32 consecutive int32 (128 B) per
// step: fully coalesced, and the per-row accumulation order is exactly the
// reference's ascending-index order.  The kernels are HBM-bound (about 0.17
// flop/byte): no tensor cores, no shared-memory tiling; matrix streams bypass L1
// (ld.global.nc.L1::no_allocate) so L1 is left to the gathered vector entries.
//
// Arithmetic: __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn everywhere on the parity
// surface, so nvcc never contracts a*b+c into an FMA: the reference is built
// without FMA (x86-64 baseline, 03_mg_solver/CMakeLists.txt:2,21-24) and bit
// parity in wavefront mode depends on it.
#include "kernels.hpp"
#include "mcf_core.hpp"
#include "patch.hpp"

#include <algorithm>
#include <string>
#include <vector>

namespace smg {

namespace {

constexpr int kBlock = 256;  // threads per CTA = 8 slices
constexpr int kPre = 8;      // matrix entries per row prefetched into registers
constexpr int kWide2 = 16;   // widest slice the two-rows-per-thread kernels accept (the tail beyond kPre is serial)

// ---- programmatic dependent launch (PDL) ---------------------------------------
// Every hot-path kernel is launched with programmatic stream serialisation: it may
// start while its predecessor is still running.  Before `pdl_wait()` a kernel only
// touches data that no kernel of a solve ever writes (matrix structure / values,
// diagonal); everything mutable (b, u, r) is read and written after the wait, which
// guarantees the predecessor has completed and its writes are visible.  Mutable
// vectors are read with ld.global.cg (L2 is the coherence point; an SM's L1 may hold
// lines cached by CTAs of the still-running predecessor).
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_vec(const double* p) {  // mutable vector data
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

bool g_use_pdl = true;
bool g_use_tma = true;
int g_gs_rows = 2;  // rows per thread of the Gauss-Seidel phase kernel on large phases (1, 2, 4)
int g_multi_rows = 150000;  // rows of a colour phase from which the Gauss-Seidel kernel takes g_gs_rows rows per thread
int g_apply2_rows = 600000;  // rows from which the residual / norm / restriction kernels take two rows per thread
bool g_gs_attr_set = false;

// ---- optional in-kernel timeline (smg_trace_*): every CTA folds %globaltimer at its
// start / end into [min start, max end] of the launch's slot.  slot < 0: off.
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ __forceinline__ unsigned long long global_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_begin(int slot) {
  if (slot >= 0 && threadIdx.x == 0) atomicMin(g_trace_buf + 2 * slot, global_timer());
}
__device__ __forceinline__ void trace_end(int slot) {
  if (slot >= 0 && threadIdx.x == 0) atomicMax(g_trace_buf + 2 * slot + 1, global_timer());
}
// per host thread: handles driven by different threads (one rank per thread) label and
// trace independently
struct TraceState {
  bool on = false;
  int next = 0, cap = 0;
  std::string label;
  std::vector<std::string> names;
};
thread_local TraceState g_trace;

int g_trace_sub = 0;  // extra trace slots of the next launch (stages inside the kernel)
const char* const* g_trace_sub_names = nullptr;

template <class... KArgs, class... Args>
void launch_kernel(const char* name, void (*kernel)(int, KArgs...), int grid, int block, size_t smem,
                   cudaStream_t st, Args... args) {
  int slot = -1;
  if (g_trace.on && g_trace.next + g_trace_sub < g_trace.cap) {
    slot = g_trace.next++;
    g_trace.names.push_back(g_trace.label + " " + name + " g" + std::to_string(grid));
    for (int i = 0; i < g_trace_sub; i++, g_trace.next++)
      g_trace.names.push_back(g_trace.label + " " + name + g_trace_sub_names[i]);
  }
  g_trace_sub = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, slot, static_cast<KArgs>(args)...);
}

// ---- TMA bulk staging of a CTA's matrix chunk ------------------------------------
// A CTA owns kSlices consecutive SELL-32 slices; their column indices and values are
// two contiguous ranges of the SELL arrays, so one elected thread brings each range
// into shared memory with a single cp.async.bulk (TMA, SASS: UBLKCP) that completes on
// an mbarrier.  The matrix is immutable during a solve, so the copy is issued BEFORE
// the PDL wait and overlaps the tail of the previous kernel; it carries an L2
// evict-first policy so the streamed matrix does not push the vectors out of L2.
constexpr int kSlices = kBlock / 32;
constexpr int kStageCapBytes = 46 * 1024;  // dynamic smem per CTA we are willing to use

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SMG_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SMG_DONE;\n"
      "bra SMG_WAIT;\n"
      "SMG_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Per-thread view of its row: first column / value entry (stride 32), width.
struct RowView {
  const int* cp;
  const double* vp;
  int w;
};

// Common prologue.  STAGED: issue the bulk copies of slices [slice0, slice0+kSlices),
// return pointers into shared memory; otherwise pointers into the global SELL arrays.
// Must be called by every thread of the CTA (contains __syncthreads); the caller waits
// with stage_wait() after its PDL wait.
template <bool STAGED, int NSL = kSlices>
__device__ __forceinline__ RowView stage_rows(int slice0, int nslices, int row, bool active,
                                              const int* __restrict__ slice_ptr,
                                              const int* __restrict__ col,
                                              const double* __restrict__ val, int max_chunk,
                                              unsigned char* dyn, uint64_t* bar) {
  RowView rv;
  rv.w = 0;
  rv.cp = col;
  rv.vp = val;
  int base = 0;
  if (active) {
    const int s = row >> 5;
    base = slice_ptr[s];
    rv.w = (slice_ptr[s + 1] - base) >> 5;
  }
  if (STAGED) {
    double* sval = reinterpret_cast<double*>(dyn);
    int* scol = reinterpret_cast<int*>(dyn + (size_t)max_chunk * sizeof(double));
    const int e0 = slice_ptr[slice0];
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      const int s1 = slice0 + NSL < nslices ? slice0 + NSL : nslices;
      const uint32_t nent = static_cast<uint32_t>(slice_ptr[s1] - e0);
      const uint64_t pol = l2_evict_first_policy();
      mbar_expect_tx(bar, nent * 12u);
      if (nent > 0) {
        bulk_g2s(sval, val + e0, nent * 8u, bar, pol);
        bulk_g2s(scol, col + e0, nent * 4u, bar, pol);
      }
    }
    __syncthreads();  // the initialised barrier is visible to every waiter
    rv.cp = scol + (base - e0) + (row & 31);
    rv.vp = sval + (base - e0) + (row & 31);
  } else {
    rv.cp = col + base + (row & 31);
    rv.vp = val + base + (row & 31);
  }
  return rv;
}

template <bool STAGED>
__device__ __forceinline__ void stage_wait(uint64_t* bar) {
  if (STAGED) mbar_wait(bar, 0);
}

// sum[q] += sum over the row's stored entries of val * x[col + q*ldx], entries in
// storage order, products and sums rounded separately (no FMA: the reference build has
// none).  Gathers of a batch of kPre entries are issued together.
template <int K, bool SKIP_DIAG, bool STAGED, int PRE>
__device__ __forceinline__ void row_accumulate_batched(const RowView& rv, int row, const double* x, int ldx,
                                                       double (&sum)[K]) {
  for (int j0 = 0; j0 < rv.w; j0 += PRE) {
    int c[PRE];
    double xv[PRE][K];
#pragma unroll
    for (int t = 0; t < PRE; t++)
      if (j0 + t < rv.w) {
        c[t] = STAGED ? rv.cp[(j0 + t) * 32] : ld_stream_s32(rv.cp + (j0 + t) * 32);
        if (!(SKIP_DIAG && c[t] == row)) {
#pragma unroll
#ifdef SMG_DEBUG_NOGATHER  // timing experiment only: perfectly coalesced "gathers"
          for (int q = 0; q < K; q++) xv[t][q] = ld_vec(x + (row < ldx - 2 ? row : ldx - 2) + (c[t] & 1) + (size_t)q * ldx);
#else
          for (int q = 0; q < K; q++) xv[t][q] = ld_vec(x + c[t] + (size_t)q * ldx);
#endif
        }
      }
#pragma unroll
    for (int t = 0; t < PRE; t++)
      if (j0 + t < rv.w && !(SKIP_DIAG && c[t] == row)) {
        const double v = STAGED ? rv.vp[(j0 + t) * 32] : ld_stream_f64(rv.vp + (j0 + t) * 32);
#pragma unroll
        for (int q = 0; q < K; q++) sum[q] = __dadd_rn(sum[q], __dmul_rn(v, xv[t][q]));
      }
  }
}

// Rows of the fine-level cotangent matrices hold ~7 entries (one batch of kPre); the Galerkin
// operators of mesh-decimated hierarchies hold 15-40: their gathers go out in batches of kPreWide,
// because every batch is one dependent round trip to L2 / DRAM.
constexpr int kPreWide = 24;
template <int K, bool SKIP_DIAG, bool STAGED>
__device__ __forceinline__ void row_accumulate(const RowView& rv, int row, const double* x, int ldx,
                                               double (&sum)[K]) {
  if (K <= 2 && rv.w > kPre) row_accumulate_batched<K, SKIP_DIAG, STAGED, (K <= 2 ? kPreWide : kPre)>(rv, row, x, ldx, sum);
  else row_accumulate_batched<K, SKIP_DIAG, STAGED, kPre>(rv, row, x, ldx, sum);
}

enum { MODE_SPMV = 0, MODE_RESIDUAL = 1, MODE_ADD = 2, MODE_SPMV_ZERO = 3, MODE_NORM = 4 };

// y = M x | y = b - M x | y += M x | y = M x and z = 0  (z: same shape as y)
// rows [rb, re); `blk` = index of this CTA among the CTAs of the range
template <int K, int MODE, bool STAGED>
__device__ __forceinline__ void sell_apply_body(int trace_slot, int rb, int re, int blk, int nslices,
                                                int max_chunk, const int* __restrict__ slice_ptr,
                                                const int* __restrict__ col, const double* __restrict__ val,
                                                const double* x, int ldx, const double* b, double* y,
                                                int ldy, double* z) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  pdl_launch_dependents();
  const int row = (rb & ~31) + blk * kBlock + threadIdx.x;
  const bool active = row >= rb && row < re;
  const RowView rv = stage_rows<STAGED>((rb >> 5) + blk * kSlices, nslices, row, active, slice_ptr, col,
                                        val, max_chunk, dyn, &bar);
  pdl_wait();
  stage_wait<STAGED>(&bar);
  if (active) {
    double sum[K];
#pragma unroll
    for (int q = 0; q < K; q++) sum[q] = 0.0;
    row_accumulate<K, false, STAGED>(rv, row, x, ldx, sum);
#pragma unroll
    for (int q = 0; q < K; q++) {
      const size_t o = row + (size_t)q * ldy;
      if (MODE == MODE_SPMV) y[o] = sum[q];
      if (MODE == MODE_RESIDUAL) y[o] = __dsub_rn(ld_vec(b + o), sum[q]);
      if (MODE == MODE_ADD) y[o] = __dadd_rn(ld_vec(y + o), sum[q]);
      if (MODE == MODE_SPMV_ZERO) {
        y[o] = sum[q];
        z[o] = 0.0;
      }
    }
  }
  trace_end(trace_slot);
}

template <int K, int MODE, bool STAGED>
__global__ void __launch_bounds__(kBlock)
sell_apply_kernel(int trace_slot, int rb, int re, int nslices, int max_chunk, const int* __restrict__ slice_ptr,
                  const int* __restrict__ col, const double* __restrict__ val, const double* x,
                  int ldx, const double* b, double* y, int ldy, double* z) {
  sell_apply_body<K, MODE, STAGED>(trace_slot, rb, re, blockIdx.x, nslices, max_chunk, slice_ptr, col, val, x,
                                   ldx, b, y, ldy, z);
}

// y = M x on several row ranges in ONE launch (the rows a rank restricts into on the
// replicated level below a partitioned one are one range per colour)
template <int K, bool STAGED>
__global__ void __launch_bounds__(kBlock)
sell_spmv_ranges_kernel(int trace_slot, RowRanges rr, int nslices, int max_chunk,
                        const int* __restrict__ slice_ptr, const int* __restrict__ col,
                        const double* __restrict__ val, const double* x, int ldx, double* y, int ldy) {
  int j = 0;
  while (j + 1 < rr.n && (int)blockIdx.x >= rr.blk0[j + 1]) j++;
  sell_apply_body<K, MODE_SPMV, STAGED>(trace_slot, rr.rb[j], rr.re[j], blockIdx.x - rr.blk0[j], nslices,
                                        max_chunk, slice_ptr, col, val, x, ldx, nullptr, y, ldy, nullptr);
}

// Tail of the residual-norm kernels: per-CTA partial sum (fixed-shape warp / CTA reduction),
// then the LAST CTA to finish adds all partial sums in a fixed order, so the result does not
// depend on which CTA that is: deterministic without a second launch.  counter self-resets.
__device__ __forceinline__ void norm_finish(double d2, double* partial, unsigned int* counter, double* out) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_down_sync(0xffffffffu, d2, o);
  __shared__ double wsum[kBlock / 32];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = d2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kBlock / 32; i++) t += wsum[i];
    partial[blockIdx.x] = t;
    __threadfence();
    is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t = 0.0;
  for (int i0 = threadIdx.x; i0 < (int)gridDim.x; i0 += kBlock * 8) {  // eight loads in flight per thread
    double v[8];
#pragma unroll
    for (int a = 0; a < 8; a++) v[a] = i0 + a * kBlock < (int)gridDim.x ? ld_vec(partial + i0 + a * kBlock) : 0.0;
#pragma unroll
    for (int a = 0; a < 8; a++) t += v[a];
  }
  __shared__ double sm[kBlock];
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = kBlock / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out = sm[0];
    *counter = 0u;
  }
}

// Transfer operators have very short rows (a prolongation row holds at most three
// entries): one row per thread makes CTAs that move only a few KB each and the kernel
// becomes bound by CTA turnover.  Here a thread owns R rows (256 apart inside the CTA's
// R*8 slices, staged by ONE bulk copy); all gathers of all its rows are issued before
// the first sum.  W entries per row are held in registers; wider rows finish entry by entry.
template <int K, int MODE, int R, int W>
__global__ void __launch_bounds__(kBlock)
sell_apply_short_kernel(int trace_slot, int rb, int re, int nslices, int max_chunk,
                        const int* __restrict__ slice_ptr, const int* __restrict__ col,
                        const double* __restrict__ val, const double* x, int ldx, const double* b,
                        double* y, int ldy, double* z, unsigned int* counter) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  pdl_launch_dependents();
  const int slice0 = (rb >> 5) + blockIdx.x * (kSlices * R);
  const int row0 = (rb & ~31) + blockIdx.x * (kBlock * R) + threadIdx.x;
  double* sval = reinterpret_cast<double*>(dyn);
  int* scol = reinterpret_cast<int*>(dyn + (size_t)max_chunk * sizeof(double));
  const int e0 = slice_ptr[slice0];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    const int s1 = slice0 + kSlices * R < nslices ? slice0 + kSlices * R : nslices;
    const uint32_t nent = static_cast<uint32_t>(slice_ptr[s1] - e0);
    const uint64_t pol = l2_evict_first_policy();
    mbar_expect_tx(&bar, nent * 12u);
    if (nent > 0) {
      bulk_g2s(sval, val + e0, nent * 8u, &bar, pol);
      bulk_g2s(scol, col + e0, nent * 4u, &bar, pol);
    }
  }
  int off[R], w[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = row0 + r * kBlock;
    w[r] = 0;
    off[r] = 0;
    if (row >= rb && row < re) {
      const int s = row >> 5;
      const int base = slice_ptr[s];
      w[r] = (slice_ptr[s + 1] - base) >> 5;
      off[r] = base - e0 + (row & 31);
    }
  }
  __syncthreads();
  pdl_wait();
  mbar_wait(&bar, 0);
  double xv[R][W][K], yv[R][K];
#pragma unroll
  for (int r = 0; r < R; r++) {
#pragma unroll
    for (int j = 0; j < W; j++)
      if (j < w[r]) {
        const int c = scol[off[r] + j * 32];
#pragma unroll
        for (int q = 0; q < K; q++) xv[r][j][q] = ld_vec(x + c + (size_t)q * ldx);
      }
    if ((MODE == MODE_ADD || MODE == MODE_RESIDUAL || MODE == MODE_NORM) && row0 + r * kBlock >= rb &&
        row0 + r * kBlock < re) {
      const double* src = MODE == MODE_ADD ? y : b;
#pragma unroll
      for (int q = 0; q < K; q++) yv[r][q] = ld_vec(src + row0 + r * kBlock + (size_t)q * ldy);
    }
  }
  double d2 = 0.0;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = row0 + r * kBlock;
    if (row < rb || row >= re) continue;
#pragma unroll
    for (int q = 0; q < K; q++) {
      double sum = 0.0;
#pragma unroll
      for (int j = 0; j < W; j++)
        if (j < w[r]) sum = __dadd_rn(sum, __dmul_rn(sval[off[r] + j * 32], xv[r][j][q]));
      // rows wider than the register window (irregular meshes: a few slices per level): the rest
      // entry by entry, in storage order
      for (int j = W; j < w[r]; j++)
        sum = __dadd_rn(sum, __dmul_rn(sval[off[r] + j * 32], ld_vec(x + scol[off[r] + j * 32] + (size_t)q * ldx)));
      const size_t o = row + (size_t)q * ldy;
      if (MODE == MODE_SPMV) y[o] = sum;
      if (MODE == MODE_ADD) y[o] = __dadd_rn(yv[r][q], sum);
      if (MODE == MODE_RESIDUAL) y[o] = __dsub_rn(yv[r][q], sum);
      if (MODE == MODE_NORM) {
        const double d = __dsub_rn(yv[r][q], sum);
        d2 += d * d;
      }
      if (MODE == MODE_SPMV_ZERO) {
        y[o] = sum;
        z[o] = 0.0;
      }
    }
  }
  if (MODE == MODE_NORM) norm_finish(d2, y, counter, z);  // y = per-CTA partial sums, z = the sum
  trace_end(trace_slot);
}

template <int K, bool STAGED>
__global__ void __launch_bounds__(kBlock)
sell_residual_norm_kernel(int trace_slot, int rb, int re, int nslices, int max_chunk, const int* __restrict__ slice_ptr,
                          const int* __restrict__ col, const double* __restrict__ val,
                          const double* x, const double* b, int ld, double* partial,
                          unsigned int* counter, double* out) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  pdl_launch_dependents();
  const int row = (rb & ~31) + blockIdx.x * kBlock + threadIdx.x;
  const bool active = row >= rb && row < re;
  const RowView rv = stage_rows<STAGED>((rb >> 5) + blockIdx.x * kSlices, nslices, row, active, slice_ptr, col,
                                        val, max_chunk, dyn, &bar);
  pdl_wait();
  stage_wait<STAGED>(&bar);
  double d2 = 0.0;
  if (active) {
    double sum[K];
#pragma unroll
    for (int q = 0; q < K; q++) sum[q] = 0.0;
    row_accumulate<K, false, STAGED>(rv, row, x, ld, sum);
#pragma unroll
    for (int q = 0; q < K; q++) {
      const double d = __dsub_rn(ld_vec(b + row + (size_t)q * ld), sum[q]);
      d2 += d * d;
    }
  }
  norm_finish(d2, partial, counter, out);
  trace_end(trace_slot);
}

// Gauss-Seidel phase with R rows per thread (rows 256 apart inside the CTA's R*8 slices,
// staged by one TMA bulk copy pair): R times fewer, fatter CTAs.  A phase of a large
// level is bound by CTA turnover (launch + set-up + one DRAM round trip per CTA), not by
// bandwidth; this amortises it.  Phase-barrier synchronisation only; kPre entries per row are
// gathered at once, wider rows (a few slices on irregular meshes) finish entry by entry.
template <int K, int R>
__global__ void __launch_bounds__(kBlock)
sell_gs_phase_multi_kernel(int trace_slot, int row0, int ps, int pe, int nslices, int max_chunk,
                           const int* __restrict__ slice_ptr, const int* __restrict__ col,
                           const double* __restrict__ val, const double* __restrict__ diag,
                           const double* b, double* u, int ld, int pf_slice0, int pf_slice_end) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  pdl_launch_dependents();
  const int slice0 = (row0 >> 5) + blockIdx.x * (kSlices * R);
  const int rbase = row0 + blockIdx.x * (kBlock * R) + threadIdx.x;
  double* sval = reinterpret_cast<double*>(dyn);
  int* scol = reinterpret_cast<int*>(dyn + (size_t)max_chunk * sizeof(double));
  const int e0 = slice_ptr[slice0 < nslices ? slice0 : nslices];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    const int s1 = slice0 + kSlices * R < nslices ? slice0 + kSlices * R : nslices;
    const uint32_t nent = static_cast<uint32_t>(slice_ptr[s1] - e0);
    const uint64_t pol = l2_evict_first_policy();
    mbar_expect_tx(&bar, nent * 12u);
    if (nent > 0) {
      bulk_g2s(sval, val + e0, nent * 8u, &bar, pol);
      bulk_g2s(scol, col + e0, nent * 4u, &bar, pol);
    }
  }
  if (pf_slice0 >= 0 && threadIdx.x == 32) {  // next phase's chunk -> L2 (see the R = 1 kernel)
    const int s0 = pf_slice0 + blockIdx.x * (kSlices * R);
    if (s0 < pf_slice_end) {
      const int s1 = s0 + kSlices * R < pf_slice_end ? s0 + kSlices * R : pf_slice_end;
      const int p0 = slice_ptr[s0];
      const uint32_t nent = static_cast<uint32_t>(slice_ptr[s1] - p0);
      if (nent > 0) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(val + p0), "r"(nent * 8u) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(col + p0), "r"(nent * 4u) : "memory");
      }
    }
  }
  int off[R], w[R];
  double d[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = rbase + r * kBlock;
    w[r] = 0;
    off[r] = 0;
    d[r] = 1.0;
    if (row >= ps && row < pe) {
      const int s = row >> 5;
      const int base = slice_ptr[s];
      w[r] = (slice_ptr[s + 1] - base) >> 5;
      off[r] = base - e0 + (row & 31);
      d[r] = ld_stream_f64(diag + row);
    }
  }
  __syncthreads();
  pdl_wait();
  mbar_wait(&bar, 0);
  double xv[R][kPre][K], bv[R][K];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = rbase + r * kBlock;
#pragma unroll
    for (int j = 0; j < kPre; j++)
      if (j < w[r]) {
        const int c = scol[off[r] + j * 32];
        if (c != row) {
#pragma unroll
          for (int q = 0; q < K; q++) xv[r][j][q] = ld_vec(u + c + (size_t)q * ld);
        }
      }
    if (w[r] > 0) {
#pragma unroll
      for (int q = 0; q < K; q++) bv[r][q] = ld_vec(b + row + (size_t)q * ld);
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int row = rbase + r * kBlock;
    if (w[r] == 0) continue;
#pragma unroll
    for (int q = 0; q < K; q++) {
      double sum = 0.0;
#pragma unroll
      for (int j = 0; j < kPre; j++)
        if (j < w[r] && scol[off[r] + j * 32] != row)
          sum = __dadd_rn(sum, __dmul_rn(sval[off[r] + j * 32], xv[r][j][q]));
      for (int j = kPre; j < w[r]; j++) {  // rows wider than the register window, in storage order
        const int c = scol[off[r] + j * 32];
        if (c != row) sum = __dadd_rn(sum, __dmul_rn(sval[off[r] + j * 32], ld_vec(u + c + (size_t)q * ld)));
      }
      u[row + (size_t)q * ld] = __ddiv_rn(__dsub_rn(bv[r][q], sum), d[r]);
    }
  }
  trace_end(trace_slot);
}


// One phase (colour / wavefront level) of Gauss-Seidel: rows [ps,pe) are mutually
// independent, so updating them in place and in parallel is exactly the sequential
// sweep of mg_VCycle.cpp:147-158 restricted to those rows.  Phases are separated by
// kernel boundaries (PDL: the matrix chunk is fetched before the wait).
template <int K, bool STAGED>
__global__ void __launch_bounds__(kBlock)
sell_gs_phase_kernel(int trace_slot, int row0, int ps, int pe, int nslices, int max_chunk,
                     const int* __restrict__ slice_ptr, const int* __restrict__ col,
                     const double* __restrict__ val, const double* __restrict__ diag,
                     const double* b, double* u, int ld, const GsFlow flow) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  pdl_launch_dependents();
  const int row = row0 + blockIdx.x * kBlock + threadIdx.x;
  const bool active = row >= ps && row < pe;
  const RowView rv = stage_rows<STAGED>((row0 >> 5) + blockIdx.x * kSlices, nslices, row, active,
                                        slice_ptr, col, val, max_chunk, dyn, &bar);
  const double d = active ? ld_stream_f64(diag + row) : 1.0;
  // While this phase gathers, stores and drains, DRAM would idle: ask for the matrix chunk
  // that the same CTA index of the NEXT phase will stream, so that its TMA hits L2.
  if (flow.pf_slice0 >= 0 && threadIdx.x == 32) {
    const int s0 = flow.pf_slice0 + blockIdx.x * kSlices;
    if (s0 < flow.pf_slice_end) {
      const int s1 = s0 + kSlices < flow.pf_slice_end ? s0 + kSlices : flow.pf_slice_end;
      const int e0 = slice_ptr[s0];
      const uint32_t nent = static_cast<uint32_t>(slice_ptr[s1] - e0);
      if (nent > 0) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(val + e0), "r"(nent * 8u) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(col + e0), "r"(nent * 4u) : "memory");
      }
    }
  }
  pdl_wait();
  stage_wait<STAGED>(&bar);
  if (active) {
    double sum[K];
#pragma unroll
    for (int q = 0; q < K; q++) sum[q] = 0.0;
    row_accumulate<K, true, STAGED>(rv, row, u, ld, sum);
#pragma unroll
    for (int q = 0; q < K; q++) {
      const size_t o = row + (size_t)q * ld;
      u[o] = __ddiv_rn(__dsub_rn(ld_vec(b + o), sum[q]), d);
    }
  }
  trace_end(trace_slot);
}

// ---- multi-GPU halo exchange (see kernels.hpp::XchgPeer) --------------------------------
// Low-latency protocol without fences (the scheme of NCCL's LL protocol): a value travels as
// two 8-byte words {low half | epoch << 32} and {high half | epoch << 32} into the peer's
// staging slot, and the receiver polls the two words of its own slot until BOTH carry the
// expected epoch.  Correctness needs only what PTX guarantees on every interconnect (NVLink,
// NVSwitch, PCIe peer-to-peer, one GPU's L2): an aligned 8-byte access is single-copy atomic.
// The two words are issued as one 16-byte vector store / load for throughput, but nothing
// depends on the 16 bytes arriving together.  A fence + flag protocol costs ~10 us per
// exchange on B200 (system-scope fences wait for the NVLink round trip); this one costs one
// store latency.  After the payload every pair of ranks exchanges one extra sync word, so a
// rank can never run two exchanges ahead of a peer even when a pair has no payload, which
// is what makes the parity double-buffering of the slots safe.
// Grid (ctas per peer, npeers); ctrl layout: [2] error flag, [8 + 64 + peer] exit counter,
// [8 + 128 + peer] epoch of the peer group (advanced by the last CTA of the group to exit,
// i.e. after every CTA of the group has read it).  A wait that exceeds the timeout (a peer
// died) sets ctrl[2] and falls through instead of hanging the GPU; once set, later
// exchanges do not wait again.
// late_trigger: release the dependent kernel only after the wait.  Ranks that share one
// device need it: with the early trigger the whole chain of later kernels of a CUDA graph
// becomes resident (blocked in griddepcontrol.wait) and can fill the device while the peer
// rank, whose push would unblock it, cannot get a CTA scheduled.
unsigned long long g_xchg_timeout_ns = 20ull * 1000 * 1000 * 1000;
constexpr int kXchgMaxPeers = 64;

__device__ __forceinline__ void st_ll(double* slot, size_t i, double v, unsigned long long epoch) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long w0 = (bits & 0xffffffffull) | (epoch << 32);
  const unsigned long long w1 = (bits >> 32) | (epoch << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot + 2 * i), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ld_ll(const double* slot, size_t i, unsigned long long epoch, double* v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot + 2 * i) : "memory");
  *v = __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
  return (w0 >> 32) == epoch && (w1 >> 32) == epoch;
}

__global__ void __launch_bounds__(kXchgThreads)
halo_exchange_kernel(int trace_slot, const XchgPeer* __restrict__ peers, double* vec, int ld, int k,
                     size_t parity_stride, int* ctrl, unsigned long long timeout_ns, int late_trigger,
                     int phases) {
  trace_begin(trace_slot);
  if (!late_trigger) pdl_launch_dependents();
  const XchgPeer pr = peers[blockIdx.y];
  int* leave = ctrl + 8 + kXchgMaxPeers + blockIdx.y;
  int* group_epoch = ctrl + 8 + 2 * kXchgMaxPeers + blockIdx.y;
  const int nthreads = gridDim.x * kXchgThreads;
  const int tid = blockIdx.x * kXchgThreads + threadIdx.x;
  // the send list is immutable: fetch this thread's first index before the dependency wait
  const int first_idx = tid < pr.n_send ? pr.send_idx[tid] : 0;
  pdl_wait();  // vec is final; the previous exchange kernel has completed
  const int epoch32 = ld_acquire_gpu(group_epoch) + 1;
  const unsigned long long epoch = static_cast<unsigned int>(epoch32);
  const size_t par = (epoch32 & 1) ? parity_stride : 0;
  // push: payload, then the sync word
  if (phases & 1) {
    double* dst = pr.remote_slot + par;
    for (int q = 0; q < k; q++)
      for (int i = tid; i < pr.n_send; i += nthreads) {
        const int idx = i == tid ? first_idx : pr.send_idx[i];
        st_ll(dst, (size_t)q * pr.n_send + i, ld_vec(vec + idx + (size_t)q * ld), epoch);
      }
    if (tid == 0) st_ll(dst, (size_t)k * pr.n_send, 0.0, epoch);
  }
  if (!(phases & 2)) {  // push-only launch (host-synchronised mode): the epoch stays
    trace_end(trace_slot);
    return;
  }
  // receive: poll every word of the own slot, scatter
  const double* src = pr.local_slot + par;
  const unsigned long long t0 = global_timer();
  bool dead = false;
  const int total = k * pr.n_recv + 1;  // + sync word
  for (int j = tid; j < total; j += nthreads) {
    double v;
    unsigned spins = 0;
    while (!ld_ll(src, (size_t)j, epoch, &v)) {
      if (dead) break;
      if ((++spins & 1023u) == 0) {
        if (ld_acquire_gpu(ctrl + 2) != 0 || global_timer() - t0 > timeout_ns) {
          atomicExch(ctrl + 2, 1);
          dead = true;
        }
      }
    }
    if (j < total - 1) {
      const int q = j / pr.n_recv, i = j - q * pr.n_recv;
      vec[pr.recv_idx[i] + (size_t)q * ld] = v;
    }
  }
  __syncthreads();
  if (late_trigger) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    if (gridDim.x == 1) {
      st_release_gpu(group_epoch, epoch32);
    } else if (atomicAdd(leave, 1) == (int)gridDim.x - 1) {
      atomicExch(leave, 0);
      st_release_gpu(group_epoch, epoch32);
    }
  }
  trace_end(trace_slot);
}

inline int blocks_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

}  // namespace

void set_pdl_enabled(bool on) { g_use_pdl = on; }
void trace_start(unsigned long long* dev_buf, int cap) {
  cudaMemcpyToSymbol(g_trace_buf, &dev_buf, sizeof(dev_buf));
  g_trace.on = true;
  g_trace.next = 0;
  g_trace.cap = cap;
  g_trace.names.clear();
}
void trace_stop() { g_trace.on = false; }
void trace_label(const char* label) { g_trace.label = label; }
int trace_count() { return g_trace.next; }
const char* trace_name(int i) { return g_trace.names[i].c_str(); }
void set_tma_enabled(bool on) { g_use_tma = on; }
void set_apply2_rows(int rows) { g_apply2_rows = rows > 0 ? rows : 600000; }
void set_multi_rows(int rows) { g_multi_rows = rows > 0 ? rows : 150000; }
void set_gs_rows(int r) {
  g_gs_rows = r >= 4 ? 4 : (r >= 2 ? 2 : 1);
  g_gs_attr_set = true;
  // more than 48 KB of dynamic shared memory needs an opt-in
  cudaFuncSetAttribute(sell_gs_phase_multi_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(sell_gs_phase_multi_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int big = 100 * 1024;
  cudaFuncSetAttribute(sell_apply_short_kernel<1, MODE_SPMV, 2, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_short_kernel<1, MODE_RESIDUAL, 2, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_short_kernel<1, MODE_ADD, 2, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_short_kernel<1, MODE_SPMV_ZERO, 2, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_short_kernel<1, MODE_NORM, 2, kPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
}

#define SMG_DISPATCH_K(k, ...)                 \
  switch (k) {                                 \
    case 1: { constexpr int K = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int K = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int K = 3; __VA_ARGS__; } break; \
    default: { constexpr int K = 4; __VA_ARGS__; } break; \
  }

namespace {
const char* const kApplyNames[5] = {"spmv", "residual", "prolong_add", "restrict_zero", "residual_norm"};
inline size_t stage_bytes(const SellDev& M) { return static_cast<size_t>(M.max_chunk) * 12; }
// Large levels keep the stage small (several resident CTAs per SM).  A level with at most one
// CTA per SM may use most of the SM's shared memory: its rows can be wide (Galerkin operators of
// decimated meshes), and without staging every batch of entries costs two dependent global
// round trips (indices, then gathers) instead of one.
constexpr int kStageCapSmallLevel = 200 * 1024;
constexpr int kSmallLevelRows = 148 * kBlock;
template <int K>
void allow_big_stage_k() {
  const int big = kStageCapSmallLevel;
  cudaFuncSetAttribute(sell_apply_kernel<K, MODE_SPMV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_kernel<K, MODE_RESIDUAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_kernel<K, MODE_ADD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_apply_kernel<K, MODE_SPMV_ZERO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_spmv_ranges_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_residual_norm_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(sell_gs_phase_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
}
void allow_big_stage() {
  static bool done[64] = {false};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done[dev]) return;
  allow_big_stage_k<1>();
  allow_big_stage_k<2>();
  allow_big_stage_k<3>();
  allow_big_stage_k<4>();
  done[dev] = true;
}
inline bool use_staged(const SellDev& M) {
  if (!g_use_tma || M.max_chunk <= 0) return false;
  if (stage_bytes(M) <= static_cast<size_t>(kStageCapBytes)) return true;
  if (M.nrows <= kSmallLevelRows && stage_bytes(M) <= static_cast<size_t>(kStageCapSmallLevel)) {
    allow_big_stage();
    return true;
  }
  // large level with wide rows (block 3n x 3n systems: ~21 entries per row): two resident CTAs
  // per SM with their chunks staged beat more CTAs that pay two round trips per batch
  if (M.max_width > kPre && stage_bytes(M) <= static_cast<size_t>(100 * 1024)) {
    allow_big_stage();
    return true;
  }
  return false;
}

template <int MODE>
void launch_apply(const SellDev& M, const double* v, const double* x, int ldx, const double* b,
                  double* y, int ldy, double* z, int k, cudaStream_t st) {
  const int rb = M.row_begin(), re = M.row_end();
  if (re <= rb) return;
  const int span = re - (rb & ~31);  // rows covered by the grid (it starts at a slice boundary)
  constexpr int R = 4, W = 3;
  if (g_use_tma && M.max_width <= W && M.max_chunk32 > 0 &&
      static_cast<size_t>(M.max_chunk32) * 12 <= static_cast<size_t>(kStageCapBytes)) {
    const int gs = blocks_for(span, kBlock * R);
    SMG_DISPATCH_K(k, launch_kernel(kApplyNames[MODE], sell_apply_short_kernel<K, MODE, R, W>, gs, kBlock,
                                    static_cast<size_t>(M.max_chunk32) * 12, st, rb, re, M.nslices,
                                    M.max_chunk32, M.slice_ptr, M.col, v, x, ldx, b, y, ldy, z, nullptr));
    return;
  }
  // large levels, k = 1: two rows per thread (fewer, fatter CTAs; see the Gauss-Seidel kernel)
  if (g_gs_rows > 1 && g_use_tma && k == 1 && M.max_width <= kWide2 && span >= g_apply2_rows &&
      M.max_chunk16 > 0 && static_cast<size_t>(M.max_chunk16) * 12 <= 100 * 1024) {
    if (!g_gs_attr_set) set_gs_rows(g_gs_rows);
    launch_kernel(kApplyNames[MODE], sell_apply_short_kernel<1, MODE, 2, kPre>, blocks_for(span, kBlock * 2),
                  kBlock, static_cast<size_t>(M.max_chunk16) * 12, st, rb, re, M.nslices, M.max_chunk16,
                  M.slice_ptr, M.col, v, x, ldx, b, y, ldy, z, nullptr);
    return;
  }
  const int g = blocks_for(span, kBlock);
  if (use_staged(M)) {
    SMG_DISPATCH_K(k, launch_kernel(kApplyNames[MODE], sell_apply_kernel<K, MODE, true>, g, kBlock, stage_bytes(M), st,
                                    rb, re, M.nslices, M.max_chunk, M.slice_ptr, M.col, v, x, ldx,
                                    b, y, ldy, z));
  } else {
    SMG_DISPATCH_K(k, launch_kernel(kApplyNames[MODE], sell_apply_kernel<K, MODE, false>, g, kBlock, 0, st, rb,
                                    re, M.nslices, M.max_chunk, M.slice_ptr, M.col, v, x, ldx, b, y,
                                    ldy, z));
  }
}
}  // namespace

void launch_spmv(const SellDev& M, bool use_valT, const double* x, int ldx, double* y, int ldy,
                 int k, cudaStream_t st) {
  launch_apply<MODE_SPMV>(M, use_valT ? M.valT : M.val, x, ldx, nullptr, y, ldy, nullptr, k, st);
}

void launch_spmv_ranges(const SellDev& M, const RowRanges& ranges, const double* x, int ldx, double* y,
                        int ldy, int k, cudaStream_t st) {
  RowRanges rr = ranges;
  rr.blk0[0] = 0;
  for (int j = 0; j < rr.n; j++) {
    const int span = rr.re[j] > rr.rb[j] ? rr.re[j] - (rr.rb[j] & ~31) : 0;
    rr.blk0[j + 1] = rr.blk0[j] + blocks_for(span, kBlock);
  }
  const int g = rr.blk0[rr.n];
  if (g <= 0) return;
  if (use_staged(M)) {
    SMG_DISPATCH_K(k, launch_kernel("spmv_ranges", sell_spmv_ranges_kernel<K, true>, g, kBlock, stage_bytes(M), st,
                                    rr, M.nslices, M.max_chunk, M.slice_ptr, M.col, M.val, x, ldx, y, ldy));
  } else {
    SMG_DISPATCH_K(k, launch_kernel("spmv_ranges", sell_spmv_ranges_kernel<K, false>, g, kBlock, 0, st, rr,
                                    M.nslices, M.max_chunk, M.slice_ptr, M.col, M.val, x, ldx, y, ldy));
  }
}

void launch_spmv_zero(const SellDev& M, const double* x, int ldx, double* y, double* z, int ldy,
                      int k, cudaStream_t st) {
  launch_apply<MODE_SPMV_ZERO>(M, M.val, x, ldx, nullptr, y, ldy, z, k, st);
}

void launch_residual(const SellDev& M, const double* b, const double* x, double* r, int ld, int k,
                     cudaStream_t st) {
  launch_apply<MODE_RESIDUAL>(M, M.valT, x, ld, b, r, ld, nullptr, k, st);
}

void launch_prolong_add(const SellDev& M, const double* x, int ldx, double* u, int ldu, int k,
                        cudaStream_t st) {
  launch_apply<MODE_ADD>(M, M.val, x, ldx, nullptr, u, ldu, nullptr, k, st);
}

int residual_norm_blocks(int nrows) { return blocks_for(nrows > 0 ? nrows + 32 : 1, kBlock); }

void launch_residual_norm2(const SellDev& M, const double* b, const double* x, int ld, int k,
                           double* scratch, unsigned int* counter, double* out, cudaStream_t st) {
  const int rb = M.row_begin(), re = M.row_end();
  const int span = re > rb ? re - (rb & ~31) : 0;
  if (span <= 0) {  // a rank without rows on this level contributes 0
    launch_fill(out, 0.0, 1, st);
    return;
  }
  int g = blocks_for(span, kBlock);
  if (g_gs_rows > 1 && g_use_tma && k == 1 && M.max_width <= kWide2 && span >= g_apply2_rows &&
      M.max_chunk16 > 0 && static_cast<size_t>(M.max_chunk16) * 12 <= 100 * 1024) {
    if (!g_gs_attr_set) set_gs_rows(g_gs_rows);
    g = blocks_for(span, kBlock * 2);
    launch_kernel("residual_norm", sell_apply_short_kernel<1, MODE_NORM, 2, kPre>, g, kBlock,
                  static_cast<size_t>(M.max_chunk16) * 12, st, rb, re, M.nslices, M.max_chunk16,
                  M.slice_ptr, M.col, M.valT, x, ld, b, scratch, ld, out, counter);
    return;
  }
  if (use_staged(M)) {
    SMG_DISPATCH_K(k, launch_kernel("residual_norm", sell_residual_norm_kernel<K, true>, g, kBlock, stage_bytes(M),
                                    st, rb, re, M.nslices, M.max_chunk, M.slice_ptr, M.col, M.valT,
                                    x, b, ld, scratch, counter, out));
  } else {
    SMG_DISPATCH_K(k, launch_kernel("residual_norm", sell_residual_norm_kernel<K, false>, g, kBlock, 0, st, rb,
                                    re, M.nslices, M.max_chunk, M.slice_ptr, M.col, M.valT, x, b, ld,
                                    scratch, counter, out));
  }
}

void set_xchg_timeout_ms(long long ms) {
  if (ms > 0) g_xchg_timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
}

int xchg_ctrl_ints() { return 8 + 3 * kXchgMaxPeers; }
int xchg_max_peers() { return kXchgMaxPeers; }

void launch_halo_exchange(const XchgPeer* d_peers, int npeers, int ctas_per_peer, double* vec, int ld,
                          int k, size_t parity_stride, int* ctrl, bool late_trigger, int phases,
                          cudaStream_t st) {
  if (npeers <= 0 || npeers > kXchgMaxPeers) return;
  ctas_per_peer = std::max(1, std::min(ctas_per_peer, 32));
  int slot = -1;
  if (g_trace.on && g_trace.next < g_trace.cap) {
    slot = g_trace.next++;
    g_trace.names.push_back(g_trace.label + " halo_exchange g" + std::to_string(npeers * ctas_per_peer));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas_per_peer, npeers);
  cfg.blockDim = dim3(kXchgThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, halo_exchange_kernel, slot, d_peers, vec, ld, k, parity_stride, ctrl,
                     g_xchg_timeout_ns, late_trigger ? 1 : 0, phases);
}

void launch_gs_phase(const SellDev& M, const double* diag, const double* b, double* u, int ld,
                     int k, int ps, int pe, const GsFlow& flow, cudaStream_t st) {
  if (pe <= ps) return;
  const int row0 = ps & ~31;
  // large phases: R rows per thread (fewer, fatter CTAs)
  if (g_gs_rows > 1 && g_use_tma && k == 1 && M.max_width <= kWide2 &&
      pe - row0 >= g_multi_rows) {  // at least one full wave of fat CTAs
    if (!g_gs_attr_set) set_gs_rows(g_gs_rows);
    const int R = g_gs_rows;
    const int mc = R == 2 ? M.max_chunk16 : M.max_chunk32s;
    if (mc > 0 && static_cast<size_t>(mc) * 12 <= 100 * 1024) {
      const int gr = blocks_for(pe - row0, kBlock * R);
      if (R == 2)
        launch_kernel("gs_phase", sell_gs_phase_multi_kernel<1, 2>, gr, kBlock, static_cast<size_t>(mc) * 12, st,
                      row0, ps, pe, M.nslices, mc, M.slice_ptr, M.col, M.val, diag, b, u, ld,
                      flow.pf_slice0, flow.pf_slice_end);
      else
        launch_kernel("gs_phase", sell_gs_phase_multi_kernel<1, 4>, gr, kBlock, static_cast<size_t>(mc) * 12, st,
                      row0, ps, pe, M.nslices, mc, M.slice_ptr, M.col, M.val, diag, b, u, ld,
                      flow.pf_slice0, flow.pf_slice_end);
      return;
    }
  }
  const int g = blocks_for(pe - row0, kBlock);
  if (use_staged(M)) {
    SMG_DISPATCH_K(k, launch_kernel("gs_phase", sell_gs_phase_kernel<K, true>, g, kBlock, stage_bytes(M), st,
                                    row0, ps, pe, M.nslices, M.max_chunk, M.slice_ptr, M.col, M.val,
                                    diag, b, u, ld, flow));
  } else {
    SMG_DISPATCH_K(k, launch_kernel("gs_phase", sell_gs_phase_kernel<K, false>, g, kBlock, 0, st, row0, ps, pe,
                                    M.nslices, M.max_chunk, M.slice_ptr, M.col, M.val, diag, b, u,
                                    ld, flow));
  }
}

// ---------------------------------------------------------------------------
// setup-time numeric kernels
// ---------------------------------------------------------------------------
namespace {

__global__ void gather_values_kernel(const double* __restrict__ in, const int* __restrict__ idx,
                                     double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

__global__ void fill_sell_kernel(const double* __restrict__ csc, const int* __restrict__ src,
                                 const int* __restrict__ tmap, double* __restrict__ val,
                                 double* __restrict__ valT, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = src[i];
  val[i] = s >= 0 ? csc[s] : 0.0;
  if (tmap) valT[i] = s >= 0 ? csc[tmap[s]] : 0.0;
}

__global__ void extract_diag_kernel(const double* __restrict__ csc,
                                    const int* __restrict__ diag_pos,
                                    const int* __restrict__ perm, double* __restrict__ diag,
                                    int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) diag[r] = csc[diag_pos[perm[r]]];
}

__global__ void shift_diag_kernel(double* csc, const int* __restrict__ diag_pos, int n,
                                  double shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) csc[diag_pos[i]] = __dadd_rn(csc[diag_pos[i]], shift);
}

// T1(i,j) = sum_{k in A(:,j), ascending} PT(i,k) * A(k,j), PT(i,k) = P(k,i) looked up
// in row k of P.  First touch assigns, later touches add (Eigen's conservative
// product), products and sums rounded separately.
__global__ void galerkin_t1_kernel(int nnz, const int* __restrict__ t_row,
                                   const int* __restrict__ t_col,
                                   const int* __restrict__ a_colptr,
                                   const int* __restrict__ a_rowidx,
                                   const double* __restrict__ a_val,
                                   const int* __restrict__ prow_ptr,
                                   const int* __restrict__ pcol, const double* __restrict__ pval,
                                   double* __restrict__ t_val) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int i = t_row[e], j = t_col[e];
  double acc = 0.0;
  bool first = true;
  for (int p = a_colptr[j]; p < a_colptr[j + 1]; p++) {
    const int k = a_rowidx[p];
    for (int q = prow_ptr[k]; q < prow_ptr[k + 1]; q++)
      if (pcol[q] == i) {
        const double t = __dmul_rn(pval[q], a_val[p]);
        acc = first ? t : __dadd_rn(acc, t);
        first = false;
        break;
      }
  }
  t_val[e] = acc;
}

// Ac(i,j) = sum_{k in P(:,j), ascending} T1(i,k) * P(k,j)
__global__ void galerkin_ac_kernel(int nnz, const int* __restrict__ c_row,
                                   const int* __restrict__ c_col,
                                   const int* __restrict__ p_colptr,
                                   const int* __restrict__ p_rowidx,
                                   const double* __restrict__ p_val,
                                   const int* __restrict__ t_colptr,
                                   const int* __restrict__ t_rowidx,
                                   const double* __restrict__ t_val,
                                   double* __restrict__ c_val) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  const int i = c_row[e], j = c_col[e];
  double acc = 0.0;
  bool first = true;
  for (int p = p_colptr[j]; p < p_colptr[j + 1]; p++) {
    const int k = p_rowidx[p];
    int lo = t_colptr[k], hi = t_colptr[k + 1];
    while (lo < hi) {  // binary search for row i in column k of T1
      const int mid = (lo + hi) >> 1;
      if (t_rowidx[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo < t_colptr[k + 1] && t_rowidx[lo] == i) {
      const double t = __dmul_rn(t_val[lo], p_val[p]);
      acc = first ? t : __dadd_rn(acc, t);
      first = false;
    }
  }
  c_val[e] = acc;
}

__global__ void csc_to_dense_kernel(int nnz, const int* __restrict__ rowidx,
                                    const int* __restrict__ colidx,
                                    const double* __restrict__ val,
                                    const int* __restrict__ iperm, double* __restrict__ D,
                                    int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  D[(size_t)iperm[rowidx[e]] + (size_t)iperm[colidx[e]] * n] = val[e];
}

__global__ void symmetrize_lower_kernel(double* D, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y * blockDim.y + threadIdx.y;
  if (r < n && c < n && r > c) D[(size_t)c + (size_t)r * n] = D[(size_t)r + (size_t)c * n];
}

// ---- coarse direct solve: u += Ainv * b with Ainv symmetric -----------------------
// Only the lower-triangular 64x64 tiles of the dense inverse are read (half the
// bytes): tile (I,J), I >= J, contributes  y_I += T b_J  and, for I != J,
// y_J += T^T b_I.  Every contribution to block-row X is written to its own slot
// (slot = index of the other block), and dense_sym_reduce_kernel sums the slots in a
// fixed order: deterministic, no floating-point atomics.
constexpr int kTile = 64;

template <int K>
__global__ void __launch_bounds__(256)
dense_sym_tile_kernel(int trace_slot, const double* __restrict__ tiles, const double* b,
                      double* __restrict__ partial, int n, int ldb) {
  trace_begin(trace_slot);
  pdl_launch_dependents();
  // triangular tile index -> (I, J), J <= I
  const int t = blockIdx.x;
  int I = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
  while ((I + 1) * (I + 2) / 2 <= t) I++;
  while (I * (I + 1) / 2 > t) I--;
  const int J = t - I * (I + 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = I * kTile + lane, r1 = r0 + 32;
  const int c0 = J * kTile + warp * 8;
  // The launch is programmatic: the first CTAs start long before the coarse right-hand
  // side exists (while the latency-bound coarse levels run and HBM idles).  They ask for
  // ALL tiles to be brought into L2 (TMA bulk prefetch), so the CTAs that only become
  // resident after the PDL wait stream from L2 instead of DRAM.
  constexpr int kPrefetchers = 512;
  if (threadIdx.x == 0 && t < kPrefetchers) {
    for (int tt = t + kPrefetchers; tt < (int)gridDim.x; tt += kPrefetchers) {
      const double* p = tiles + static_cast<size_t>(tt) * (kTile * kTile);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p),
                   "r"(static_cast<uint32_t>(kTile * kTile * sizeof(double)))
                   : "memory");
    }
  }
  double m0[8], m1[8];
  // tile t is one contiguous 32 KB block (column-major inside, zero beyond n); it is
  // immutable during a solve and fetched before the PDL wait
  const double* tp = tiles + static_cast<size_t>(t) * (kTile * kTile) + (warp * 8) * kTile + lane;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    m0[c] = ld_stream_f64(tp + c * kTile);
    m1[c] = ld_stream_f64(tp + c * kTile + 32);
  }
  pdl_wait();
  __shared__ double red[8][kTile];
#pragma unroll
  for (int q = 0; q < K; q++) {
    const double* bq = b + static_cast<size_t>(q) * ldb;
    // row part: y_I[r] += sum_c T[r][c] * b_J[c]
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const double bj = (c0 + c < n) ? ld_vec(bq + c0 + c) : 0.0;
      a0 += m0[c] * bj;
      a1 += m1[c] * bj;
    }
    red[warp][lane] = a0;
    red[warp][lane + 32] = a1;
    __syncthreads();
    if (threadIdx.x < kTile) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < 8; w++) sum += red[w][threadIdx.x];
      const int r = I * kTile + threadIdx.x;
      if (r < n) partial[(static_cast<size_t>(J) * K + q) * n + r] = sum;
    }
    __syncthreads();
    // column part: y_J[c] += sum_r T[r][c] * b_I[r]
    if (I != J) {
      const double bi0 = r0 < n ? ld_vec(bq + r0) : 0.0;
      const double bi1 = r1 < n ? ld_vec(bq + r1) : 0.0;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double v = m0[c] * bi0 + m1[c] * bi1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && c0 + c < n) partial[(static_cast<size_t>(I) * K + q) * n + c0 + c] = v;
      }
    }
  }
  trace_end(trace_slot);
}

// u[block-row X] += sum over the nblk slots, in slot order: 64 rows x 4 slot groups per CTA
template <int K>
__global__ void __launch_bounds__(256)
dense_sym_reduce_kernel(int trace_slot, const double* partial, double* u, int n, int nblk, int ldu) {
  trace_begin(trace_slot);
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[4][kTile];
  const int r = threadIdx.x & (kTile - 1), g = threadIdx.x >> 6;
  const int i = blockIdx.x * kTile + r;
#pragma unroll
  for (int q = 0; q < K; q++) {
    double sum = 0.0;
    if (i < n)
      for (int s0 = g; s0 < nblk; s0 += 32) {  // eight loads in flight, summed in slot order
        double v[8];
#pragma unroll
        for (int a = 0; a < 8; a++)
          v[a] = s0 + 4 * a < nblk ? ld_vec(partial + (static_cast<size_t>(s0 + 4 * a) * K + q) * n + i) : 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++) sum += v[a];
      }
    red[g][r] = sum;
    __syncthreads();
    if (g == 0 && i < n) {
      const double tot = ((red[0][r] + red[1][r]) + red[2][r]) + red[3][r];
      u[i + static_cast<size_t>(q) * ldu] = ld_vec(u + i + static_cast<size_t>(q) * ldu) + tot;
    }
    __syncthreads();
  }
  trace_end(trace_slot);
}

// pack the lower-triangular 64x64 tiles of a symmetric matrix whose LOWER triangle is valid
// (the output of syrk / potri) into contiguous tiles, mirroring inside the diagonal tiles
__global__ void __launch_bounds__(256)
pack_sym_tiles_kernel(const double* __restrict__ A, double* __restrict__ tiles, int n) {
  const int t = blockIdx.x;
  int I = static_cast<int>((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
  while ((I + 1) * (I + 2) / 2 <= t) I++;
  while (I * (I + 1) / 2 > t) I--;
  const int J = t - I * (I + 1) / 2;
  for (int e = threadIdx.x; e < kTile * kTile; e += 256) {
    const int c = e / kTile, r = e % kTile;
    int gr = I * kTile + r, gc = J * kTile + c;
    if (gr < gc) {  // only inside a diagonal tile: take the mirrored entry
      const int tmp = gr;
      gr = gc;
      gc = tmp;
    }
    tiles[static_cast<size_t>(t) * (kTile * kTile) + e] =
        (gr < n && gc < n) ? A[static_cast<size_t>(gc) * n + gr] : 0.0;
  }
}

__global__ void gather_system_kernel(const double* __restrict__ RHS,
                                     const double* __restrict__ z0,
                                     const double* __restrict__ kv, int n_full, int n_known,
                                     const int* __restrict__ g, const int* __restrict__ auk_ptr,
                                     const int* __restrict__ auk_q,
                                     const double* __restrict__ auk_val, double* __restrict__ bu,
                                     double* __restrict__ zu, int nu, int k) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nu) return;
  const int src = g[r];
  for (int q = 0; q < k; q++) {
    zu[r + (size_t)q * nu] = z0[src + (size_t)q * n_full];
    double rhs = RHS[src + (size_t)q * n_full];
    if (auk_ptr) {
      // RHS_unknown - (Auk * known_val): the product accumulates from zero in
      // ascending known-column order (cpp:316-318)
      double t = 0.0;
      for (int p = auk_ptr[r]; p < auk_ptr[r + 1]; p++)
        t = __dadd_rn(t, __dmul_rn(auk_val[p], kv[auk_q[p] + (size_t)q * n_known]));
      rhs = __dsub_rn(rhs, t);
    }
    bu[r + (size_t)q * nu] = rhs;
  }
}

__global__ void scatter_solution_kernel(const double* __restrict__ zu, const int* __restrict__ g,
                                        double* __restrict__ z, int n_full, int nu, int k) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nu) return;
  const int dst = g[r];
  for (int q = 0; q < k; q++) z[dst + (size_t)q * n_full] = zu[r + (size_t)q * nu];
}

__global__ void scatter_known_kernel(const double* __restrict__ kv, const int* __restrict__ kidx,
                                     const int* __restrict__ ksrc, double* __restrict__ z,
                                     int n_full, int n_known, int n_distinct, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_distinct) return;
  for (int q = 0; q < k; q++)
    z[kidx[i] + (size_t)q * n_full] = kv[ksrc[i] + (size_t)q * n_known];
}

__global__ void permute_in_kernel(const double* __restrict__ in, const int* __restrict__ perm,
                                  double* __restrict__ out, int n, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = perm[i];
  for (int q = 0; q < k; q++) out[i + (size_t)q * n] = in[s + (size_t)q * n];
}

__global__ void permute_out_kernel(const double* __restrict__ in, const int* __restrict__ perm,
                                   double* __restrict__ out, int n, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = perm[i];
  for (int q = 0; q < k; q++) out[d + (size_t)q * n] = in[i + (size_t)q * n];
}

__global__ void fill_kernel(int trace_slot, double* p, double v, int64_t n) {
  trace_begin(trace_slot);
  pdl_launch_dependents();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
  trace_end(trace_slot);
}

}  // namespace

void launch_gather_values(const double* in, const int* idx, double* out, int n, cudaStream_t st) {
  if (n > 0) gather_values_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, idx, out, n);
}
void launch_fill_sell(const double* csc, const int* src, const int* tmap, double* val,
                      double* valT, int64_t n, cudaStream_t st) {
  if (n > 0) fill_sell_kernel<<<blocks_for(n, 256), 256, 0, st>>>(csc, src, tmap, val, valT, n);
}
void launch_extract_diag(const double* csc, const int* diag_pos, const int* perm, double* diag,
                         int n, cudaStream_t st) {
  if (n > 0) extract_diag_kernel<<<blocks_for(n, 256), 256, 0, st>>>(csc, diag_pos, perm, diag, n);
}
void launch_shift_diag(double* csc, const int* diag_pos, int n, double shift, cudaStream_t st) {
  if (n > 0) shift_diag_kernel<<<blocks_for(n, 256), 256, 0, st>>>(csc, diag_pos, n, shift);
}
void launch_galerkin_t1(int nnz_t1, const int* t_row, const int* t_col, const int* a_colptr,
                        const int* a_rowidx, const double* a_val, const int* prow_ptr,
                        const int* pcol, const double* pval, double* t_val, cudaStream_t st) {
  if (nnz_t1 > 0)
    galerkin_t1_kernel<<<blocks_for(nnz_t1, 256), 256, 0, st>>>(
        nnz_t1, t_row, t_col, a_colptr, a_rowidx, a_val, prow_ptr, pcol, pval, t_val);
}
void launch_galerkin_ac(int nnz_ac, const int* c_row, const int* c_col, const int* p_colptr,
                        const int* p_rowidx, const double* p_val, const int* t_colptr,
                        const int* t_rowidx, const double* t_val, double* c_val,
                        cudaStream_t st) {
  if (nnz_ac > 0)
    galerkin_ac_kernel<<<blocks_for(nnz_ac, 256), 256, 0, st>>>(
        nnz_ac, c_row, c_col, p_colptr, p_rowidx, p_val, t_colptr, t_rowidx, t_val, c_val);
}
void launch_csc_to_dense(int nnz, const int* rowidx, const int* colidx, const double* val,
                         const int* iperm, double* D, int n, cudaStream_t st) {
  if (nnz > 0)
    csc_to_dense_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, rowidx, colidx, val, iperm, D, n);
}
namespace {
__global__ void set_identity_kernel(double* D, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) D[(size_t)i * n + i] = 1.0;
}
}  // namespace
void launch_set_identity(double* D, int n, cudaStream_t st) {
  if (n > 0) set_identity_kernel<<<blocks_for(n, 256), 256, 0, st>>>(D, n);
}
void launch_symmetrize_lower(double* D, int n, cudaStream_t st) {
  if (n <= 0) return;
  dim3 b(32, 8), g((n + 31) / 32, (n + 7) / 8);
  symmetrize_lower_kernel<<<g, b, 0, st>>>(D, n);
}
size_t dense_sym_scratch_doubles(int n, int k) {
  const int nblk = (n + kTile - 1) / kTile;
  return static_cast<size_t>(nblk) * (k < kMaxK ? k : kMaxK) * n;
}
void launch_dense_sym_add(const double* Ainv, const double* b, double* u, double* scratch, int n,
                          int k, cudaStream_t st) {
  if (n <= 0) return;
  const int nblk = (n + kTile - 1) / kTile;
  const int ntiles = nblk * (nblk + 1) / 2;
  SMG_DISPATCH_K(k, launch_kernel("coarse_tiles", dense_sym_tile_kernel<K>, ntiles, 256, 0, st, Ainv,
                                  b, scratch, n, n));
  SMG_DISPATCH_K(k, launch_kernel("coarse_reduce", dense_sym_reduce_kernel<K>, nblk, 256, 0, st,
                                  scratch, u, n, nblk, n));
}
size_t dense_sym_tiles_doubles(int n) {
  const size_t nblk = (n + kTile - 1) / kTile;
  return nblk * (nblk + 1) / 2 * kTile * kTile;
}
void launch_pack_sym_tiles(const double* A_lower, double* tiles, int n, cudaStream_t st) {
  if (n <= 0) return;
  const int nblk = (n + kTile - 1) / kTile;
  pack_sym_tiles_kernel<<<nblk * (nblk + 1) / 2, 256, 0, st>>>(A_lower, tiles, n);
}
void launch_gather_system(const double* RHS, const double* z0, const double* kv, int n_full,
                          int n_known, const int* g, const int* auk_ptr, const int* auk_q,
                          const double* auk_val, double* bu, double* zu, int nu, int k,
                          cudaStream_t st) {
  if (nu > 0)
    gather_system_kernel<<<blocks_for(nu, 256), 256, 0, st>>>(
        RHS, z0, kv, n_full, n_known, g, auk_ptr, auk_q, auk_val, bu, zu, nu, k);
}
void launch_scatter_solution(const double* zu, const int* g, double* z, int n_full, int nu, int k,
                             cudaStream_t st) {
  if (nu > 0) scatter_solution_kernel<<<blocks_for(nu, 256), 256, 0, st>>>(zu, g, z, n_full, nu, k);
}
void launch_scatter_known(const double* kv, const int* kidx, const int* ksrc, double* z,
                          int n_full, int n_known, int n_distinct, int k, cudaStream_t st) {
  if (n_distinct > 0)
    scatter_known_kernel<<<blocks_for(n_distinct, 256), 256, 0, st>>>(kv, kidx, ksrc, z, n_full,
                                                                     n_known, n_distinct, k);
}
void launch_permute_in(const double* in, const int* perm, double* out, int n, int k,
                       cudaStream_t st) {
  if (n > 0) permute_in_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, perm, out, n, k);
}
void launch_permute_out(const double* in, const int* perm, double* out, int n, int k,
                        cudaStream_t st) {
  if (n > 0) permute_out_kernel<<<blocks_for(n, 256), 256, 0, st>>>(in, perm, out, n, k);
}
// Force the (lazily loaded) solve-time kernels into the context now.  Loading a kernel may
// synchronise the whole context; when ranks share a device that must not happen while
// another rank already spins in a halo exchange.
namespace {
template <class F>
void preload_one(F* f) {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, f);
}
template <int K>
void preload_k() {
  preload_one(sell_apply_kernel<K, MODE_SPMV, true>);
  preload_one(sell_apply_kernel<K, MODE_SPMV, false>);
  preload_one(sell_apply_kernel<K, MODE_RESIDUAL, true>);
  preload_one(sell_apply_kernel<K, MODE_RESIDUAL, false>);
  preload_one(sell_apply_kernel<K, MODE_ADD, true>);
  preload_one(sell_apply_kernel<K, MODE_ADD, false>);
  preload_one(sell_apply_kernel<K, MODE_SPMV_ZERO, true>);
  preload_one(sell_apply_kernel<K, MODE_SPMV_ZERO, false>);
  preload_one(sell_apply_short_kernel<K, MODE_SPMV, 4, 3>);
  preload_one(sell_apply_short_kernel<K, MODE_RESIDUAL, 4, 3>);
  preload_one(sell_apply_short_kernel<K, MODE_ADD, 4, 3>);
  preload_one(sell_apply_short_kernel<K, MODE_SPMV_ZERO, 4, 3>);
  preload_one(sell_spmv_ranges_kernel<K, true>);
  preload_one(sell_spmv_ranges_kernel<K, false>);
  preload_one(sell_residual_norm_kernel<K, true>);
  preload_one(sell_residual_norm_kernel<K, false>);
  preload_one(sell_gs_phase_kernel<K, true>);
  preload_one(sell_gs_phase_kernel<K, false>);
  preload_one(dense_sym_tile_kernel<K>);
  preload_one(dense_sym_reduce_kernel<K>);
}
}  // namespace

void preload_kernels() {
  preload_k<1>();
  preload_k<2>();
  preload_k<3>();
  preload_k<4>();
  if (!g_gs_attr_set) set_gs_rows(g_gs_rows);
  preload_one(sell_gs_phase_multi_kernel<1, 2>);
  preload_one(sell_gs_phase_multi_kernel<1, 4>);
  preload_one(sell_apply_short_kernel<1, MODE_SPMV, 2, kPre>);
  preload_one(sell_apply_short_kernel<1, MODE_RESIDUAL, 2, kPre>);
  preload_one(sell_apply_short_kernel<1, MODE_ADD, 2, kPre>);
  preload_one(sell_apply_short_kernel<1, MODE_SPMV_ZERO, 2, kPre>);
  preload_one(sell_apply_short_kernel<1, MODE_NORM, 2, kPre>);
  preload_one(halo_exchange_kernel);
  preload_one(fill_kernel);
  preload_one(gather_system_kernel);
  preload_one(scatter_solution_kernel);
  preload_one(scatter_known_kernel);
  preload_one(permute_in_kernel);
  preload_one(permute_out_kernel);
  cudaGetLastError();
}

// ---- mean-curvature-flow assembly on the device (mcf_core.hpp holds the arithmetic) ------
namespace {
__global__ void mcf_face_kernel(int nV, int nF, const int* __restrict__ F, const double* __restrict__ U,
                                double* __restrict__ dblA) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nF) dblA[f] = mcf_face_doublearea(U, nV, F[f], F[f + nF], F[f + 2 * nF]);
}
__global__ void mcf_vertex_kernel(int nV, const int* __restrict__ vf_ptr, const int* __restrict__ vf_face,
                                  const double* __restrict__ dblA, double* __restrict__ mass) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nV) mass[v] = mcf_vertex_mass(dblA, vf_face, vf_ptr[v], vf_ptr[v + 1]);
}
__global__ void mcf_lhs_kernel(int nnz, const int* __restrict__ rowidx, const int* __restrict__ colidx,
                               const double* __restrict__ mass, double delta, const double* __restrict__ Lval,
                               double* __restrict__ a_val) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) a_val[p] = mcf_lhs_entry(rowidx[p] == colidx[p] ? mass[colidx[p]] : 0.0, delta, Lval[p]);
}
__global__ void mcf_rhs_kernel(int nV, int k, const double* __restrict__ mass, const double* __restrict__ U,
                               double* __restrict__ rhs) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nV)
    for (int c = 0; c < k; c++) rhs[v + (size_t)c * nV] = __dmul_rn(mass[v], U[v + (size_t)c * nV]);
}
}  // namespace

void launch_mcf_assemble(int nV, int nF, const int* F, const double* U, const int* vf_ptr, const int* vf_face,
                         double* dblA, double* mass, int nnz, const int* rowidx, const int* colidx, double delta,
                         const double* Lval, double* a_val, int k, double* rhs, cudaStream_t st) {
  if (nF > 0) mcf_face_kernel<<<blocks_for(nF, 256), 256, 0, st>>>(nV, nF, F, U, dblA);
  if (nV > 0) mcf_vertex_kernel<<<blocks_for(nV, 256), 256, 0, st>>>(nV, vf_ptr, vf_face, dblA, mass);
  if (nnz > 0) mcf_lhs_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, rowidx, colidx, mass, delta, Lval, a_val);
  if (nV > 0) mcf_rhs_kernel<<<blocks_for(nV, 256), 256, 0, st>>>(nV, k, mass, U, rhs);
}

void launch_fill(double* p, double v, int64_t n, cudaStream_t st) {
  if (n > 0) launch_kernel("fill", fill_kernel, blocks_for(n, 256), 256, 0, st, p, v, n);
}

// ---- communication-avoiding patch smoother (patch.hpp) -----------------------------------
// One CTA per patch.  The patch blob (header, local matrix in ELL form, index lists) is
// immutable during a solve: ONE cp.async.bulk brings it into shared memory before the PDL
// wait, i.e. while the previous kernel still runs.  After the wait the CTA gathers u (and b)
// of its local rows, runs all colour phases of the relax call out of shared memory
// (__syncthreads between colours instead of kernel boundaries), and writes the rows it owns.
//   PATCH_DOWN: u_out = relax(u_in); r = b - A u_out on the rows its coarse rows restrict
//               from; bc = PT r and uc = 0 on those coarse rows    (mg_VCycle.cpp:36-47)
//   PATCH_UP:   u_out = relax(u_in + P uc)                           (mg_VCycle.cpp:52-56)
// u_in and u_out are different buffers: other CTAs still read the old values of rows this CTA
// owns.  Arithmetic per row is that of sell_gs_phase_kernel / sell_apply_* (same entry order,
// products and sums rounded separately, true division), so results are bit-identical.
namespace {
// sum[q] += sum_j val[j * stride + r] * x[col[j * stride + r] + q * ldx] over the first w entries of
// ELL row r, in entry order, products and sums rounded separately.  The first WMAX entries are
// fetched together (indices, then values and gathers) before the dependent chain of additions.
template <int K, int WMAX, class ColT>
__device__ __forceinline__ void ell_row_sum(const ColT* __restrict__ col, const double* __restrict__ val, int stride,
                                            int r, int w, const double* x, int ldx, double (&sum)[K]) {
  int c[WMAX];
  double v[WMAX], xv[WMAX][K];
#pragma unroll
  for (int j = 0; j < WMAX; j++)
    if (j < w) {
      c[j] = col[j * stride + r];
      v[j] = val[j * stride + r];
    }
#pragma unroll
  for (int j = 0; j < WMAX; j++)
    if (j < w) {
#pragma unroll
      for (int q = 0; q < K; q++) xv[j][q] = x[c[j] + q * ldx];
    }
#pragma unroll
  for (int j = 0; j < WMAX; j++)
    if (j < w) {
#pragma unroll
      for (int q = 0; q < K; q++) sum[q] = __dadd_rn(sum[q], __dmul_rn(v[j], xv[j][q]));
    }
  for (int j = WMAX; j < w; j++) {
    const int cc = col[j * stride + r];
    const double vv = val[j * stride + r];
#pragma unroll
    for (int q = 0; q < K; q++) sum[q] = __dadd_rn(sum[q], __dmul_rn(vv, x[cc + q * ldx]));
  }
}

template <int K, int KIND>
__global__ void __launch_bounds__(kPatchThreads)
patch_kernel(int trace_slot, const unsigned char* __restrict__ blob, const long long* __restrict__ off,
             const double* u_in, double* u_out, const double* b, int ld, const double* uc, double* bc,
             double* uc_zero, int ldc, const unsigned char* __restrict__ pf_blob,
             const long long* __restrict__ pf_off, int pf_n, int use_tma) {
  extern __shared__ __align__(128) unsigned char dyn[];
  __shared__ uint64_t bar;
  trace_begin(trace_slot);
  trace_begin(trace_slot < 0 ? -1 : trace_slot + 1);
  pdl_launch_dependents();
  const long long o0 = off[blockIdx.x];
  const uint32_t bytes = static_cast<uint32_t>(off[blockIdx.x + 1] - o0);
  if (use_tma) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_expect_tx(&bar, bytes);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dyn)),
          "l"(blob + o0), "r"(bytes), "r"(smem_u32(&bar))
          : "memory");
    }
  } else {  // SMG_NO_TMA=1 (tools that do not model bulk copies): plain cooperative copy
    const uint4* src = reinterpret_cast<const uint4*>(blob + o0);
    uint4* dst = reinterpret_cast<uint4*>(dyn);
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (use_tma) mbar_wait(&bar, 0);
  trace_end(trace_slot < 0 ? -1 : trace_slot + 1);  // sub-slots (tracing only): blob in shared memory
  const PatchHeader& H = *reinterpret_cast<const PatchHeader*>(dyn);
  const int n_loc = H.n_loc, n_b = H.n_b;
  const int nthr = blockDim.x, tid = threadIdx.x;
  double* u_loc = reinterpret_cast<double*>(dyn + ((bytes + 15u) & ~15u));
  double* b_loc = u_loc + static_cast<size_t>(K) * n_loc;
  double* r_loc = b_loc + static_cast<size_t>(K) * n_b;
  const int* gid = reinterpret_cast<const int*>(dyn + H.o_gid);
  pdl_wait();
  trace_begin(trace_slot < 0 ? -1 : trace_slot + 2);  // gather
  // ---- gather the local rows ----------------------------------------------------------
  // All global loads of a batch of U rows are issued before the first one is used: a patch
  // is a few rows per thread, and one dependent round trip per row would dominate the launch.
  constexpr int U = K == 1 ? 8 : (K == 2 ? 4 : 2);
  if (KIND == PATCH_UP) {
    constexpr int WP = 3;  // prolongation rows hold at most three entries (get_prolong.cpp:45-56)
    const unsigned char* pw = dyn + H.o_pw;
    const int* pcol = reinterpret_cast<const int*>(dyn + H.o_pcol);
    const double* pval = reinterpret_cast<const double*>(dyn + H.o_pval);
    for (int i0 = tid; i0 < n_loc; i0 += nthr * U) {
      double y[U][K], x[U][WP][K], bv[U][K];
      int w[U];
#pragma unroll
      for (int a = 0; a < U; a++) {
        const int i = i0 + a * nthr;
        w[a] = 0;
        if (i < n_loc) {
          const int g = gid[i];
          w[a] = pw[i];
#pragma unroll
          for (int q = 0; q < K; q++) y[a][q] = ld_vec(u_in + g + (size_t)q * ld);
#pragma unroll
          for (int j = 0; j < WP; j++)
            if (j < w[a]) {
              const int c = pcol[(size_t)j * n_loc + i];
#pragma unroll
              for (int q = 0; q < K; q++) x[a][j][q] = ld_vec(uc + c + (size_t)q * ldc);
            }
          if (i < n_b) {
#pragma unroll
            for (int q = 0; q < K; q++) bv[a][q] = ld_vec(b + g + (size_t)q * ld);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < U; a++) {
        const int i = i0 + a * nthr;
        if (i >= n_loc) continue;
        double sum[K];
#pragma unroll
        for (int q = 0; q < K; q++) sum[q] = 0.0;
#pragma unroll
        for (int j = 0; j < WP; j++)
          if (j < w[a]) {
            const double v = pval[(size_t)j * n_loc + i];
#pragma unroll
            for (int q = 0; q < K; q++) sum[q] = __dadd_rn(sum[q], __dmul_rn(v, x[a][j][q]));
          }
        for (int j = WP; j < w[a]; j++) {  // wider rows (general P): one at a time
          const int c = pcol[(size_t)j * n_loc + i];
          const double v = pval[(size_t)j * n_loc + i];
#pragma unroll
          for (int q = 0; q < K; q++) sum[q] = __dadd_rn(sum[q], __dmul_rn(v, ld_vec(uc + c + (size_t)q * ldc)));
        }
#pragma unroll
        for (int q = 0; q < K; q++) u_loc[i + q * n_loc] = __dadd_rn(y[a][q], sum[q]);
        if (i < n_b) {
#pragma unroll
          for (int q = 0; q < K; q++) b_loc[i + q * n_b] = bv[a][q];
        }
      }
    }
  } else {
    for (int i0 = tid; i0 < n_loc; i0 += nthr * U) {
      double y[U][K], bv[U][K];
#pragma unroll
      for (int a = 0; a < U; a++) {
        const int i = i0 + a * nthr;
        if (i < n_loc) {
          const int g = gid[i];
#pragma unroll
          for (int q = 0; q < K; q++) y[a][q] = ld_vec(u_in + g + (size_t)q * ld);
          if (i < n_b) {
#pragma unroll
            for (int q = 0; q < K; q++) bv[a][q] = ld_vec(b + g + (size_t)q * ld);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < U; a++) {
        const int i = i0 + a * nthr;
        if (i >= n_loc) continue;
#pragma unroll
        for (int q = 0; q < K; q++) u_loc[i + q * n_loc] = y[a][q];
        if (i < n_b) {
#pragma unroll
          for (int q = 0; q < K; q++) b_loc[i + q * n_b] = bv[a][q];
        }
      }
    }
  }
  __syncthreads();
  trace_end(trace_slot < 0 ? -1 : trace_slot + 2);
  trace_begin(trace_slot < 0 ? -1 : trace_slot + 3);  // colour phases
  // ---- colour phases ------------------------------------------------------------------
  {
    const unsigned char* wv = dyn + H.o_w;
    const unsigned short* col = reinterpret_cast<const unsigned short*>(dyn + H.o_col);
    const double* val = reinterpret_cast<const double*>(dyn + H.o_val);
    const double* diag = reinterpret_cast<const double*>(dyn + H.o_diag);
    const int T = H.T, C = H.C;
    int g = 0;
    for (int t = 0; t < T; t++) {
      const int na = H.n_active[t];
      const int g0 = H.grp_row[g], gn = H.grp_row[g + 1] - g0, e0 = H.grp_ent[g];
      for (int r = tid; r < na; r += nthr) {
        const int row = g0 + r;
        const int w = wv[row];
        double sum[K];
#pragma unroll
        for (int q = 0; q < K; q++) sum[q] = 0.0;
        ell_row_sum<K, 8>(col + e0, val + e0, gn, r, w, u_loc, n_loc, sum);
        const double d = diag[row];
#pragma unroll
        for (int q = 0; q < K; q++) u_loc[row + q * n_loc] = __ddiv_rn(__dsub_rn(b_loc[row + q * n_b], sum[q]), d);
      }
      __syncthreads();
      if (++g == C) g = 0;
    }
  }
  trace_end(trace_slot < 0 ? -1 : trace_slot + 3);
  trace_begin(trace_slot < 0 ? -1 : trace_slot + 4);  // write-back, residual, restriction
  // ---- the rows this patch owns --------------------------------------------------------
  {
    const unsigned short* own = reinterpret_cast<const unsigned short*>(dyn + H.o_own);
    for (int i = tid; i < H.n_own; i += nthr) {
      const int li = own[i];
      const int g = gid[li];
#pragma unroll
      for (int q = 0; q < K; q++) u_out[g + (size_t)q * ld] = u_loc[li + q * n_loc];
    }
  }
  if (KIND == PATCH_DOWN) {
    // residual of the rows the owned coarse rows restrict from, then the restriction
    const int n_R = H.n_R, n_C = H.n_C;
    const unsigned short* ridx = reinterpret_cast<const unsigned short*>(dyn + H.o_ridx);
    const unsigned char* rw = dyn + H.o_rw;
    const unsigned short* rcol = reinterpret_cast<const unsigned short*>(dyn + H.o_rcol);
    const double* rval = reinterpret_cast<const double*>(dyn + H.o_rval);
    for (int r = tid; r < n_R; r += nthr) {
      const int w = rw[r];
      double sum[K];
#pragma unroll
      for (int q = 0; q < K; q++) sum[q] = 0.0;
      ell_row_sum<K, 8>(rcol, rval, n_R, r, w, u_loc, n_loc, sum);
      const int li = ridx[r];
#pragma unroll
      for (int q = 0; q < K; q++) r_loc[r + q * n_R] = __dsub_rn(b_loc[li + q * n_b], sum[q]);
    }
    __syncthreads();
    const int* cgid = reinterpret_cast<const int*>(dyn + H.o_cgid);
    const unsigned char* ptw = dyn + H.o_ptw;
    const unsigned short* ptcol = reinterpret_cast<const unsigned short*>(dyn + H.o_ptcol);
    const double* ptval = reinterpret_cast<const double*>(dyn + H.o_ptval);
    for (int r = tid; r < n_C; r += nthr) {
      const int w = ptw[r];
      double sum[K];
#pragma unroll
      for (int q = 0; q < K; q++) sum[q] = 0.0;
      ell_row_sum<K, 8>(ptcol, ptval, n_C, r, w, r_loc, n_R, sum);
      const int I = cgid[r];
#pragma unroll
      for (int q = 0; q < K; q++) {
        bc[I + (size_t)q * ldc] = sum[q];
        uc_zero[I + (size_t)q * ldc] = 0.0;
      }
    }
  }
  // The blobs of the NEXT patch launch of the V-cycle -> L2 (immutable data; one 32 KB piece per
  // thread).  Last statement of the kernel: ptxas serialises the uniform-datapath prefetch
  // instruction over the lanes, which compute-sanitizer's synccheck misreads as divergence at
  // any CTA barrier that follows it.
  for (int p = blockIdx.x; p < pf_n; p += gridDim.x) {
    const long long q0 = pf_off[p], q1 = pf_off[p + 1];
    for (long long q = q0 + 32768ll * threadIdx.x; q < q1; q += 32768ll * blockDim.x) {
      const uint32_t nb = static_cast<uint32_t>(q1 - q < 32768 ? q1 - q : 32768);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_blob + q), "r"(nb) : "memory");
    }
  }
  trace_end(trace_slot < 0 ? -1 : trace_slot + 4);
  trace_end(trace_slot);
}

// device-side numeric fill of the matrix values of the patch blobs (numeric precompute)
__global__ void patch_fill_kernel(double* __restrict__ blob, const int* __restrict__ dst,
                                  const int* __restrict__ src, const double* __restrict__ csc, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) blob[dst[i]] = csc[src[i]];
}

template <int K, int KIND>
void patch_set_attr(size_t smem) {
  static size_t cur[64] = {0};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > cur[dev]) {
    cudaFuncSetAttribute(patch_kernel<K, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cur[dev] = smem;
  }
}
}  // namespace

size_t patch_smem_bytes(const PatchDev& P, int k) {
  return static_cast<size_t>((P.max_blob_bytes + 15) & ~15) + static_cast<size_t>(k) * 8 * P.max_vec_doubles + 16;
}

void launch_patch_fill(double* blob, const int* dst, const int* src, const double* csc, int64_t n,
                       cudaStream_t st) {
  if (n > 0) patch_fill_kernel<<<blocks_for(n, 256), 256, 0, st>>>(blob, dst, src, csc, n);
}

void launch_patch(const PatchDev& P, int kind, const double* u_in, double* u_out, const double* b, int ld,
                  const double* uc, double* bc, double* uc_zero, int ldc, int k, const PatchDev* next,
                  cudaStream_t st) {
  if (P.n_patches <= 0) return;
  const size_t smem = patch_smem_bytes(P, k);
  const int threads = P.max_active > 256 ? kPatchThreads : 256;
  const char* name = kind == PATCH_DOWN ? "patch_down" : "patch_up";
  const unsigned char* pf_blob = next ? next->blob : nullptr;
  const long long* pf_off = next ? next->off : nullptr;
  const int pf_n = next ? next->n_patches : 0;
  static const char* const kStages[4] = {".blob", ".gather", ".phases", ".tail"};
  g_trace_sub = 4;
  g_trace_sub_names = kStages;
  SMG_DISPATCH_K(k, {
    if (kind == PATCH_DOWN) {
      patch_set_attr<K, PATCH_DOWN>(smem);
      launch_kernel(name, patch_kernel<K, PATCH_DOWN>, P.n_patches, threads, smem, st, P.blob, P.off, u_in, u_out, b,
                    ld, uc, bc, uc_zero, ldc, pf_blob, pf_off, pf_n, g_use_tma ? 1 : 0);
    } else {
      patch_set_attr<K, PATCH_UP>(smem);
      launch_kernel(name, patch_kernel<K, PATCH_UP>, P.n_patches, threads, smem, st, P.blob, P.off, u_in, u_out, b,
                    ld, uc, bc, uc_zero, ldc, pf_blob, pf_off, pf_n, g_use_tma ? 1 : 0);
    }
  });
}


namespace {
__global__ void solve_decide_kernel(cudaGraphConditionalHandle handle, SolveCtl* ctl, const double* norm2,
                                    int nchunks, int in_body) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (in_body && ctl->n_his >= ctl->max_iter) {
    cudaGraphSetConditional(handle, 0);
    return;
  }
  double ss = 0.0;
  for (int c = 0; c < nchunks; c++) ss += ld_vec(norm2 + c);
  const double r = sqrt(ss);
  ctl->r_his[ctl->n_his++] = r;
  const bool finite = isfinite(r);
  if (!finite) ctl->nonfinite = 1;
  cudaGraphSetConditional(handle, finite && !(r < ctl->tol) ? 1u : 0u);
}
}  // namespace

void launch_solve_decide(cudaGraphConditionalHandle handle, SolveCtl* ctl, const double* norm2, int nchunks,
                         int in_body, cudaStream_t st) {
  solve_decide_kernel<<<1, 32, 0, st>>>(handle, ctl, norm2, nchunks, in_body);
}

}  // namespace smg

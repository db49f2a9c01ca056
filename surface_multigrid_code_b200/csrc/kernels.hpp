// kernels.hpp -- launch wrappers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace smg {

// Device view of a SELL-32 matrix (see plan.hpp::Sell). `val` holds, for row i,
// the entries of CSC column i (what the reference's smoother reads); `valT` holds
// the true row i (entries A(i,j)), which is what A*x needs. For exactly symmetric
// matrices the two arrays have identical contents.
struct SellDev {
  int nrows = 0;
  int nslices = 0;
  int max_chunk = 0;  // most stored entries in any 8 consecutive slices (one CTA's chunk)
  int max_width = 0;  // widest slice (entries per row)
  int max_chunk32 = 0;  // most stored entries in any aligned group of 32 slices
  int max_chunk16 = 0, max_chunk32s = 0;  // ... in any 16 / 32 consecutive slices
  const int* slice_ptr = nullptr;
  const int* col = nullptr;
  const double* val = nullptr;
  const double* valT = nullptr;
  // rows [rb, re) are processed (re < 0: all rows); a rank of a row-partitioned level
  // applies only its own rows of the full matrix
  int rb = 0, re = -1;
  int row_begin() const { return rb; }
  int row_end() const { return re < 0 ? nrows : re; }
  SellDev rows(int b, int e) const {
    SellDev d = *this;
    d.rb = b;
    d.re = e;
    return d;
  }
};

constexpr int kMaxK = 4;  // right-hand sides handled per kernel pass

// up to kMaxRanges row ranges handled by one launch (launch_spmv_ranges)
constexpr int kMaxRanges = 8;
struct RowRanges {
  int n = 0;
  int rb[kMaxRanges] = {0}, re[kMaxRanges] = {0};
  int blk0[kMaxRanges + 1] = {0};  // filled by the launcher
};

// y = M x  (y: nrows x k, ldy; x: ldx)
void launch_spmv(const SellDev& M, bool use_valT, const double* x, int ldx, double* y, int ldy,
                 int k, cudaStream_t st);
// y = M x on the rows of `ranges` only, one launch
void launch_spmv_ranges(const SellDev& M, const RowRanges& ranges, const double* x, int ldx, double* y,
                        int ldy, int k, cudaStream_t st);
// y = M x and z = 0 (z has the shape of y): restriction fused with zeroing the coarse guess
void launch_spmv_zero(const SellDev& M, const double* x, int ldx, double* y, double* z, int ldy,
                      int k, cudaStream_t st);
// programmatic dependent launch of the hot-path kernels (default on)
void set_pdl_enabled(bool on);
// TMA (cp.async.bulk) staging of the matrix chunks in shared memory (default on)
void set_tma_enabled(bool on);
// rows from which the residual / norm / restriction kernels take two rows per thread (default 600000)
void set_apply2_rows(int rows);
// rows of a colour phase from which the Gauss-Seidel kernel takes two rows per thread (default 150000)
void set_multi_rows(int rows);
// rows per thread of the Gauss-Seidel phase kernel on large phases (1, 2 or 4)
void set_gs_rows(int r);
// in-kernel timeline: slots of 2 x u64 [min start, max end] in nanoseconds (%globaltimer);
// the caller initialises the buffer to {~0, 0} per slot
void trace_start(unsigned long long* dev_buf, int cap);
void trace_stop();
void trace_label(const char* label);
int trace_count();
const char* trace_name(int i);
// r = b - M x
void launch_residual(const SellDev& M, const double* b, const double* x, double* r, int ld, int k,
                     cudaStream_t st);
// u = u + M x   (u: nrows x k, ldu; x: ldx)
void launch_prolong_add(const SellDev& M, const double* x, int ldx, double* u, int ldu, int k,
                        cudaStream_t st);
// *out = || b - M x ||_F^2, deterministic; scratch must hold >= residual_norm_blocks(nrows) doubles
int residual_norm_blocks(int nrows);
// counter: one zero-initialised word (the last CTA to finish does the final sum and resets it)
void launch_residual_norm2(const SellDev& M, const double* b, const double* x, int ld, int k,
                           double* scratch, unsigned int* counter, double* out, cudaStream_t st);
// Per-launch extras of a Gauss-Seidel phase kernel.
struct GsFlow {
  // L2 prefetch of the NEXT launch's matrix chunk: CTA b asks for the slices
  // [pf_slice0 + 8 b, pf_slice0 + 8 b + 8) clipped to pf_slice_end; pf_slice0 < 0: none
  int pf_slice0 = -1, pf_slice_end = 0;
};
// one Gauss-Seidel phase: rows [ps, pe) of the permuted matrix, in place on u
void launch_gs_phase(const SellDev& M, const double* diag, const double* b, double* u, int ld,
                     int k, int ps, int pe, const GsFlow& flow, cudaStream_t st);

// ---- multi-GPU halo exchange over peer-mapped memory ------------------------------------
// One (exchange, peer) pair.  The sender gathers vec[send_idx[i]] and stores it as two 8-byte
// words {half of the value, epoch} into the PEER's staging slot (stores over NVLink); the receiver
// polls the words of its own slot until both carry the epoch and scatters the values into
// vec[recv_idx[i]].  A slot therefore holds (k * n + 1) pairs of words (16 bytes per value; the
// last pair is the per-pair sync word).
// Staging is double-buffered by epoch parity: a rank can be at most one exchange ahead of
// a peer, because it cannot leave exchange e before the peer has entered it.
struct XchgPeer {
  int n_send = 0, n_recv = 0;
  const int* send_idx = nullptr;
  const int* recv_idx = nullptr;
  double* remote_slot = nullptr;       // peer memory: parity 0; parity 1 at + parity_stride
  const double* local_slot = nullptr;  // own memory, written by the peer
};
constexpr int kXchgThreads = 512;
// ints of control memory (zero-initialised) an exchange context needs
int xchg_ctrl_ints();
int xchg_max_peers();
// how long an exchange waits for a peer before it gives up and poisons the context (20 s)
void set_xchg_timeout_ms(long long ms);
// vec (ld, k columns) is both source and destination; ctrl[2] != 0 after a wait timed out.
// ctas_per_peer in 1..32 may differ from exchange to exchange.  phases: 3 = push and
// receive in one launch (normal); 1 = push only, 2 = receive only (host-synchronised mode).
void launch_halo_exchange(const XchgPeer* d_peers, int npeers, int ctas_per_peer, double* vec, int ld,
                          int k, size_t parity_stride, int* ctrl, bool late_trigger, int phases,
                          cudaStream_t st);

// load every solve-time kernel into the current context (see kernels.cu)
void preload_kernels();

// ---- setup-time numeric kernels ---------------------------------------------
// out[i] = in[idx[i]]
void launch_gather_values(const double* in, const int* idx, double* out, int n, cudaStream_t st);
// SELL values from CSC values: val[s] = src[s] >= 0 ? csc[src[s]] : 0 ;
// valT[s] = src[s] >= 0 ? csc[tmap[src[s]]] : 0 (tmap may be null: valT not written)
void launch_fill_sell(const double* csc, const int* src, const int* tmap, double* val,
                      double* valT, int64_t n, cudaStream_t st);
// diag[r] = csc[diag_pos[perm[r]]]
void launch_extract_diag(const double* csc, const int* diag_pos, const int* perm, double* diag,
                         int n, cudaStream_t st);
// csc[diag_pos[i]] += shift
void launch_shift_diag(double* csc, const int* diag_pos, int n, double shift, cudaStream_t st);
// T1 = PT * A (values): thread per T1 entry.
//   t_row/t_col: entry coordinates (coarse i, fine j); A: CSC of the fine matrix;
//   P by fine row: prow_ptr/pcol/pval (= CSC arrays of PT)
void launch_galerkin_t1(int nnz_t1, const int* t_row, const int* t_col, const int* a_colptr,
                        const int* a_rowidx, const double* a_val, const int* prow_ptr,
                        const int* pcol, const double* pval, double* t_val, cudaStream_t st);
// Ac = T1 * P (values): thread per Ac entry (i, j).
void launch_galerkin_ac(int nnz_ac, const int* c_row, const int* c_col, const int* p_colptr,
                        const int* p_rowidx, const double* p_val, const int* t_colptr,
                        const int* t_rowidx, const double* t_val, double* c_val,
                        cudaStream_t st);
// dense (permuted) copy of a CSC matrix: D[iperm[r] + iperm[c]*n] = val
void launch_csc_to_dense(int nnz, const int* rowidx, const int* colidx, const double* val,
                         const int* iperm, double* D, int n, cudaStream_t st);
// mirror the lower triangle of a column-major n x n matrix into the upper one
void launch_symmetrize_lower(double* D, int n, cudaStream_t st);
// D (n x n, zeroed) gets ones on its diagonal
void launch_set_identity(double* D, int n, cudaStream_t st);
// u = u + Ainv * b from the packed lower-triangular 64x64 tiles of the symmetric Ainv
// (launch_pack_sym_tiles); scratch must hold dense_sym_scratch_doubles(n, k) doubles
size_t dense_sym_scratch_doubles(int n, int k);
size_t dense_sym_tiles_doubles(int n);
void launch_pack_sym_tiles(const double* A_lower, double* tiles, int n, cudaStream_t st);
void launch_dense_sym_add(const double* tiles, const double* b, double* u, double* scratch, int n,
                          int k, cudaStream_t st);

// ---- mean-curvature-flow step: device-side assembly (05_example_mean_curvature_flow/main.cpp:66-69)
// dblA (nF), mass (nV) are scratch; a_val receives M - delta * L in the CSC order of
// (rowidx, colidx); rhs = M * U (nV x k column-major).  Four launches.
void launch_mcf_assemble(int nV, int nF, const int* F, const double* U, const int* vf_ptr, const int* vf_face,
                         double* dblA, double* mass, int nnz, const int* rowidx, const int* colidx, double delta,
                         const double* Lval, double* a_val, int k, double* rhs, cudaStream_t st);

// ---- solve-time gather / scatter ----------------------------------------------
// zu[r] = z0[g[r]] ; bu[r] = RHS[g[r]] - sum_q Auk(row r, q) kv[q]   (per column)
void launch_gather_system(const double* RHS, const double* z0, const double* kv, int n_full,
                          int n_known, const int* g, const int* auk_ptr, const int* auk_q,
                          const double* auk_val, double* bu, double* zu, int nu, int k,
                          cudaStream_t st);
// z[g[r]] = zu[r]
void launch_scatter_solution(const double* zu, const int* g, double* z, int n_full, int nu, int k,
                             cudaStream_t st);
// z[kidx[i]] = kv[ksrc[i]]  (distinct kidx; ksrc = last occurrence in `known`)
void launch_scatter_known(const double* kv, const int* kidx, const int* ksrc, double* z,
                          int n_full, int n_known, int n_distinct, int k, cudaStream_t st);
// out[i] = in[perm[i]] per column (perm: new->old)  /  out[perm[i]] = in[i]
void launch_permute_in(const double* in, const int* perm, double* out, int n, int k,
                       cudaStream_t st);
void launch_permute_out(const double* in, const int* perm, double* out, int n, int k,
                        cudaStream_t st);
void launch_fill(double* p, double v, int64_t n, cudaStream_t st);


// ---- communication-avoiding patch smoother (patch.hpp, patch.cpp) ------------------------
constexpr int kPatchThreads = 512;
struct PatchDev {
  int n_patches = 0;
  const unsigned char* blob = nullptr;
  const long long* off = nullptr;  // n_patches + 1 byte offsets
  int max_blob_bytes = 0, max_vec_doubles = 0, max_active = 0;
};
// dynamic shared memory one CTA needs for k right-hand-side columns
size_t patch_smem_bytes(const PatchDev& P, int k);
// kind = PATCH_DOWN: u_out = relax(u_in), bc = PT (b - A u_out), uc_zero = 0
// kind = PATCH_UP:   u_out = relax(u_in + P uc)
// next: the patch launch that follows in the V-cycle (its blobs are prefetched into L2), or null
void launch_patch(const PatchDev& P, int kind, const double* u_in, double* u_out, const double* b, int ld,
                  const double* uc, double* bc, double* uc_zero, int ldc, int k, const PatchDev* next,
                  cudaStream_t st);
void launch_patch_fill(double* blob, const int* dst, const int* src, const double* csc, int64_t n,
                       cudaStream_t st);


// ---- device-side solve loop (min_quad_with_fixed_mg.cpp:330-347 as ONE graph launch) -------
// The residual test of every iteration runs on the device and drives a conditional WHILE node
// of the solve graph, so a solve needs no host round trip per iteration.
constexpr int kSolveCtlHis = 4096;  // residual measurements a device-side solve can record
struct SolveCtl {
  double tol;
  int max_iter, n_his, nonfinite, pad;
  double r_his[kSolveCtlHis];
};
// One thread: r = sqrt(sum of norm2[0..nchunks)), recorded in ctl->r_his; the loop goes on while
// r is finite and not < tol.  in_body: the call sits at the end of the loop body (after a
// V-cycle); once max_iter measurements exist it only ends the loop (the reference does not
// measure after its last cycle).
void launch_solve_decide(cudaGraphConditionalHandle handle, SolveCtl* ctl, const double* norm2, int nchunks,
                         int in_body, cudaStream_t st);

}  // namespace smg

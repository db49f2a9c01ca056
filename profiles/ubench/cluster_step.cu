// Microbenchmark: cost of one dependent "step" inside a thread-block cluster.
// Variants: (A) global-memory exchange (st.global + barrier.cluster release/acquire),
// (B) DSMEM exchange (st/ld.shared::cluster + barrier.cluster), (C) barrier only.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int T = 512;

__device__ __forceinline__ void cbar() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cbar_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;\n" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(T, 1) step_kernel(double* g, int n, int steps, long long* out) {
  extern __shared__ double sm[];  // n / ncta doubles per CTA
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), ncta = cl.num_blocks();
  const int chunk = n / ncta;
  const int gtid = rank * T + threadIdx.x, nth = ncta * T;
  for (int i = threadIdx.x; i < chunk; i += T) sm[i] = 1.0;
  cl.sync();
  long long t0 = clock64();
  for (int s = 0; s < steps; s++) {
    for (int row = gtid; row < n; row += nth) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const int c = (row * 7 + j * 977 + s) % n;
        if (MODE == 0) {
          acc += __ldcg(g + c);
        } else if (MODE == 1) {
          double* base = cl.map_shared_rank(sm, c / chunk);
          acc += base[c % chunk];
        }
      }
      if (MODE == 0) g[row] = acc * 0.1;
      if (MODE == 1) {
        double* base = cl.map_shared_rank(sm, row / chunk);
        base[row % chunk] = acc * 0.1;
      }
    }
    if (MODE == 3) cbar_relaxed(); else cbar();
  }
  long long t1 = clock64();
  if (gtid == 0) out[0] = t1 - t0;
}

template <int MODE>
float run(int ncta, int n, int steps, double* g, long long* out) {
  cudaFuncSetAttribute(step_kernel<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  size_t smem = (size_t)(n / ncta) * 8;
  cudaFuncSetAttribute(step_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; w++) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, step_kernel<MODE>, g, n, steps, out);
    cudaEventRecord(e1);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return -1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("sync failed %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / steps;
}

int main() {
  double* g; long long* out;
  const int n = 65536;
  cudaMalloc(&g, n * 8); cudaMemset(g, 0, n * 8); cudaMalloc(&out, 64);
  const int steps = 200;
  for (int ncta : {16, 8}) {
    for (int nn : {65536, 16384, 4096}) {
      printf("ncta %2d n %6d : global %.3f us/step  dsmem %.3f us/step  barrier(release) %.3f  barrier(relaxed) %.3f\n", ncta, nn,
             run<0>(ncta, nn, steps, g, out), run<1>(ncta, nn, steps, g, out), run<2>(ncta, nn, steps, g, out),
             run<3>(ncta, nn, steps, g, out));
    }
  }
  return 0;
}

// Dependent-chain latencies on the SM that bound one colour phase of the patch smoother
// (profiles/r2_patch_stages.md): FP64 add / mul / fma / divide, shared-memory load chains,
// __syncthreads.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_latency.cu -o fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chains(double* out, long long* cyc, double a0, double b0, int n) {
  __shared__ double sm[1024];
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    sm[i] = 1.0 + i * 1e-3;
    idx[i] = (i * 37 + 11) & 1023;
  }
  __syncthreads();
  double a = a0 + threadIdx.x, b = b0;
  long long t0, t1;
  // 1. dependent DADD
  t0 = clock64();
  for (int i = 0; i < n; i++) a = __dadd_rn(a, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 2. dependent DMUL
  t0 = clock64();
  for (int i = 0; i < n; i++) a = __dmul_rn(a, 1.0000001);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 3. dependent DFMA
  t0 = clock64();
  for (int i = 0; i < n; i++) a = fma(a, 1.0000001, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 4. dependent DDIV
  t0 = clock64();
  for (int i = 0; i < n; i++) a = __ddiv_rn(a, 1.0000001 + b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 5. dependent shared-memory load chain (index -> index)
  int j = threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < n; i++) j = idx[j];
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 6. __syncthreads
  t0 = clock64();
  for (int i = 0; i < n; i++) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 7. one smoother row: idx load -> value gather -> mul/add chain of 6 -> sub -> div -> store -> barrier
  t0 = clock64();
  for (int i = 0; i < n; i++) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 6; e++) s = __dadd_rn(s, __dmul_rn(sm[(j + e * 5) & 1023], sm[idx[(j + e) & 1023]]));
    sm[threadIdx.x] = __ddiv_rn(__dsub_rn(b, s), a + 2.0);
    __syncthreads();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = a + j + sm[(threadIdx.x + 1) & 1023];
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMallocManaged(&cyc, 8 * sizeof(long long));
  const int n = 2000;
  const char* names[7] = {"DADD chain", "DMUL chain", "DFMA chain", "DDIV chain", "LDS index chain", "__syncthreads",
                          "smoother row (6 entries + div + barrier)"};
  for (int threads : {32, 256, 512}) {
    chains<<<1, threads>>>(out, cyc, 1.5, 1e-9, n);
    cudaDeviceSynchronize();
    printf("threads %d (1 CTA):\n", threads);
    for (int i = 0; i < 7; i++) printf("  %-42s %8.1f cycles per step\n", names[i], (double)cyc[i] / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

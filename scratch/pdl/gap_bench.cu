// Launch-gap microbenchmark: how long does a chain of tiny dependent kernels take per launch on this GPU, with plain stream
// order, with programmatic dependent launch (griddepcontrol), and when kernels alternate between a 188 KB and a 0 KB
// shared-memory configuration (the generator alternates persistent tensor-core convolutions with small grid kernels).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_plain(float* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f; }
__global__ void k_pdl(float* p) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}
__global__ void k_bigsmem(float* p) { extern __shared__ float s[]; if (threadIdx.x == 0 && blockIdx.x == 0) { s[0] = p[0]; p[0] = s[0] + 1.f; } }
__global__ void k_bigsmem_pdl(float* p) {
  extern __shared__ float s[];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0 && blockIdx.x == 0) { s[0] = p[0]; p[0] = s[0] + 1.f; }
}
template <typename K>
static void launch(K k, int grid, int block, size_t smem, cudaStream_t st, bool pdl, float* p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, p);
}
int main() {
  float* p; cudaMalloc(&p, 4); cudaMemset(p, 0, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaFuncSetAttribute(k_bigsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 188 * 1024);
  cudaFuncSetAttribute(k_bigsmem_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, 188 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int N = 2000;
  for (int variant = 0; variant < 6; ++variant) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a, st);
      for (int i = 0; i < N; ++i) {
        switch (variant) {
          case 0: launch(k_plain, 148, 256, 0, st, false, p); break;
          case 1: launch(k_pdl, 148, 256, 0, st, true, p); break;
          case 2: if (i & 1) launch(k_bigsmem, 148, 480, 188 * 1024, st, false, p); else launch(k_plain, 592, 256, 0, st, false, p); break;
          case 3: if (i & 1) launch(k_bigsmem_pdl, 148, 480, 188 * 1024, st, true, p); else launch(k_pdl, 592, 256, 0, st, true, p); break;
          case 4: launch(k_plain, 1, 32, 0, st, false, p); break;
          case 5: launch(k_pdl, 1, 32, 0, st, true, p); break;
        }
      }
      cudaEventRecord(b, st);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 1) {
        const char* names[] = {"148x256 plain", "148x256 PDL", "alternating 188KB/0KB plain", "alternating 188KB/0KB PDL", "1x32 plain", "1x32 PDL"};
        printf("%-32s %.2f us per launch\n", names[variant], ms * 1e3 / N);
      }
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

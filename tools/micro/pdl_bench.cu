// Launch-floor microbenchmark: a chain of N dependent tiny kernels, plain stream vs CUDA graph, with and
// without programmatic dependent launch.  Prints us per launch.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

__global__ void tiny(float* p, int use_pdl) {
  if (use_pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  p[i] = p[i] * 1.0001f + 1.f;
}

static void launch(float* p, int blocks, int threads, size_t smem, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, tiny, p, pdl ? 1 : 0);
}

int main() {
  const int N = 300;
  float* p;
  cudaMalloc(&p, 148 * 8 * 256 * sizeof(float));
  cudaMemset(p, 0, 148 * 8 * 256 * sizeof(float));
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaFuncSetAttribute(tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int blocks : {1, 148, 1184}) {
    for (size_t smem : {size_t(0), size_t(100 * 1024), size_t(200 * 1024)}) {
      for (int pdl = 0; pdl < 2; ++pdl) {
        for (int graph = 0; graph < 2; ++graph) {
          cudaGraphExec_t ge = nullptr;
          if (graph) {
            cudaGraph_t g;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            for (int i = 0; i < N; ++i) launch(p, blocks, 256, smem, st, pdl);
            cudaStreamEndCapture(st, &g);
            cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
            if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return 1; }
            cudaGraphDestroy(g);
          }
          float best = 1e9f;
          for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0, st);
            if (graph) cudaGraphLaunch(ge, st);
            else for (int i = 0; i < N; ++i) launch(p, blocks, 256, smem, st, pdl);
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
          }
          printf("blocks=%5d smem=%6zu pdl=%d graph=%d : %.2f us/launch\n", blocks, smem, pdl, graph, best * 1e3f / N);
          if (ge) cudaGraphExecDestroy(ge);
        }
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("final: %s\n", cudaGetErrorString(e));
  return 0;
}

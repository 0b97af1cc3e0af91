// Micro-probe: what does a dependent kernel launch cost inside a CUDA graph on this GPU, with and without programmatic
// dependent launch (PDL)?  A chain of N small kernels (each ~W us of work on all SMs) captured into a graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_probe pdl_probe.cu && ./pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void work_kernel(float* buf, int iters) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  float v = buf[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < iters; ++i) v = fmaf(v, 1.0001f, 0.5f);
  buf[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <bool PDL>
static float run(int n, int iters, int blocks, cudaStream_t st, float* buf) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = PDL ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, work_kernel<PDL>, buf, iters);
    if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); break; }
  }
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture error %s\n", cudaGetErrorString(e)); return -1; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate error %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
  cudaEventRecord(e0, st);
  for (int w = 0; w < 10; ++w) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms / 10.f * 1000.f / n;   // us per kernel
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  float* buf;
  cudaMalloc(&buf, 148 * 8 * 256 * 4);
  cudaMemset(buf, 0, 148 * 8 * 256 * 4);
  const int n = 500;
  for (int blocks : {148, 148 * 8}) {
    for (int iters : {1, 2000, 20000}) {
      float a = run<false>(n, iters, blocks, st, buf);
      float b = run<true>(n, iters, blocks, st, buf);
      printf("blocks %5d iters %6d : plain %7.2f us/kernel   PDL %7.2f us/kernel   saved %6.2f us\n", blocks, iters, a, b, a - b);
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

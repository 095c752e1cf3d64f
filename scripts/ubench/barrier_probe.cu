// Which block-barrier forms of a warp-specialised kernel does compute-sanitizer synccheck accept?
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o barrier_probe barrier_probe.cu
//   compute-sanitizer --tool synccheck ./barrier_probe <variant 0..3>
// Every variant: warps 0..7 ("rankers") and warp 8 ("look-back") meet at two block-wide barriers reached from different branches.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __noinline__ void barrier_fn(int n) { asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); }

template <int V>
__global__ void __launch_bounds__(288) probe(int* out)
{
  __shared__ int s[2];
  const int tid = threadIdx.x;
  const bool ranker = tid < 256;
  if (tid == 0) s[0] = s[1] = 0;
  __syncthreads();
  if (!ranker) {
    // look-back role: waits for the rankers' data (S2), publishes its own (S4), leaves
    if (V == 0) __syncthreads();
    if (V == 1) asm volatile("bar.sync 2, %0;" ::"r"(288) : "memory");
    if (V == 2) asm volatile("bar.sync 2, %0;" ::"r"(288) : "memory");
    if (V == 3) barrier_fn(288);
    if (tid == 256) s[1] = s[0] + 1;
    if (V == 0) __syncthreads();
    if (V == 1) asm volatile("bar.sync 2, %0;" ::"r"(288) : "memory");
    if (V == 2) { __threadfence_block(); asm volatile("bar.arrive 3, %0;" ::"r"(288) : "memory"); }
    if (V == 3) barrier_fn(288);
    return;
  }
  if (tid == 5) s[0] = 41;
  if (V == 0) __syncthreads();
  if (V == 1) asm volatile("bar.sync 2, %0;" ::"r"(288) : "memory");
  if (V == 2) { __threadfence_block(); asm volatile("bar.arrive 2, %0;" ::"r"(288) : "memory"); asm volatile("bar.sync 1, %0;" ::"r"(256) : "memory"); }
  if (V == 3) barrier_fn(288);
  if (V == 0) __syncthreads();
  if (V == 1) asm volatile("bar.sync 2, %0;" ::"r"(288) : "memory");
  if (V == 2) asm volatile("bar.sync 3, %0;" ::"r"(288) : "memory");
  if (V == 3) barrier_fn(288);
  if (tid == 0) out[blockIdx.x] = s[1];
}

int main(int argc, char** argv)
{
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  int* d;
  cudaMalloc(&d, 64 * sizeof(int));
  if (v == 0) probe<0><<<64, 288>>>(d);
  if (v == 1) probe<1><<<64, 288>>>(d);
  if (v == 2) probe<2><<<64, 288>>>(d);
  if (v == 3) probe<3><<<64, 288>>>(d);
  int h[64];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int ok = 1;
  for (int i = 0; i < 64; ++i) ok &= h[i] == 42;
  printf("variant %d: %s, values %s\n", v, cudaGetErrorString(e), ok ? "ok" : "WRONG");
  return 0;
}

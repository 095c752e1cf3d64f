// Micro-benchmark: random 8-byte gather over a large array under different L2 fetch-granularity limits
// and load flavours. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__host__ __device__ inline uint64_t mix(uint64_t x){ x += 0x9E3779B97F4A7C15ull; x=(x^(x>>30))*0xBF58476D1CE4E5B9ull; x=(x^(x>>27))*0x94D049BB133111EBull; return x^(x>>31);} 
__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t m){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) idx[i]=(uint32_t)(mix(i)%m);} 
template<int MODE,int U> __global__ void gather(const uint64_t* __restrict__ src,const uint32_t* __restrict__ idx,uint64_t* __restrict__ out,size_t n){
  size_t stride=(size_t)gridDim.x*blockDim.x*U;
  for(size_t b=(blockIdx.x*(size_t)blockDim.x+threadIdx.x)*U;b<n;b+=stride){
    uint32_t m[U]; uint64_t v[U];
    #pragma unroll
    for(int j=0;j<U;++j) m[j]= b+j<n? idx[b+j]:0;
    #pragma unroll
    for(int j=0;j<U;++j){
      const uint64_t* p=src+m[j];
      if(MODE==0) v[j]=*p;
      else if(MODE==1) asm volatile("ld.global.nc.L1::no_allocate.u64 %0,[%1];":"=l"(v[j]):"l"(p));
      else if(MODE==2) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u64 %0,[%1];":"=l"(v[j]):"l"(p));
      else if(MODE==3) asm volatile("ld.global.cv.u64 %0,[%1];":"=l"(v[j]):"l"(p));
      else if(MODE==4) asm volatile("ld.global.nc.L1::evict_first.u64 %0,[%1];":"=l"(v[j]):"l"(p));
    }
    #pragma unroll
    for(int j=0;j<U;++j) if(b+j<n) out[b+j]=v[j];
  }
}
template<int MODE,int U> float run(const uint64_t* s,const uint32_t* i,uint64_t* o,size_t n,int grid){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE,U><<<grid,256>>>(s,i,o,n); cudaDeviceSynchronize();
  cudaEventRecord(a); gather<MODE,U><<<grid,256>>>(s,i,o,n); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){
  size_t n=(size_t)1<<29; uint32_t m=(uint32_t)(1u<<30);
  uint64_t *src,*out; uint32_t* idx; cudaMalloc(&src,(size_t)m*8); cudaMalloc(&out,n*8); cudaMalloc(&idx,n*4);
  cudaMemset(src,1,(size_t)m*8); fill_idx<<<1184,256>>>(idx,n,m); cudaDeviceSynchronize();
  size_t lims[4]={128,64,32,0};
  for(int li=0;li<3;++li){
    cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,lims[li]); size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity);
    printf("limit set %zu -> %s, get %zu\n",lims[li],cudaGetErrorName(e),got);
    printf("  plain U4 %.2f ms | nc.noalloc U4 %.2f | nc.L2::64B U4 %.2f | cv U4 %.2f | evict U4 %.2f | plain U8 %.2f | nc U8 %.2f | plain U16 %.2f\n",
      run<0,4>(src,idx,out,n,2368),run<1,4>(src,idx,out,n,2368),run<2,4>(src,idx,out,n,2368),run<3,4>(src,idx,out,n,2368),run<4,4>(src,idx,out,n,2368),run<0,8>(src,idx,out,n,2368),run<1,8>(src,idx,out,n,2368),run<0,16>(src,idx,out,n,1184));
  }
  // locality probe: indices confined to a 64 MB window (fits L2)
  fill_idx<<<1184,256>>>(idx,n,(uint32_t)(1u<<23)); cudaDeviceSynchronize();
  printf("L2-resident source (64 MB): plain U4 %.2f ms\n",run<0,4>(src,idx,out,n,2368));
  printf("rows %zu (x8B), algorithmic 20 B/row = %.1f GB\n",n,n*20/1e9);
  return 0; }

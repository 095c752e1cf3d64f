// Micro-benchmark of warp-level digit ranking strategies on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rank_probe rank_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lanemask_lt(){unsigned m; asm("mov.u32 %0, %%lanemask_lt;":"=r"(m)); return m;}
__device__ __forceinline__ uint32_t hash32(uint32_t x){x^=x>>16;x*=0x7feb352dU;x^=x>>15;x*=0x846ca68bU;x^=x>>16;return x;}
constexpr int IPT=16, NW=12, THREADS=NW*32;
// MODE 0: MATCH.ANY ; 1: 8 ballots ; 2: smem atomicOr bitmap (cub style) ; 3: unordered atomicAdd-with-return (order probe)
template<int MODE>
__global__ void __launch_bounds__(THREADS,2) k(uint32_t* out, int iters, uint32_t dmask, unsigned long long* disorder){
  __shared__ uint32_t hist[NW][256];
  __shared__ uint32_t bm[NW][256];
  const int warp=threadIdx.x>>5, lane=threadIdx.x&31;
  uint32_t* h=hist[warp]; uint32_t* m=bm[warp];
  for(int j=lane;j<256;j+=32){h[j]=0;m[j]=0;}
  __syncwarp();
  uint32_t acc=0; unsigned long long bad=0;
  for(int it=0;it<iters;++it){
    uint32_t d[IPT];
    #pragma unroll
    for(int i=0;i<IPT;++i) d[i]=hash32((blockIdx.x*THREADS+threadIdx.x)*977u+it*131u+i)&dmask;
    #pragma unroll
    for(int i=0;i<IPT;++i){
      unsigned peers;
      if(MODE==0){ peers=__match_any_sync(0xffffffffu,d[i]); }
      else if(MODE==1){ peers=0xffffffffu;
        #pragma unroll
        for(int b=0;b<8;++b){unsigned bit=(d[i]>>b)&1u; unsigned v=__ballot_sync(0xffffffffu,bit); peers&=v^(bit-1u);} }
      else if(MODE==2){ atomicOr(&m[d[i]],1u<<lane); __syncwarp(); peers=m[d[i]]; }
      if(MODE<=2){
        unsigned lt=__popc(peers&lanemask_lt()); uint32_t prev=0;
        if(lt==0){prev=h[d[i]]; h[d[i]]=prev+__popc(peers); if(MODE==2) m[d[i]]=0;}
        __syncwarp();
        prev=__shfl_sync(0xffffffffu,prev,__ffs(peers)-1);
        acc+=prev+lt;
      } else {
        uint32_t r=atomicAdd(&h[d[i]],1u);
        // order probe: among lanes with the same digit, ranks must ascend with lane
        unsigned peers2=__match_any_sync(0xffffffffu,d[i]);
        unsigned lt=__popc(peers2&lanemask_lt());
        uint32_t base=__shfl_sync(0xffffffffu,r,__ffs(peers2)-1);
        if(r!=base+lt) ++bad;
        acc+=r;
      }
    }
  }
  out[blockIdx.x*THREADS+threadIdx.x]=acc;
  if(MODE==3 && bad) atomicAdd(disorder,bad);
}
template<int MODE> void run(const char* name,uint32_t dmask,uint32_t* out,unsigned long long* dis){
  int iters=200, grid=148*2; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(dis,0,8);
  k<MODE><<<grid,THREADS>>>(out,10,dmask,dis); cudaDeviceSynchronize();
  cudaMemset(dis,0,8);
  cudaEventRecord(a); k<MODE><<<grid,THREADS>>>(out,iters,dmask,dis); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); unsigned long long h=0; cudaMemcpy(&h,dis,8,cudaMemcpyDeviceToHost);
  double items=(double)grid*THREADS*IPT*iters;
  printf("%-28s dmask=%3u  %.3f ms  %.1f Gitems/s  (%.2f cycles/warp-item/SM @1.9GHz)  disorder=%llu\n",name,dmask,ms,items/ms/1e6, ms*1e-3*1.9e9/(items/32/148),h);
}
int main(){ uint32_t* out; unsigned long long* dis; cudaMalloc(&out,148*2*THREADS*4); cudaMalloc(&dis,8);
  uint32_t masks[4]={255,15,3,0};
  for(int mi=0;mi<4;++mi){
    run<0>("MATCH.ANY",masks[mi],out,dis); run<1>("8 ballots",masks[mi],out,dis); run<2>("smem atomicOr bitmap",masks[mi],out,dis); run<3>("atomicAdd ret + order probe",masks[mi],out,dis);
  }
  return 0; }

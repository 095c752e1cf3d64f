// SIMT emulator for CPU-side functional tests of the CUDA kernels (TEST INFRASTRUCTURE ONLY).
//
// tests/emu/build_emu.py compiles cudf_b200/csrc/*.cu as plain C++ (g++ -x c++ -DB2_EMU) against THIS header, which
// stands in for <cuda_runtime.h>: "device memory" is host memory, every kernel launch runs its CTAs one after the
// other, and the threads of a CTA are fibers that meet at __syncthreads / named barriers / warp collectives
// (emu_runtime.cpp).  It checks index arithmetic, barrier structure and protocol logic of kernels that have not yet
// run on hardware; it says nothing about memory ordering, races or speed.  The product never loads this library:
// only tests/test_emu_*.py do, in a subprocess, by rebinding the ctypes entry points (tests/emu/harness.py).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <cmath>
using std::sqrt;

// ---- qualifiers ---------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

// ---- vector types -------------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// ---- runtime (emu_runtime.cpp) ---------------------------------------------------------------------------
namespace emu {
struct thread_ctx {
  uint3 tid;
  unsigned linear, lane, warp;
};
extern thread_ctx* cur;       // the running fiber
extern uint3 g_block_idx;
extern dim3 g_block_dim, g_grid_dim;
void launch(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
unsigned char* dynamic_smem();
void sync_threads();
void named_barrier(int id, int nthreads);
// all lanes of `mask` deposit `v`; returns after everyone arrived with the 32 deposited values in out[]
// (lanes outside the mask or already exited read as 0); returns the mask of lanes that took part
unsigned warp_exchange(unsigned mask, uint64_t v, uint64_t out[32]);
void yield();
}  // namespace emu

#define threadIdx (::emu::cur->tid)
#define blockIdx (::emu::g_block_idx)
#define blockDim (::emu::g_block_dim)
#define gridDim (::emu::g_grid_dim)

// ---- host API ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef void* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef void* cudaMemPool_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaMemPoolAttrReleaseThreshold = 4 };
enum { cudaLimitMaxL2FetchGranularity = 5 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t);
cudaError_t cudaFreeAsync(void* p, cudaStream_t);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char* cudaGetErrorName(cudaError_t);
const char* cudaGetErrorString(cudaError_t);
cudaError_t cudaGetDevice(int* dev);
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int dev);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int attr, void* value);
cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t);
cudaError_t cudaDeviceSetLimit(int, size_t);
cudaError_t cudaEventCreate(cudaEvent_t*);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t*, unsigned);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
constexpr unsigned cudaEventDisableTiming = 2;
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*);
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void*);
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- device intrinsics -----------------------------------------------------------------------------------
inline void __syncthreads() { ::emu::sync_threads(); }
inline void __threadfence() {}
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }  // only named in discarded branches
inline void __nanosleep(unsigned) { ::emu::yield(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (unsigned)(x >> (sh & 31));
}
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }

// min / max over mixed arithmetic types (CUDA's global overloads)
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)b < (C)a ? (C)b : (C)a; }
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)b : (C)a; }

// atomics: one OS thread runs all fibers and a fiber is never preempted, so plain read-modify-write is atomic
template <typename T, typename V> inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename V> inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename V> inline T atomicMax(T* p, V v) { T o = *p; if (o < (T)v) *p = (T)v; return o; }
template <typename T, typename V> inline T atomicOr(T* p, V v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename V> inline T atomicExch(T* p, V v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename A, typename B> inline T atomicCAS(T* p, A cmp, B val) { T o = *p; if (o == (T)cmp) *p = (T)val; return o; }

// warp collectives
inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t o[32]; ::emu::warp_exchange(mask, 0, o); }
template <typename T> inline uint64_t emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T emu_unbits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src)
{
  uint64_t o[32];
  ::emu::warp_exchange(mask, emu_bits(v), o);
  return emu_unbits<T>(o[src & 31]);
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int x)
{
  uint64_t o[32];
  ::emu::warp_exchange(mask, emu_bits(v), o);
  return emu_unbits<T>(o[(::emu::cur->lane ^ (unsigned)x) & 31]);
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d)
{
  uint64_t o[32];
  ::emu::warp_exchange(mask, emu_bits(v), o);
  const unsigned l = ::emu::cur->lane;
  return l >= d ? emu_unbits<T>(o[l - d]) : v;
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, pred ? 1 : 0, o);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (((part >> i) & 1u) && o[i]) r |= 1u << i;
  return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <typename T> inline unsigned __match_any_sync(unsigned mask, T v)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, emu_bits(v), o);
  const uint64_t mine = emu_bits(v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (((part >> i) & 1u) && o[i] == mine) r |= 1u << i;
  return r;
}
template <typename T> inline unsigned __match_all_sync(unsigned mask, T v, int* pred)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, emu_bits(v), o);
  const uint64_t mine = emu_bits(v);
  bool all = true;
  for (int i = 0; i < 32; ++i) if (((part >> i) & 1u) && o[i] != mine) all = false;
  *pred = all ? 1 : 0;
  return all ? part : 0u;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if ((part >> i) & 1u) r |= (unsigned)o[i];
  return r;
}
inline unsigned __reduce_and_sync(unsigned mask, unsigned v)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, v, o);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; ++i) if ((part >> i) & 1u) r &= (unsigned)o[i];
  return r;
}
template <typename T> inline T __reduce_add_sync(unsigned mask, T v)
{
  uint64_t o[32];
  const unsigned part = ::emu::warp_exchange(mask, emu_bits(v), o);
  T r = 0;
  for (int i = 0; i < 32; ++i) if ((part >> i) & 1u) r += emu_unbits<T>(o[i]);
  return r;
}

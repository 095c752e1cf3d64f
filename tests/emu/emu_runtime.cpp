// Runtime of the SIMT emulator (see include/cuda_runtime.h): fibers, barriers, warp rendezvous, guarded "device" memory.
// TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>

#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <map>
#include <vector>

namespace emu {

thread_ctx* cur = nullptr;
uint3 g_block_idx{0, 0, 0};
dim3 g_block_dim, g_grid_dim;

extern const char* g_kernel_name;

namespace {

// ---- minimal x86-64 context switch (callee-saved registers + stack pointer) -----------------------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

constexpr size_t STACK_BYTES = 128 * 1024;

struct fiber {
  thread_ctx ctx;
  void* sp        = nullptr;
  char* stack     = nullptr;
  int state       = 0;  // 0 runnable, 1 waiting, 2 done
  const volatile unsigned* wait_gen = nullptr;
  unsigned wait_val = 0;
};

struct barrier_t {
  unsigned count = 0, gen = 0, expected = 0;  // expected == 0: all live threads
};

struct warp_t {
  uint64_t vals[32];
  uint64_t snap[2][32];
  unsigned snap_part[2];
  unsigned arrived = 0, gen = 0, want = 0, live = 0;
};

std::vector<fiber> g_fibers;
std::vector<char*> g_stack_pool;
std::vector<warp_t> g_warps;
barrier_t g_bar[16];
unsigned g_live = 0;
void* g_sched_sp = nullptr;
fiber* g_cur_fiber = nullptr;
const std::function<void()>* g_body = nullptr;
alignas(128) unsigned char g_smem[256 * 1024];

void fiber_exit();

void fiber_entry()
{
  (*g_body)();
  fiber_exit();
}

void to_scheduler() { emu_switch(&g_cur_fiber->sp, g_sched_sp); }

void wait_on(const volatile unsigned* gen, unsigned val)
{
  g_cur_fiber->state    = 1;
  g_cur_fiber->wait_gen = gen;
  g_cur_fiber->wait_val = val;
  to_scheduler();
}

void release_block_barrier_if_complete(barrier_t& b)
{
  const unsigned need = b.expected ? b.expected : g_live;
  if (b.count > 0 && b.count >= need) {
    b.count = 0;
    b.expected = 0;
    ++b.gen;
  }
}

void complete_warp_if_ready(warp_t& w)
{
  const unsigned need = w.want & w.live;
  if (w.arrived != 0 && (w.arrived & need) == need) {
    const unsigned slot = w.gen & 1u;
    for (int i = 0; i < 32; ++i) w.snap[slot][i] = ((w.arrived >> i) & 1u) ? w.vals[i] : 0;
    w.snap_part[slot] = w.arrived;
    w.arrived = 0;
    w.want = 0;
    ++w.gen;
  }
}

void fiber_exit()
{
  fiber* f = g_cur_fiber;
  f->state = 2;
  --g_live;
  warp_t& w = g_warps[f->ctx.warp];
  w.live &= ~(1u << f->ctx.lane);
  complete_warp_if_ready(w);  // lanes waiting for an exited lane are released (sm_70+ semantics)
  for (auto& b : g_bar) release_block_barrier_if_complete(b);
  to_scheduler();
  std::abort();  // never resumed
}

char* get_stack()
{
  if (!g_stack_pool.empty()) {
    char* s = g_stack_pool.back();
    g_stack_pool.pop_back();
    return s;
  }
  return static_cast<char*>(std::malloc(STACK_BYTES));
}

void run_block(unsigned nthreads)
{
  g_fibers.assign(nthreads, fiber{});
  g_warps.assign((nthreads + 31) / 32, warp_t{});
  for (auto& b : g_bar) b = barrier_t{};
  g_live = nthreads;
  for (unsigned t = 0; t < nthreads; ++t) {
    fiber& f = g_fibers[t];
    f.ctx.linear = t;
    f.ctx.tid    = uint3{t % g_block_dim.x, (t / g_block_dim.x) % g_block_dim.y, t / (g_block_dim.x * g_block_dim.y)};
    f.ctx.lane   = t & 31u;
    f.ctx.warp   = t >> 5;
    g_warps[f.ctx.warp].live |= 1u << f.ctx.lane;
    f.stack = get_stack();
    // initial frame: six callee-saved registers, then the return address = fiber_entry; keep (rsp + 8) % 16 == 0 at entry
    uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + STACK_BYTES) & ~uintptr_t(15);
    void** sp = reinterpret_cast<void**>(top);
    *--sp = nullptr;                                   // fake return address of fiber_entry (alignment slot)
    *--sp = reinterpret_cast<void*>(&fiber_entry);     // `ret` target
    for (int i = 0; i < 6; ++i) *--sp = nullptr;       // rbp rbx r12 r13 r14 r15
    f.sp = sp;
  }
  unsigned done = 0;
  while (done < nthreads) {
    bool progressed = false;
    for (unsigned t = 0; t < nthreads; ++t) {
      fiber& f = g_fibers[t];
      if (f.state == 2) continue;
      if (f.state == 1) {
        if (*f.wait_gen == f.wait_val) continue;  // still blocked
        f.state = 0;
      }
      g_cur_fiber = &f;
      cur = &f.ctx;
      emu_switch(&g_sched_sp, f.sp);
      progressed = true;
      if (f.state == 2) ++done;
    }
    if (!progressed) {
      std::fprintf(stderr, "emu: DEADLOCK in %s block (%u,%u,%u): %u of %u threads finished\n", g_kernel_name, g_block_idx.x, g_block_idx.y,
                   g_block_idx.z, done, nthreads);
      for (unsigned t = 0; t < nthreads && t < 2048; ++t)
        if (g_fibers[t].state == 1 && (t % 32 == 0 || t < 4))
          std::fprintf(stderr, "  thread %u waits (bar0 count %u gen %u, warp arrived %08x want %08x live %08x)\n", t, g_bar[0].count,
                       g_bar[0].gen, g_warps[t >> 5].arrived, g_warps[t >> 5].want, g_warps[t >> 5].live);
      std::abort();
    }
  }
  for (auto& f : g_fibers) g_stack_pool.push_back(f.stack);
  cur = nullptr;
  g_cur_fiber = nullptr;
}

// ---- guarded allocations ------------------------------------------------------------------------------------
constexpr size_t GUARD = 256;
constexpr unsigned char GUARD_BYTE = 0xA5;
std::map<void*, size_t> g_allocs;

void check_guards(const char* when)
{
  for (auto& kv : g_allocs) {
    const unsigned char* p = static_cast<const unsigned char*>(kv.first);
    for (size_t i = 0; i < GUARD; ++i) {
      if (p[-(ptrdiff_t)GUARD + (ptrdiff_t)i] != GUARD_BYTE || p[kv.second + i] != GUARD_BYTE) {
        std::fprintf(stderr, "emu: out-of-bounds write next to a %zu-byte allocation (%s, %s guard, byte %zu)\n", kv.second, when,
                     p[-(ptrdiff_t)GUARD + (ptrdiff_t)i] != GUARD_BYTE ? "front" : "back", i);
        std::abort();
      }
    }
  }
}

}  // namespace

unsigned char* dynamic_smem() { return g_smem; }

void yield()
{
  g_cur_fiber->state = 0;
  to_scheduler();
}

void sync_threads()
{
  barrier_t& b = g_bar[0];
  const unsigned gen = b.gen;
  ++b.count;
  release_block_barrier_if_complete(b);
  if (b.gen == gen) wait_on(&b.gen, gen);
}

void named_barrier(int id, int nthreads)
{
  barrier_t& b = g_bar[id & 15];
  const unsigned gen = b.gen;
  b.expected = (unsigned)nthreads;
  ++b.count;
  release_block_barrier_if_complete(b);
  if (b.gen == gen) wait_on(&b.gen, gen);
}

unsigned warp_exchange(unsigned mask, uint64_t v, uint64_t out[32])
{
  warp_t& w = g_warps[cur->warp];
  const unsigned lane = cur->lane;
  const unsigned gen = w.gen;
  w.vals[lane] = v;
  w.arrived |= 1u << lane;
  w.want |= mask;
  complete_warp_if_ready(w);
  if (w.gen == gen) wait_on(&w.gen, gen);
  const unsigned slot = gen & 1u;
  std::memcpy(out, w.snap[slot], sizeof(uint64_t) * 32);
  return w.snap_part[slot];
}

const char* g_kernel_name = "";

void launch(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
  g_kernel_name = name;
  static const bool trace = std::getenv("B2_EMU_TRACE") != nullptr;
  if (trace) std::fprintf(stderr, "emu: launch %s grid %u block %u smem %zu\n", name, grid.x, block.x, smem);
  if (smem > sizeof(g_smem)) {
    std::fprintf(stderr, "emu: %zu bytes of dynamic shared memory requested\n", smem);
    std::abort();
  }
  g_grid_dim  = grid;
  g_block_dim = block;
  g_body      = &body;
  const unsigned nthreads = block.x * block.y * block.z;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        g_block_idx = uint3{x, y, z};
        std::memset(g_smem, 0xCD, smem);  // shared memory starts undefined
        run_block(nthreads);
      }
  g_body = nullptr;
  check_guards(name);
}

}  // namespace emu

// ---- host API --------------------------------------------------------------------------------------------------
using emu::g_allocs;

// B2_EMU_GUARDPAGE=1: every allocation ends (up to 16-byte alignment) at an inaccessible page, so that reads or writes
// past the end fault immediately instead of going unnoticed.
static bool guard_pages()
{
  static const bool on = std::getenv("B2_EMU_GUARDPAGE") != nullptr;
  return on;
}
static std::map<void*, std::pair<void*, size_t>> g_maps;  // user pointer -> (mapping base, mapping bytes)

static cudaError_t guarded_alloc(void** p, size_t bytes)
{
  const size_t page = 4096;
  const size_t rounded = (bytes + 15) / 16 * 16;
  const size_t body = (rounded + page - 1) / page * page;
  unsigned char* base = static_cast<unsigned char*>(mmap(nullptr, body + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
  if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
  mprotect(base + body, page, PROT_NONE);
  unsigned char* user = base + body - rounded;
  std::memset(base, 0xCD, body);
  *p = user;
  g_maps[user] = {base, body + page};
  return cudaSuccess;
}

extern "C" void* emu_alloc(size_t bytes)
{
  void* p = nullptr;
  return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
}
extern "C" void emu_free(void* p) { cudaFree(p); }

cudaError_t cudaMalloc(void** p, size_t bytes)
{
  if (guard_pages()) return guarded_alloc(p, bytes ? bytes : 1);
  unsigned char* raw = static_cast<unsigned char*>(std::malloc(bytes + 2 * emu::GUARD + 64));
  if (!raw) return cudaErrorMemoryAllocation;
  std::memset(raw, emu::GUARD_BYTE, emu::GUARD);
  std::memset(raw + emu::GUARD, 0xCD, bytes);  // fresh device memory is undefined
  std::memset(raw + emu::GUARD + bytes, emu::GUARD_BYTE, emu::GUARD + 64);
  *p = raw + emu::GUARD;
  g_allocs[*p] = bytes;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
  if (!p) return cudaSuccess;
  if (guard_pages()) {
    auto it = g_maps.find(p);
    if (it == g_maps.end()) {
      std::fprintf(stderr, "emu: cudaFree of an unknown pointer\n");
      std::abort();
    }
    munmap(it->second.first, it->second.second);
    g_maps.erase(it);
    return cudaSuccess;
  }
  auto it = g_allocs.find(p);
  if (it == g_allocs.end()) {
    std::fprintf(stderr, "emu: cudaFree of an unknown pointer\n");
    std::abort();
  }
  emu::check_guards("at free");
  std::memset(p, 0xDD, it->second);  // poison
  g_allocs.erase(it);
  std::free(static_cast<unsigned char*>(p) - emu::GUARD);
  return cudaSuccess;
}
cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { std::memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) { std::memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaErrorEmu"; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulator error"; }
cudaError_t cudaGetDevice(int* dev) { *dev = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
struct emu_event { std::chrono::steady_clock::time_point t; };
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }

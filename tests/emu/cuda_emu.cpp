// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY: fiber-based SIMT emulator (see cuda_emu.h).
#include "cuda_emu.h"
#include <ucontext.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/mman.h>

namespace ojb_emu {

static const size_t kStack = 192 * 1024;

struct WarpState {
  unsigned long long slot[2][32];
  unsigned count = 0, gen = 0, mask = 0;
};

struct Fiber {
  ucontext_t uc;
  ThreadCtx tc;
  bool done = false;
  unsigned lane = 0, warp = 0;
  void* stack = nullptr;
};

struct BlockRun {
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  unsigned bar_count = 0, bar_gen = 0;
  unsigned nbar_count[16] = {0}, nbar_gen[16] = {0};      // named barriers (bar.sync / bar.arrive id, count)
  unsigned nthreads = 0;
  ucontext_t sched;
  Fiber* cur = nullptr;
  const std::function<void()>* body = nullptr;
  void* dyn = nullptr;
  unsigned long progress = 0;
};

static thread_local BlockRun* t_run = nullptr;
static thread_local ThreadCtx t_idle;

ThreadCtx& ctx() { return t_run && t_run->cur ? t_run->cur->tc : t_idle; }
void* dyn_smem() { return t_run->dyn; }

static void yield_fiber() {
  BlockRun* r = t_run;
  Fiber* f = r->cur;
  swapcontext(&f->uc, &r->sched);
}

unsigned long long collective(unsigned mask, unsigned long long v, unsigned long long* all32) {
  BlockRun* r = t_run;
  Fiber* f = r->cur;
  WarpState& w = r->warps[f->warp];
  unsigned my_gen = w.gen;
  if (w.count == 0) w.mask = mask;
  else if (w.mask != mask) {
    fprintf(stderr, "[cuda_emu] warp collective with inconsistent masks %08x vs %08x\n", w.mask, mask);
    abort();
  }
  if (!((mask >> f->lane) & 1u)) {
    fprintf(stderr, "[cuda_emu] lane %u calls a collective whose mask %08x excludes it\n", f->lane, mask);
    abort();
  }
  w.slot[my_gen & 1][f->lane] = v;
  unsigned need = (unsigned)__builtin_popcount(mask);
  // lanes of a partial last warp that do not exist never arrive
  unsigned first = f->warp * 32, exist = std::min(32u, r->nthreads - first);
  if (exist < 32) need = (unsigned)__builtin_popcount(mask & ((1u << exist) - 1u));
  r->progress++;
  if (++w.count == need) { w.count = 0; w.gen++; }
  else while (w.gen == my_gen) yield_fiber();
  for (int i = 0; i < 32; ++i) all32[i] = w.slot[my_gen & 1][i];
  return v;
}

void block_barrier() {
  BlockRun* r = t_run;
  unsigned my_gen = r->bar_gen;
  r->progress++;
  if (++r->bar_count == r->nthreads) { r->bar_count = 0; r->bar_gen++; }
  else while (r->bar_gen == my_gen) yield_fiber();
}

// bar.sync id, count (wait = true) / bar.arrive id, count (wait = false): the barrier completes when `count`
// threads have arrived or waited
void named_barrier(unsigned id, unsigned count, bool wait) {
  BlockRun* r = t_run;
  unsigned my_gen = r->nbar_gen[id & 15];
  r->progress++;
  if (++r->nbar_count[id & 15] == count) { r->nbar_count[id & 15] = 0; r->nbar_gen[id & 15]++; }
  else if (wait) while (r->nbar_gen[id & 15] == my_gen) yield_fiber();
}

static void fiber_entry() {
  BlockRun* r = t_run;
  Fiber* f = r->cur;
  (*r->body)();
  f->done = true;
  r->progress++;
  swapcontext(&f->uc, &r->sched);
}

struct StackPool {
  std::vector<void*> free_list;
  ~StackPool() { for (void* p : free_list) munmap(p, kStack); }
  void* get() {
    if (!free_list.empty()) { void* p = free_list.back(); free_list.pop_back(); return p; }
    void* p = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) { perror("mmap"); abort(); }
    return p;
  }
  void put(void* p) { free_list.push_back(p); }
};
static thread_local StackPool t_stacks;

static void run_block(dim3 grid, dim3 block, uint3 bid, size_t smem, const std::function<void()>& body) {
  BlockRun run;
  unsigned n = block.x * block.y * block.z;
  run.nthreads = n;
  run.fibers.resize(n);
  run.warps.resize((n + 31) / 32);
  run.body = &body;
  std::vector<unsigned long long> dyn((smem + 7) / 8 + 2, 0);
  run.dyn = dyn.data();
  t_run = &run;
  for (unsigned i = 0; i < n; ++i) {
    Fiber& f = run.fibers[i];
    f.tc.tid = uint3{ i % block.x, (i / block.x) % block.y, i / (block.x * block.y) };
    f.tc.bid = bid; f.tc.bdim = block; f.tc.gdim = grid;
    f.lane = i & 31; f.warp = i >> 5;
    f.stack = t_stacks.get();
    getcontext(&f.uc);
    f.uc.uc_stack.ss_sp = f.stack; f.uc.uc_stack.ss_size = kStack; f.uc.uc_link = &run.sched;
    makecontext(&f.uc, (void (*)())fiber_entry, 0);
  }
  unsigned live = n;
  while (live) {
    unsigned long before = run.progress;
    for (unsigned i = 0; i < n; ++i) {
      Fiber& f = run.fibers[i];
      if (f.done) continue;
      run.cur = &f;
      swapcontext(&run.sched, &f.uc);
      if (f.done) --live;
    }
    if (live && run.progress == before) {
      fprintf(stderr, "[cuda_emu] deadlock in block (%u,%u,%u): %u threads stuck at a collective\n",
              bid.x, bid.y, bid.z, live);
      abort();
    }
  }
  for (unsigned i = 0; i < n; ++i) t_stacks.put(run.fibers[i].stack);
  run.cur = nullptr;
  t_run = nullptr;
}

static int pool_size() {
  static int n = [] {
    const char* e = getenv("OJB_EMU_THREADS");
    int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
    return v < 1 ? 1 : (v > 64 ? 64 : v);
  }();
  return n;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
  if (total == 0) return;
  std::atomic<unsigned long long> next(0);
  auto worker = [&]() {
    for (;;) {
      unsigned long long b = next.fetch_add(1);
      if (b >= total) break;
      uint3 bid{ (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y)) };
      run_block(grid, block, bid, smem, body);
    }
  };
  int nt = (int)std::min<unsigned long long>((unsigned long long)pool_size(), total);
  if (nt <= 1) { worker(); return; }
  std::vector<std::thread> th;
  for (int i = 0; i < nt; ++i) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

} // namespace ojb_emu

// ---- runtime API subset (device memory == host memory; everything is synchronous) --------
struct emuStream_st { int dummy; };
struct emuEvent_st { std::chrono::steady_clock::time_point t; };
// Fresh device memory is filled with a poison pattern (a kernel that reads what nobody wrote shows up
// as a test failure instead of passing by luck on a zeroed heap) and sits between two guard zones that
// are checked when it is freed (a kernel that writes just outside its buffer aborts the test run).
static const size_t kGuard = 4096;
struct AllocHeader { size_t size; unsigned long long magic; };
cudaError_t cudaMalloc(void** p, size_t n) {
  const size_t sz = (n + 255) & ~(size_t)255;
  unsigned char* raw = (unsigned char*)aligned_alloc(4096, sz + 2 * kGuard);
  if (!raw) { *p = nullptr; return cudaErrorMemoryAllocation; }
  static const int pat = [] { const char* e = getenv("OJB_EMU_POISON"); return e ? (int)strtol(e, nullptr, 0) : 0xA5; }();
  memset(raw, 0xEE, kGuard);
  memset(raw + kGuard, pat, sz);
  memset(raw + kGuard + sz, 0xEE, kGuard);
  AllocHeader* h = (AllocHeader*)raw; h->size = sz; h->magic = 0x0A110CA7EDull;
  *p = raw + kGuard;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
  if (!p) return cudaSuccess;
  unsigned char* raw = (unsigned char*)p - kGuard;
  AllocHeader* h = (AllocHeader*)raw;
  if (h->magic != 0x0A110CA7EDull) { fprintf(stderr, "[cuda_emu] free of a pointer cudaMalloc did not return, or its header was overwritten\n"); abort(); }
  const size_t sz = h->size;
  for (size_t i = sizeof(AllocHeader); i < kGuard; ++i)
    if (raw[i] != 0xEE) { fprintf(stderr, "[cuda_emu] write BEFORE a device buffer of %zu bytes (guard offset -%zu)\n", sz, kGuard - i); abort(); }
  for (size_t i = 0; i < kGuard; ++i)
    if (raw[kGuard + sz + i] != 0xEE) { fprintf(stderr, "[cuda_emu] write PAST a device buffer of %zu bytes (+%zu)\n", sz, i); abort(); }
  free(raw);
  return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t y = 0; y < height; ++y) memmove((char*)d + y * dpitch, (const char*)s + y * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = new emuStream_st(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream_st(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent_st(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent_st(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{ *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; strcpy(p->name, "cuda_emu (CPU fibers)");
  p->totalGlobalMem = (size_t)8 << 30; p->major = 10; p->minor = 0; return cudaSuccess;
}

// cuda_emu.h — deterministic single-threaded SIMT emulator (DEVELOPMENT / TEST TOOLING ONLY).
//
// Compiles armada_b200/csrc/armada_round.cu with plain g++ (-DARMADA_EMU) so the *same* kernel
// source can be stepped through the parity suite on a machine without a GPU before GPU minutes are
// spent.  It is NOT a product path and NOT a CPU fallback: the product library
// (libarmada_b200.so) is built by nvcc only and fails loudly without a CUDA device; the library
// produced from this header is loaded exclusively by tests/emu_lib.py.
//
// Model: every CUDA thread is a user-level coroutine on one host thread; blocks run one after the
// other, warps of a block are interleaved at warp collectives / barriers / explicit yields.
// Warp collectives require all 32 lanes of the warp (full mask), like the kernels use them.
// Memory is sequentially consistent, so this finds logic errors, not memory-ordering races.
#pragma once
#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define ARMADA_NOINLINE __attribute__((noinline))
#define ARMADA_PREFETCH_L1(p) ((void)(p))
#define __align__(n) alignas(n)
#define __shared__ static
#define __constant__ static
#define __restrict__

struct uint4 {
  unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct uint2 {
  unsigned x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
struct EmuEvent {
  std::chrono::steady_clock::time_point t;
};
typedef EmuEvent* cudaEvent_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {

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

enum State { RUNNABLE, WAIT_WARP, WAIT_BAR, DONE };

struct Warp {
  uint64_t vals[2][32];
  void* site[2][32];  // call site of each lane's collective (divergence check)
  int count[2] = {0, 0};
  uint64_t done = 0;  // completed generations
  int nlanes = 32;
};
struct Barrier {
  unsigned count = 0;
  uint64_t gen = 0;
};
struct Thread {
  void* sp = nullptr;
  State state = RUNNABLE;
  uint3 tidx{0, 0, 0};
  Warp* warp = nullptr;
  int lane = 0;
  uint64_t gen = 0;       // next warp-collective generation of this lane
  uint64_t wait_gen = 0;  // generation waited for (warp or barrier)
  int wait_bar = 0;
  int mark = 0;  // last ARMADA_EMU_MARK() passed (debug aid for the watchdog)
  char* stack = nullptr;
};

struct Machine {
  std::vector<Thread> threads;
  std::vector<Warp> warps;
  Barrier bars[16];
  void* sched_sp = nullptr;
  Thread* cur = nullptr;
  uint3 block_idx{0, 0, 0};
  dim3 block_dim, grid_dim;
  unsigned char* dyn_smem = nullptr;
  size_t dyn_cap = 0;
  char* stacks = nullptr;
  size_t stack_bytes = 256 * 1024, stacks_cap = 0;
  std::function<void()> body;
  long long clock = 0;
  bool reverse = false;
};
inline Machine& M() {
  static Machine m;
  return m;
}

inline void yield_to_scheduler() {
  Machine& m = M();
  emu_switch(&m.cur->sp, m.sched_sp);
}
static void trampoline() {
  Machine& m = M();
  m.body();
  m.cur->state = DONE;
  yield_to_scheduler();
  abort();
}

inline void run_block() {
  Machine& m = M();
  const unsigned B = m.block_dim.x;
  if (m.threads.size() < B) m.threads.resize(B);
  if (m.warps.size() < (B + 31) / 32) m.warps.resize((B + 31) / 32);
  if (m.stacks_cap < (size_t)B * m.stack_bytes) {
    if (m.stacks) munmap(m.stacks, m.stacks_cap);
    m.stacks_cap = (size_t)B * m.stack_bytes;
    m.stacks = (char*)mmap(nullptr, m.stacks_cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m.stacks == MAP_FAILED) abort();
  }
  for (auto& b : m.bars) b = Barrier{};
  for (unsigned w = 0; w < (B + 31) / 32; ++w) {
    m.warps[w] = Warp{};
    m.warps[w].nlanes = (int)std::min(32u, B - w * 32);
  }
  for (unsigned t = 0; t < B; ++t) {
    Thread& th = m.threads[t];
    th = Thread{};
    th.tidx = uint3{t, 0, 0};
    th.warp = &m.warps[t / 32];
    th.lane = (int)(t % 32);
    th.stack = m.stacks + (size_t)t * m.stack_bytes;
    uintptr_t top = ((uintptr_t)th.stack + m.stack_bytes) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                // fake return address of trampoline
    *--sp = (void*)&trampoline;     // `ret` target
    for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12..r15
    th.sp = (void*)sp;
  }
  unsigned live = B;
  unsigned long long rounds = 0;
  const auto t_start = std::chrono::steady_clock::now();
  const double limit_s = getenv("EMU_LAUNCH_TIMEOUT") ? atof(getenv("EMU_LAUNCH_TIMEOUT")) : 120.0;
  while (live) {
    if ((++rounds & 0xFFFF) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > limit_s) {
      fprintf(stderr, "simt_emu: kernel launch exceeded %.0f s (livelock?) in block %u; thread states:\n", limit_s, m.block_idx.x);
      for (unsigned t = 0; t < B; t += 32) {
        fprintf(stderr, "  warp %u done=%llu lanes(state:gen):", t / 32, (unsigned long long)m.threads[t].warp->done);
        for (unsigned i = t; i < t + 32 && i < B; ++i) fprintf(stderr, " %d:%llu@%d", (int)m.threads[i].state, (unsigned long long)m.threads[i].gen, m.threads[i].mark);
        fprintf(stderr, "\n    pending collective sites:");
        for (int i = 0; i < 32; ++i) fprintf(stderr, " %p", m.threads[t].warp->site[m.threads[t].warp->done & 1][i]);
        fprintf(stderr, "\n");
      }
      abort();
    }
    bool progressed = false;
    for (unsigned tt = 0; tt < B; ++tt) {
      // EMU_ORDER=reverse runs the lanes of every warp from 31 down to 0: code that is only correct
      // under one lane order (a missing __syncwarp) fails under the other.
      unsigned t = m.reverse ? (tt & ~31u) + (std::min(B - (tt & ~31u), 32u) - 1 - (tt & 31u)) : tt;
      Thread& th = m.threads[t];
      if (th.state == DONE) continue;
      if (th.state == WAIT_WARP && th.warp->done <= th.wait_gen) continue;
      if (th.state == WAIT_BAR && m.bars[th.wait_bar].gen == th.wait_gen) continue;
      th.state = RUNNABLE;
      m.cur = &th;
      emu_switch(&m.sched_sp, th.sp);
      progressed = true;
      if (th.state == DONE) --live;
    }
    if (!progressed) {
      fprintf(stderr, "simt_emu: deadlock in block %u (%u live threads)\n", m.block_idx.x, live);
      abort();
    }
  }
}

static void segv_handler(int, siginfo_t* si, void*) {
  void* bt[64];
  int n = backtrace(bt, 64);
  Machine& m = M();
  fprintf(stderr, "simt_emu: SIGSEGV addr=%p block=%u thread=%u stack=[%p,%p)\n", si->si_addr, m.block_idx.x,
          m.cur ? m.cur->tidx.x : 0u, m.cur ? (void*)m.cur->stack : nullptr, m.cur ? (void*)(m.cur->stack + m.stack_bytes) : nullptr);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
inline void install_segv() {
  static bool done = false;
  if (done || !getenv("EMU_TRACE")) return;
  done = true;
  static char alt[1 << 16];
  stack_t ss{};
  ss.ss_sp = alt;
  ss.ss_size = sizeof(alt);
  sigaltstack(&ss, nullptr);
  struct sigaction sa{};
  sa.sa_sigaction = segv_handler;
  sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  Machine& m = M();
  install_segv();
  m.reverse = getenv("EMU_ORDER") && !strcmp(getenv("EMU_ORDER"), "reverse");
  m.grid_dim = grid;
  m.block_dim = block;
  if (m.dyn_cap < smem + 64) {
    free(m.dyn_smem);
    m.dyn_cap = smem + 64;
    m.dyn_smem = (unsigned char*)aligned_alloc(128, (m.dyn_cap + 127) & ~(size_t)127);
  }
  m.body = std::function<void()>(f);
  for (unsigned b = 0; b < grid.x; ++b) {
    m.block_idx = uint3{b, 0, 0};
    memset(m.dyn_smem, 0xCD, smem);  // uninitialised shared memory is garbage on the device too
    run_block();
  }
}

// ---- warp collectives (all lanes of the warp participate) ---------------------------------------
__attribute__((noinline)) inline const uint64_t* collect(uint64_t v) {
  Machine& m = M();
  Thread* t = m.cur;
  Warp* W = t->warp;
  uint64_t g = t->gen++;
  int b = (int)(g & 1);
  W->vals[b][t->lane] = v;
  W->site[b][t->lane] = __builtin_return_address(0);
  if (++W->count[b] == W->nlanes) {
    W->count[b] = 0;
    W->done = g + 1;
  } else {
    t->state = WAIT_WARP;
    t->wait_gen = g;
    yield_to_scheduler();
  }
  m.clock += 1;
  return W->vals[b];
}
inline void barrier(int id, unsigned expected) {
  Machine& m = M();
  Thread* t = m.cur;
  Barrier& b = m.bars[id];
  if (++b.count == expected) {
    b.count = 0;
    b.gen++;
  } else {
    t->state = WAIT_BAR;
    t->wait_bar = id;
    t->wait_gen = b.gen;
    yield_to_scheduler();
  }
}
template <class T>
inline uint64_t to_bits(T v) {
  uint64_t b = 0;
  static_assert(sizeof(T) <= 8, "collective value too wide");
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
}  // namespace emu

#define threadIdx (emu::M().cur->tidx)
#define blockIdx (emu::M().block_idx)
#define blockDim (emu::M().block_dim)
#define gridDim (emu::M().grid_dim)

// explicit cooperative yield inside polling loops (no-op on the device)
static inline void armada_emu_yield() { emu::yield_to_scheduler(); }

static inline void __syncthreads() { emu::barrier(0, blockDim.x); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::collect(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline long long clock64() { return emu::M().clock += 7; }
static inline void __nanosleep(unsigned) { armada_emu_yield(); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  const uint64_t* a = emu::collect(emu::to_bits(v));
  return emu::from_bits<T>(a[src & 31]);
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m) {
  int lane = emu::M().cur->lane;
  const uint64_t* a = emu::collect(emu::to_bits(v));
  return emu::from_bits<T>(a[(lane ^ m) & 31]);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  int lane = emu::M().cur->lane;
  const uint64_t* a = emu::collect(emu::to_bits(v));
  return emu::from_bits<T>(a[lane >= (int)d ? lane - (int)d : lane]);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  int lane = emu::M().cur->lane;
  const uint64_t* a = emu::collect(emu::to_bits(v));
  return emu::from_bits<T>(a[lane + (int)d < 32 ? lane + (int)d : lane]);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  const uint64_t* a = emu::collect(pred ? 1 : 0);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (a[i] ? 1u : 0u) << i;
  return r;
}
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xFFFFFFFFu; }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  const uint64_t* a = emu::collect(v);
  unsigned r = 0xFFFFFFFFu;
  for (int i = 0; i < 32; ++i) r = (unsigned)a[i] < r ? (unsigned)a[i] : r;
  return r;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  const uint64_t* a = emu::collect(v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r = (unsigned)a[i] > r ? (unsigned)a[i] : r;
  return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  const uint64_t* a = emu::collect(v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r += (unsigned)a[i];
  return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  const uint64_t* a = emu::collect(v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)a[i];
  return r;
}

template <class T, class U>
static inline T atomicAdd(T* p, U v) {
  T o = *p;
  *p = (T)(o + (T)v);
  return o;
}
template <class T, class U>
static inline T atomicOr(T* p, U v) {
  T o = *p;
  *p = (T)(o | (T)v);
  return o;
}
template <class T, class U>
static inline T atomicAnd(T* p, U v) {
  T o = *p;
  *p = (T)(o & (T)v);
  return o;
}
template <class T, class U>
static inline T atomicMin(T* p, U v) {
  T o = *p;
  if ((T)v < o) *p = (T)v;
  return o;
}
template <class T, class U>
static inline T atomicMax(T* p, U v) {
  T o = *p;
  if ((T)v > o) *p = (T)v;
  return o;
}
template <class T, class U>
static inline T atomicExch(T* p, U v) {
  T o = *p;
  *p = (T)v;
  return o;
}
template <class T>
static inline T __ldcg(const T* p) { return *p; }
template <class T>
static inline T __ldg(const T* p) { return *p; }

static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
static inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}

// ---- CUDA runtime shims ----------------------------------------------------------------------------
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
// allocations carry a canary so that a kernel writing past the end of a buffer is reported
struct EmuAlloc {
  unsigned char* p;
  size_t n;
};
static inline std::vector<EmuAlloc>& emu_allocs() {
  static std::vector<EmuAlloc> v;
  return v;
}
static inline void emu_check_canaries(const char* where) {
  for (auto& a : emu_allocs())
    for (size_t i = 0; i < 64; ++i)
      if (a.p[a.n + i] != 0x5C) {
        fprintf(stderr, "simt_emu: buffer overrun past a %zu-byte allocation (offset +%zu) detected after %s\n", a.n, i, where);
        abort();
      }
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  unsigned char* q = (unsigned char*)malloc(n + 64);
  if (!q) return 2;
  memset(q, getenv("EMU_POISON") ? atoi(getenv("EMU_POISON")) : 0xA5, n);  // device allocations are not zeroed
  memset(q + n, 0x5C, 64);
  emu_allocs().push_back(EmuAlloc{q, n});
  *p = q;
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) {
  auto& v = emu_allocs();
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i].p == p) {
      v[i] = v.back();
      v.pop_back();
      break;
    }
  free(p);
  return cudaSuccess;
}
enum { cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void** p, size_t n, int) {
  *p = malloc(n);
  return *p ? cudaSuccess : 2;
}
static inline cudaError_t cudaFreeHost(void* p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, int) {
  *d = h;
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) {
  memcpy(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) {
  memcpy(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
  memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  memset(d, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) {
  *s = (void*)1;
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new EmuEvent();
  return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
  e->t = std::chrono::steady_clock::now();
  return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  *v = attr == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448 : 0;
  return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMemcpyToSymbolAsync(T& sym, const void* src, size_t n, size_t off, int, cudaStream_t) {
  memcpy((char*)&sym + off, src, n);
  return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

#define ARMADA_LAUNCH(kern, grid, block, smem, stream, ...) \
  if (getenv("EMU_TRACE")) fprintf(stderr, "emu launch %s\n", #kern); \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); }); \
  emu_check_canaries(#kern)
#define ARMADA_EMU_MARK(id) (emu::M().cur->mark = (id))
#define ARMADA_DYN_SMEM(name) unsigned char* name = emu::M().dyn_smem
#define ARMADA_NAMED_BARRIER(id, count) emu::barrier((id), (count))

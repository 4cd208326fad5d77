// Shared host-side plumbing for libqwen3tts_b200: status-carrying exception, CUDA error checks, device arena.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qwen3tts_b200.h"

namespace q3 {

struct Error : std::runtime_error {
  q3tts_status status;
  Error(q3tts_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

[[noreturn]] inline void fail(q3tts_status s, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(s, buf);
}

#define Q3_CUDA(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      ::q3::fail(Q3TTS_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                 cudaGetErrorString(_e));                                                              \
  } while (0)

#define Q3_CHECK(cond, status, ...) \
  do {                              \
    if (!(cond)) ::q3::fail(status, __VA_ARGS__); \
  } while (0)

inline size_t dtype_size(int dt) { return dt == Q3TTS_F32 ? 4 : 2; }

// Programmatic dependent launch (PDL): a kernel launched with `pdl` may start (prologue, weight prefetch) while its
// predecessor in the stream is still draining; it calls pdl_wait() before touching anything the predecessor wrote.  Kernels
// call pdl_launch_dependents() early so their successor can be scheduled.  Both instructions are no-ops without the attribute.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline void launch_kernel_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  Q3_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("Q3TTS_PDL"); return !(e && atoi(e) == 0); }();
  return on;
}
#endif

// ---- chain signals: flag hand-off between consecutive kernels of a decode step -------------------------------------------------
// A kernel launched with programmatic dependent launch is resident (weights prefetched, operands dequantised) long before its
// predecessor retires, but griddepcontrol.wait only returns once the WHOLE predecessor grid has completed and flushed.  Chained
// kernels hand over earlier: every CTA of the producer, after its last global store, does fence + one increment of the launch's
// counter; the consumer spins on that counter (acquire) until all producer CTAs have signalled, and never executes
// griddepcontrol.wait.  Safe because a dependent grid only starts once ALL CTAs of its primary are resident (they called
// launch_dependents), so a spinning consumer can never keep its producer from running; every CTA signals after ALL its reads and
// writes, so later launches may reuse the producer's input buffers; visibility to later launches follows from the
// release/acquire chain.  Counters are zeroed by a memset node at the head of every frame graph.
struct ChainSig {
  unsigned* out = nullptr;        // incremented once per CTA of this launch
  const unsigned* in = nullptr;   // non-null: wait for *in >= in_target instead of griddepcontrol.wait
  unsigned in_target = 0;
};
#ifdef __CUDACC__
__device__ __forceinline__ void chain_signal(unsigned* sig) {  // ONE thread, after a CTA barrier that follows the CTA's last store
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sig) : "memory");
}
__device__ __forceinline__ void chain_wait(const unsigned* sig, unsigned target) {
  unsigned v, spins = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sig) : "memory");
    if (v >= target) break;
    if (++spins > (1u << 25)) __trap();  // a protocol bug must surface as an error, never as a hung GPU
  }
}
#endif
struct ChainState {
  unsigned* base = nullptr;       // device counters of one frame graph
  int capacity = 0, next = 0;
  const unsigned* prev = nullptr; // counter of the launch issued last, when that launch signals
  unsigned prev_ctas = 0;
};

// Counts every kernel this library launches (reported as q3tts_timing.kernel_launches / bench gpu_launches).
struct LaunchCounter {
  int64_t n = 0;
  bool capturing = false;   // while a CUDA graph is being captured launches are counted into `captured`
  int64_t captured = 0;
  inline void tick() { if (capturing) ++captured; else ++n; }
};

// Bump allocator over cudaMalloc'ed slabs: weights and state live for the handle's lifetime.
class DeviceArena {
 public:
  ~DeviceArena() { release(); }
  void* alloc(size_t bytes, size_t align = 256) {
    if (bytes == 0) bytes = align;
    size_t off = (used_ + align - 1) / align * align;
    if (slabs_.empty() || off + bytes > cap_) {
      size_t sz = bytes > kSlab ? bytes : kSlab;
      void* p = nullptr;
      Q3_CUDA(cudaMalloc(&p, sz));
      slabs_.push_back(p);
      cap_ = sz;
      used_ = 0;
      off = 0;
      total_ += sz;
    }
    used_ = off + bytes;
    return static_cast<char*>(slabs_.back()) + off;
  }
  template <typename T>
  T* alloc_n(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  void release() {
    for (void* p : slabs_) cudaFree(p);
    slabs_.clear();
    cap_ = used_ = total_ = 0;
  }
  size_t total() const { return total_; }

 private:
  static constexpr size_t kSlab = size_t(256) << 20;
  std::vector<void*> slabs_;
  size_t cap_ = 0, used_ = 0, total_ = 0;
};

}  // namespace q3

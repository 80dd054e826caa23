// Shared helpers for libbvg_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <utility>

#include "../../include/bvg_b200.h"

namespace bvg {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BVG_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::bvg::cuda_fail(_e, #expr); \
  } while (0)

#define BVG_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::bvg::set_error(__VA_ARGS__);  \
      return BVG_EINVAL;              \
    }                                 \
  } while (0)

// True exactly once per (current device, kernel): function attributes (dynamic shared-memory opt-in, carve-out
// preference) are per device, so a process that drives several GPUs must set them on each.  Thread-safe.
bool first_use_on_device(const void* kernel);

// the knobs of a call: the caller's bvg_tuning or the defaults
inline bvg_tuning default_tuning() {
  bvg_tuning t = {};
  t.amp_mma = 1;
  t.amp_packed = 1;
  t.amp_ct = 1;
  t.umma_ntile_cap = 256;
  t.umma_stack = 128;
  t.umma_pair = 1;
  return t;
}
inline bvg_tuning tune_of(const bvg_tuning* t) { return t ? *t : default_tuning(); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A program (bvg_program) is a fixed chain of dependent launches, many of them a few microseconds long (a DiffSVC
// step: 106 launches, 1.8 ms).  With the programmaticStreamSerialization launch attribute a kernel's CTAs may be
// scheduled while its predecessor still runs: every kernel of the chain fires `griddepcontrol.launch_dependents` on
// entry, sets itself up (barriers, TMEM, tensor-map prefetch, index math), and executes `griddepcontrol.wait` --
// which returns once the predecessor grid has completed and its writes are visible -- before its first access to
// global memory.  Launch latency and prologue then overlap the predecessor instead of following it.  The attribute
// is only ever given to kernels that contain the wait; `pdl_mode()` is thread-local, set by bvg_program_run for the
// duration of the call when the program asks for it (bvg_program_set_pdl) and false everywhere else.
bool& pdl_mode();
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_mode()) { pdl_mode() = on; }
  ~PdlScope() { pdl_mode() = prev; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_mode() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- device-side element helpers ---------------------------------------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

// round-to-nearest-even fp32 -> bf16, returned as the high 16 bits
__device__ __forceinline__ uint32_t float_to_bf16_bits(float f) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}

// pack two floats into one 32-bit word of two bf16 (lo half = first)
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float& a, float& b) {
  a = __uint_as_float(w << 16);
  b = __uint_as_float(w & 0xffff0000u);
}

// hi = bf16(x), lo = bf16(x - hi): the two planes of a SPLIT tensor
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(x));
  lo = x - hi;  // exact in fp32; rounded to bf16 when packed
}

}  // namespace bvg

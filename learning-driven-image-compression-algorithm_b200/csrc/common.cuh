// Shared helpers for libldic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>

#include "../../include/ldic.h"

namespace ldic {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;
extern std::mutex g_init_mu;   // guards the lazily initialised per-kernel state (function attributes, occupancy, debug buffers)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LDIC_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return LDIC_OK;
}

#define LDIC_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) return ::ldic::fail(LDIC_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

constexpr int kMaxSMs = 160;      // sizing bound of per-CTA workspaces (B200: 148); grids use num_sms()
constexpr int kMaxDevices = 64;
int current_device();             // cudaGetDevice (-1 on error)
int num_sms();                    // multiprocessor count of the current device (queried once per device)

// Tuning / diagnostic switches, read from the environment once at load time (see ldic_set_tuning in ldic.h)
struct Tuning {
  int debug_nostore, debug_timing, gdn_insert, stages_cap, tail_wide, lik_grid, first_epi, first_insert, first_tma_store;
  unsigned epoch;                 // bumped by ldic_set_tuning: cached launch plans of older epochs are not reused
};
const Tuning& tuning();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// a11 arithmetic on one sample (model/net.py:864-868): squared difference of the 8-bit levels, exact integer
__device__ __forceinline__ unsigned int sq_level_err(float x, float xt, int clamp_pm1) {
  if (clamp_pm1) xt = fminf(fmaxf(xt, -1.f), 1.f);
  float gt = rintf(__fmul_rn(__fadd_rn(x, 1.f), 127.5f));
  float xh = __fmul_rn(__fadd_rn(xt, 1.f), 127.5f);
  xh = rintf(fminf(fmaxf(xh, 0.f), 255.f));
  float d = xh - gt;   // integers: exact
  return (unsigned int)(d * d);
}

}  // namespace ldic

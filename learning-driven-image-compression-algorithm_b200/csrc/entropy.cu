// Memory-bound kernels of the rate-distortion path: LowerBound / reparametrisation,
// quantise + Gaussian likelihood + sum(ln L), MSE on 8-bit levels, layout glue.
// sm_100a; every kernel is a coalesced, vectorised streaming pass.
#include "common.cuh"
#include <stdlib.h>
#include <type_traits>
#include <string.h>

namespace ldic {
thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};
std::mutex g_init_mu;

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  return dev;
}
int num_sms() {
  static std::atomic<int> sms[kMaxDevices];
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    if (n > kMaxSMs) n = kMaxSMs;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static Tuning g_tuning = [] {
  auto env = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  Tuning t;
  t.debug_nostore = getenv("LDIC_DEBUG_NOSTORE") ? (atoi(getenv("LDIC_DEBUG_NOSTORE")) ? atoi(getenv("LDIC_DEBUG_NOSTORE")) : 1) : 0;
  t.debug_timing = getenv("LDIC_DEBUG_TIMING") != nullptr;
  t.gdn_insert = env("LDIC_GDN_INSERT", 0);
  t.stages_cap = env("LDIC_STAGES", 0);
  t.tail_wide = env("LDIC_TAIL_WIDE", 1);
  t.lik_grid = env("LDIC_LIK_GRID", 0);
  t.first_epi = env("LDIC_FIRST_EPI", 0);      // epilogue warps of the first-layer kernel: 0 = default (12 at 192 channels), 8
  // first-layer kernel: how many of tile t+1's two conv stages the MMA warp issues before the GDN contraction of tile t
  // (2 = default; 0 = the contraction never waits for the patch builders: measured no faster, 0.380 vs 0.378 ms uint8,
  // 0.346 vs 0.335 ms fp32 -- the wait for the contraction is its own ~1.5 k cycles, not the issue order)
  t.first_insert = env("LDIC_FIRST_INSERT", 2);
  t.first_tma_store = env("LDIC_FIRST_TMA_STORE", 1);   // first layer: output tile through shared memory + TMA stores
  t.epoch = 0;
  return t;
}();
const Tuning& tuning() { return g_tuning; }
}  // namespace ldic

using namespace ldic;

extern "C" int ldic_version(void) { return 100; }
extern "C" const char* ldic_last_error(void) { return g_err; }
extern "C" long long ldic_launch_count(void) { return g_launches.load(); }
extern "C" int ldic_set_tuning(const char* key, int value) {
  if (!key) return fail(LDIC_EINVAL, "set_tuning: null key");
  Tuning& t = const_cast<Tuning&>(tuning());
  int* f = nullptr;
  if (!strcmp(key, "debug_nostore")) f = &t.debug_nostore;
  else if (!strcmp(key, "debug_timing")) f = &t.debug_timing;
  else if (!strcmp(key, "gdn_insert")) f = &t.gdn_insert;
  else if (!strcmp(key, "stages")) f = &t.stages_cap;
  else if (!strcmp(key, "tail_wide")) f = &t.tail_wide;
  else if (!strcmp(key, "lik_grid")) f = &t.lik_grid;
  else if (!strcmp(key, "first_epi")) f = &t.first_epi;
  else if (!strcmp(key, "first_insert")) f = &t.first_insert;
  else if (!strcmp(key, "first_tma_store")) f = &t.first_tma_store;
  if (!f) return fail(LDIC_EINVAL, "set_tuning: unknown key '%s'", key);
  std::lock_guard<std::mutex> lk(g_init_mu);
  const int old = *f;
  *f = value;
  ++t.epoch;
  return old;
}

extern "C" int ldic_check_device(int dev) {
  cudaDeviceProp p;
  LDIC_CUDA(cudaGetDeviceProperties(&p, dev));
  if (p.major != 10) return fail(LDIC_ENOTSUP, "device %d is sm_%d%d, libldic_b200 needs sm_100", dev, p.major, p.minor);
  return LDIC_OK;
}

// ------------------------------------------------------------------------------------
// a3: LowerBound, NonNegativeParametrizer
// ------------------------------------------------------------------------------------
// torch.max propagates NaN; (x < b ? b : x) keeps a NaN x, unlike fmaxf.
__device__ __forceinline__ float lower_bound_f(float x, float b) { return (x < b) ? b : x; }

__global__ void k_lower_bound(const float* __restrict__ x, float bound, float* __restrict__ y, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) y[i] = lower_bound_f(x[i], bound);
}
__global__ void k_lower_bound_bwd(const float* __restrict__ x, float bound, const float* __restrict__ g,
                                  float* __restrict__ gi, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) gi[i] = ((x[i] >= bound) || (g[i] < 0.f)) ? g[i] : 0.f * g[i];
}
__global__ void k_nonneg(const float* __restrict__ p, float bound, float pedestal, float* __restrict__ o, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    float t = lower_bound_f(p[i], bound);
    o[i] = __fsub_rn(__fmul_rn(t, t), pedestal);
  }
}
static inline int grid_for(size_t n, int threads = 256, int max_blocks = 0) {
  if (max_blocks <= 0) max_blocks = num_sms() * 8;
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > (size_t)max_blocks ? max_blocks : b);
}

extern "C" int ldic_lower_bound(const float* x, float bound, float* y, size_t n, void* stream) {
  if (n == 0) return LDIC_OK;
  k_lower_bound<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, bound, y, n);
  return check_launch("k_lower_bound");
}
extern "C" int ldic_lower_bound_bwd(const float* x, float bound, const float* g, float* gi, size_t n, void* stream) {
  if (n == 0) return LDIC_OK;
  k_lower_bound_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, bound, g, gi, n);
  return check_launch("k_lower_bound_bwd");
}
extern "C" int ldic_nonneg_reparam(const float* p, float bound, float pedestal, float* out, size_t n, void* stream) {
  if (n == 0) return LDIC_OK;
  k_nonneg<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(p, bound, pedestal, out, n);
  return check_launch("k_nonneg");
}

// beta_eff / gamma_eff + the bf16 block-diagonal operand image for the tensor-core epilogue.
__global__ void k_gdn_prepare(const float* __restrict__ beta_p, const float* __restrict__ gamma_p, int C,
                              float bb, float gb, float ped, float* __restrict__ beta_eff,
                              float* __restrict__ gamma_eff, __nv_bfloat16* __restrict__ gamma_bf16,
                              float* __restrict__ beta_tiled, int groups, int Np, int Kp) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = tid; i < C; i += st) {
    float t = lower_bound_f(beta_p[i], bb);
    float b = __fsub_rn(__fmul_rn(t, t), ped);
    if (beta_eff) beta_eff[i] = b;
  }
  for (int i = tid; i < C * C; i += st) {
    float t = lower_bound_f(gamma_p[i], gb);
    float g = __fsub_rn(__fmul_rn(t, t), ped);
    if (gamma_eff) gamma_eff[i] = g;
  }
  if (gamma_bf16) {
    for (int i = tid; i < Np * Kp; i += st) {
      int r = i / Kp, k = i % Kp;
      float g = 0.f;
      if (r < groups * C && k < groups * C && (r / C) == (k / C)) {
        float t = lower_bound_f(gamma_p[(r % C) * C + (k % C)], gb);
        g = __fsub_rn(__fmul_rn(t, t), ped);
      }
      gamma_bf16[i] = __float2bfloat16_rn(g);
    }
  }
  if (beta_tiled) {
    for (int i = tid; i < Np; i += st) {
      float b = 1.f;
      if (i < groups * C) {
        float t = lower_bound_f(beta_p[i % C], bb);
        b = __fsub_rn(__fmul_rn(t, t), ped);
      }
      beta_tiled[i] = b;
    }
  }
}
extern "C" int ldic_gdn_prepare(const float* beta_p, const float* gamma_p, int C, float beta_bound, float gamma_bound,
                                float pedestal, float* beta_eff, float* gamma_eff, void* gamma_bf16,
                                float* beta_tiled, int groups, int Np, int Kp, void* stream) {
  if (C <= 0) return fail(LDIC_EINVAL, "gdn_prepare: C=%d", C);
  if (gamma_bf16 && (groups * C > Np || groups * C > Kp)) return fail(LDIC_EINVAL, "gdn_prepare: groups*C exceeds Np/Kp");
  size_t n = (size_t)C * C;
  if (gamma_bf16 && (size_t)Np * Kp > n) n = (size_t)Np * Kp;
  k_gdn_prepare<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(beta_p, gamma_p, C, beta_bound, gamma_bound, pedestal,
                                                                beta_eff, gamma_eff, (__nv_bfloat16*)gamma_bf16,
                                                                beta_tiled, groups, Np, Kp);
  return check_launch("k_gdn_prepare");
}

// ------------------------------------------------------------------------------------
// a2: stand-alone GDN / IGDN, NCHW fp32 (module surface).  CUDA-core fp32: the
// fused tensor-core version lives in conv_tc.cu.  One block = 64 pixels x all C.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gdn_nchw(const float* __restrict__ x, const float* __restrict__ beta,
                                                  const float* __restrict__ gamma, float* __restrict__ y, int C,
                                                  long long HW, int inverse, int use_rsqrt) {
  extern __shared__ float xs[];  // [C][64] squares
  const int b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * 64;
  const float* xb = x + (long long)b * C * HW;
  float* yb = y + (long long)b * C * HW;
  for (int i = threadIdx.x; i < C * 64; i += 256) {
    int c = i >> 6, p = i & 63;
    float v = (p0 + p < HW) ? xb[(long long)c * HW + p0 + p] : 0.f;
    xs[i] = v * v;
  }
  __syncthreads();
  const int p = threadIdx.x & 63, g = threadIdx.x >> 6;
  if (p0 + p >= HW) return;
  for (int i = g; i < C; i += 4) {
    const float* gr = gamma + (long long)i * C;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = 0;
    for (; j + 3 < C; j += 4) {
      a0 = fmaf(__ldg(gr + j), xs[(j << 6) + p], a0);
      a1 = fmaf(__ldg(gr + j + 1), xs[((j + 1) << 6) + p], a1);
      a2 = fmaf(__ldg(gr + j + 2), xs[((j + 2) << 6) + p], a2);
      a3 = fmaf(__ldg(gr + j + 3), xs[((j + 3) << 6) + p], a3);
    }
    for (; j < C; ++j) a0 = fmaf(__ldg(gr + j), xs[(j << 6) + p], a0);
    float norm = ((a0 + a1) + (a2 + a3)) + __ldg(beta + i);
    float xv = xb[(long long)i * HW + p0 + p];
    float o;
    if (inverse) o = xv * __fsqrt_rn(norm);
    else if (use_rsqrt) o = xv * __fdiv_rn(1.0f, __fsqrt_rn(norm));
    else o = __fdiv_rn(xv, __fsqrt_rn(norm));
    yb[(long long)i * HW + p0 + p] = o;
  }
}
extern "C" int ldic_gdn_nchw_f32(const float* x, const float* beta_eff, const float* gamma_eff, float* y, int B, int C,
                                 int H, int W, int inverse, int use_rsqrt, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return (B == 0 || H == 0 || W == 0) ? LDIC_OK : fail(LDIC_EINVAL, "gdn: bad shape");
  size_t smem = (size_t)C * 64 * sizeof(float);
  if (smem > 200 * 1024) return fail(LDIC_EINVAL, "gdn: C=%d too large", C);
  if (smem > 48 * 1024) LDIC_CUDA(cudaFuncSetAttribute(k_gdn_nchw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 63) / 64), B);
  k_gdn_nchw<<<grid, 256, smem, (cudaStream_t)stream>>>(x, beta_eff, gamma_eff, y, C, HW, inverse, use_rsqrt);
  return check_launch("k_gdn_nchw");
}

// ------------------------------------------------------------------------------------
// a6+a7+a8+a9: quantise + Gaussian likelihood + sum(ln L), one streaming pass.
// ------------------------------------------------------------------------------------
constexpr int kLikThreads = 256;
constexpr int kLikMaxBlocks = kMaxSMs * 32;

struct LikWs {
  unsigned int ticket;
  unsigned int pad[15];
  double partial[kLikMaxBlocks];
};
extern "C" size_t ldic_likelihood_workspace_bytes(void) { return sizeof(LikWs); }

__device__ __forceinline__ float quantise(float v, float mu, int quant) {
  switch (quant) {
    case 1: return rintf(v);
    case 2: return __fadd_rn(rintf(__fsub_rn(v, mu)), mu);
    case 3: return __fadd_rn(__fsub_rn(rintf(v), v), v);
    default: return v;
  }
}

// form 0: model/net.py:272-286.  cdf(t) = 0.5*(1+erf(t/sqrt2)), difference, clamp(min).
// form 1: CompressAI GaussianConditional (erfc form on |v-mu|, sigma lower bound).
template <int FORM>
__device__ __forceinline__ float likelihood(float vhat, float mu, float sigma, float lik_bound, float scale_bound) {
  float d = __fsub_rn(vhat, mu);
  float l;
  if (FORM == 0) {
    float r = __frcp_rn(sigma);
    float tu = __fmul_rn(__fadd_rn(d, 0.5f), r);
    float tl = __fmul_rn(__fsub_rn(d, 0.5f), r);
    const float inv_sqrt2 = 0.70710678118654752440f;
    float cu = __fmul_rn(0.5f, __fadd_rn(1.0f, erff(__fmul_rn(tu, inv_sqrt2))));
    float cl = __fmul_rn(0.5f, __fadd_rn(1.0f, erff(__fmul_rn(tl, inv_sqrt2))));
    l = __fsub_rn(cu, cl);
  } else {
    float a = fabsf(d);
    float s = lower_bound_f(sigma, scale_bound);
    float r = __frcp_rn(s);
    const float c = -0.70710678118654752440f;
    float tu = __fmul_rn(__fsub_rn(0.5f, a), r);
    float tl = __fmul_rn(__fsub_rn(-0.5f, a), r);
    float cu = __fmul_rn(0.5f, erfcf(__fmul_rn(c, tu)));
    float cl = __fmul_rn(0.5f, erfcf(__fmul_rn(c, tl)));
    l = __fsub_rn(cu, cl);
  }
  return lower_bound_f(l, lik_bound);  // NaN stays NaN, like torch.clamp / torch.max
}

struct LikParams {
  const float* v; long long v_rs, v_off;
  const float* mu; long long mu_rs, mu_off; int mu_mode;
  const float* sigma; long long sg_rs, sg_off; int sg_mode; int sg_period;
  long long rows, cols;
  int quant, sigma_is_log;
  float lik_bound, scale_bound;
  float* v_hat; long long vh_rs, vh_off;
  __nv_bfloat16* v_hat_bf16; long long vb_rs, vb_off;
  float* lik;
  float* sum_out;
  LikWs* ws;
};

template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&o)[VEC]) {
  if (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = t.x; o[1 % VEC] = t.y; o[2 % VEC] = t.z; o[3 % VEC] = t.w;
  } else {
    o[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&o)[VEC]) {
  if (VEC == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]));
  } else {
    p[0] = o[0];
  }
}

template <int VEC, int FORM>
__global__ void __launch_bounds__(kLikThreads) k_likelihood(LikParams P) {
  const long long colsv = P.cols / VEC;
  const long long total = P.rows * colsv;
  const long long stride = (long long)gridDim.x * kLikThreads;
  long long i = (long long)blockIdx.x * kLikThreads + threadIdx.x;
  long long row = i / colsv, cv = i - row * colsv;
  const long long step_r = stride / colsv, step_c = stride - step_r * colsv;
  float acc = 0.f;  // sum of log2(L) for this thread
  for (; i < total; i += stride) {
    const long long col = cv * VEC;
    float v[VEC], mu[VEC], sg[VEC], vh[VEC], lk[VEC];
    load_vec<VEC>(P.v + row * P.v_rs + P.v_off + col, v);
    if (P.mu_mode == 2) load_vec<VEC>(P.mu + row * P.mu_rs + P.mu_off + col, mu);
    else if (P.mu_mode == 1) load_vec<VEC>(P.mu + col, mu);
    else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) mu[k] = 0.f;
    }
    if (P.sg_mode == 2) load_vec<VEC>(P.sigma + row * P.sg_rs + P.sg_off + col, sg);
    else if (P.sg_mode == 1) load_vec<VEC>(P.sigma + col, sg);
    else {
      float s = __ldg(P.sigma + (row % P.sg_period));
#pragma unroll
      for (int k = 0; k < VEC; ++k) sg[k] = s;
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float s = P.sigma_is_log ? expf(sg[k]) : sg[k];
      vh[k] = quantise(v[k], mu[k], P.quant);
      lk[k] = likelihood<FORM>(vh[k], mu[k], s, P.lik_bound, P.scale_bound);
      acc += __log2f(lk[k]);
    }
    if (P.v_hat) store_vec<VEC>(P.v_hat + row * P.vh_rs + P.vh_off + col, vh);
    if (P.v_hat_bf16) {
      __nv_bfloat16* q = P.v_hat_bf16 + row * P.vb_rs + P.vb_off + col;
      if (VEC == 4) {
        uint2 t = make_uint2(pack_bf16x2(vh[0], vh[1 % VEC]), pack_bf16x2(vh[2 % VEC], vh[3 % VEC]));
        *reinterpret_cast<uint2*>(q) = t;
      } else {
        q[0] = __float2bfloat16_rn(vh[0]);
      }
    }
    if (P.lik) store_vec<VEC>(P.lik + row * P.cols + col, lk);
    cv += step_c; row += step_r;
    if (cv >= colsv) { cv -= colsv; ++row; }
  }
  // block reduction (warp shuffle -> smem -> double), then a deterministic last-block pass
  __shared__ float wsum[kLikThreads / 32];
  __shared__ bool is_last;
  float w = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kLikThreads / 32; ++k) s += (double)wsum[k];
    P.ws->partial[blockIdx.x] = s;
    __threadfence();
    unsigned int t = atomicAdd(&P.ws->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) s += *((volatile double*)&P.ws->partial[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      P.sum_out[0] = (float)(s * 0.69314718055994530942);  // sum ln L
      P.ws->ticket = 0;                                      // leave the workspace reusable
    }
  }
}

// ---- fast path: per-element mu and sigma, GaussianModel form, float4, no in-loop mode branches ----
// Branch-free erf on [-inf, inf]: erf(|t|) = 1 - 2^(-|t| g(|t|)), g = degree-8 polynomial fitted to
// -log2(erfc(t))/t on [0, 3.95] (max abs error 9.5e-8 = 1.6 ulp of 1.0, the same order as the
// CPU-erf / CUDA-erff disagreement; |t| is clamped at 3.95 where erf rounds to 1 in fp32).
__device__ __forceinline__ float erf_fast(float x) {
  float t = fabsf(x);
  t = (t > 3.95f) ? 3.95f : t;                  // keeps NaN (0/0 when sigma == 0), like torch.erf
  float p = -1.1605328836594708e-05f;
  p = fmaf(p, t, 0.00015298039943445474f);
  p = fmaf(p, t, -0.0008483482524752617f);
  p = fmaf(p, t, 0.0022751344367861748f);
  p = fmaf(p, t, -8.534445805707946e-05f);
  p = fmaf(p, t, -0.027724044397473335f);
  p = fmaf(p, t, 0.14830774068832397f);
  p = fmaf(p, t, 0.9184429049491882f);
  p = fmaf(p, t, 1.6279072761535645f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * p));
  return copysignf(1.0f - e, x);
}
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int QUANT, bool SIGLOG>
__device__ __forceinline__ float lik_fast(float v, float mu, float sg, float bound, float& vh) {
  vh = (QUANT == 1) ? rintf(v) : v;
  const float s = SIGLOG ? ex2_fast(sg * 1.4426950408889634f) : sg;
  const float d = __fsub_rn(vh, mu);
  const float r = rcp_fast(s) * 0.70710678118654752440f;
  const float cu = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_fast(__fadd_rn(d, 0.5f) * r)));
  const float cl = __fmul_rn(0.5f, __fadd_rn(1.0f, erf_fast(__fsub_rn(d, 0.5f) * r)));
  return lower_bound_f(__fsub_rn(cu, cl), bound);
}

template <int QUANT, bool SIGLOG>
__global__ void __launch_bounds__(kLikThreads) k_likelihood_fast(LikParams P) {
  const long long colsv = P.cols >> 2;
  const long long total = P.rows * colsv;
  const long long stride = (long long)gridDim.x * kLikThreads, end = total;
  long long i = (long long)blockIdx.x * kLikThreads + threadIdx.x;
  long long row = i / colsv, cv = i - row * colsv;
  const long long step_r = stride / colsv, step_c = stride - step_r * colsv;
  const float bound = P.lik_bound;
  float acc = 0.f;  // sum of log2(L) for this thread
  for (; i < end; i += stride) {
    const long long col = cv << 2;
    const float4 v4 = __ldcs(reinterpret_cast<const float4*>(P.v + row * P.v_rs + P.v_off + col));
    const float4 m4 = __ldcs(reinterpret_cast<const float4*>(P.mu + row * P.mu_rs + P.mu_off + col));
    const float4 s4 = __ldcs(reinterpret_cast<const float4*>(P.sigma + row * P.sg_rs + P.sg_off + col));
    float4 h4, l4;
    l4.x = lik_fast<QUANT, SIGLOG>(v4.x, m4.x, s4.x, bound, h4.x);
    l4.y = lik_fast<QUANT, SIGLOG>(v4.y, m4.y, s4.y, bound, h4.y);
    l4.z = lik_fast<QUANT, SIGLOG>(v4.z, m4.z, s4.z, bound, h4.z);
    l4.w = lik_fast<QUANT, SIGLOG>(v4.w, m4.w, s4.w, bound, h4.w);
    acc += (lg2_fast(l4.x) + lg2_fast(l4.y)) + (lg2_fast(l4.z) + lg2_fast(l4.w));
    if (P.v_hat) __stcs(reinterpret_cast<float4*>(P.v_hat + row * P.vh_rs + P.vh_off + col), h4);
    if (P.v_hat_bf16)
      *reinterpret_cast<uint2*>(P.v_hat_bf16 + row * P.vb_rs + P.vb_off + col) =
          make_uint2(pack_bf16x2(h4.x, h4.y), pack_bf16x2(h4.z, h4.w));
    if (P.lik) __stcs(reinterpret_cast<float4*>(P.lik + row * P.cols + col), l4);
    cv += step_c; row += step_r;
    if (cv >= colsv) { cv -= colsv; ++row; }
  }
  __shared__ float wsum[kLikThreads / 32];
  __shared__ bool is_last;
  float w = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kLikThreads / 32; ++k) s += (double)wsum[k];
    P.ws->partial[blockIdx.x] = s;
    __threadfence();
    unsigned int t = atomicAdd(&P.ws->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) s += *((volatile double*)&P.ws->partial[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      P.sum_out[0] = (float)(s * 0.69314718055994530942);  // sum ln L
      P.ws->ticket = 0;
    }
  }
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int ldic_round_likelihood_bpp(const LdicLikelihoodArgs* a, void* stream) {
  if (!a || !a->sum_ln_out) return fail(LDIC_EINVAL, "likelihood: null argument");
  if (a->rows < 0 || a->cols < 0) return fail(LDIC_EINVAL, "likelihood: negative shape");
  if (a->rows == 0 || a->cols == 0) {  // empty input: sum over nothing = 0 (torch.sum of an empty tensor)
    LDIC_CUDA(cudaMemsetAsync(a->sum_ln_out, 0, sizeof(float), (cudaStream_t)stream));
    return LDIC_OK;
  }
  if (!a->v || !a->sigma || !a->workspace) return fail(LDIC_EINVAL, "likelihood: null argument");
  if (a->mu_mode < 0 || a->mu_mode > 2 || a->sigma_mode < 1 || a->sigma_mode > 3) return fail(LDIC_EINVAL, "likelihood: bad broadcast mode");
  if (a->mu_mode != 0 && !a->mu) return fail(LDIC_EINVAL, "likelihood: mu is null");
  if (a->sigma_mode == 3 && a->sigma_period <= 0) return fail(LDIC_EINVAL, "likelihood: sigma_period");
  if (a->form < 0 || a->form > 1 || a->quant < 0 || a->quant > 3) return fail(LDIC_EINVAL, "likelihood: bad form/quant");
  cudaStream_t st = (cudaStream_t)stream;
  LikParams P;
  P.v = a->v; P.v_rs = a->v_rs; P.v_off = a->v_off;
  P.mu = a->mu; P.mu_rs = a->mu_rs; P.mu_off = a->mu_off; P.mu_mode = a->mu_mode;
  P.sigma = a->sigma; P.sg_rs = a->sigma_rs; P.sg_off = a->sigma_off; P.sg_mode = a->sigma_mode;
  P.sg_period = a->sigma_period > 0 ? a->sigma_period : 1;
  P.rows = a->rows; P.cols = a->cols; P.quant = a->quant; P.sigma_is_log = a->sigma_is_log;
  P.lik_bound = a->lik_bound; P.scale_bound = a->scale_bound;
  P.v_hat = a->v_hat; P.vh_rs = a->v_hat_rs; P.vh_off = a->v_hat_off;
  P.v_hat_bf16 = (__nv_bfloat16*)a->v_hat_bf16; P.vb_rs = a->vb_rs; P.vb_off = a->vb_off;
  P.lik = a->lik; P.sum_out = a->sum_ln_out; P.ws = (LikWs*)a->workspace;
  // dense per-element problems collapse to a single row (no per-row index math)
  bool dense = (P.v_rs == P.cols && P.v_off == 0) &&
               (P.mu_mode == 0 || (P.mu_mode == 2 && P.mu_rs == P.cols && P.mu_off == 0)) &&
               (P.sg_mode == 2 && P.sg_rs == P.cols && P.sg_off == 0) &&
               (!P.v_hat || (P.vh_rs == P.cols && P.vh_off == 0)) &&
               (!P.v_hat_bf16 || (P.vb_rs == P.cols && P.vb_off == 0));
  if (dense) {
    P.cols = P.rows * P.cols; P.rows = 1;
    P.v_rs = P.mu_rs = P.sg_rs = P.vh_rs = P.vb_rs = P.cols;
  }
  auto ok4 = [](long long x) { return (x & 3) == 0; };
  bool vec4 = ok4(P.cols) && ok4(P.v_rs) && ok4(P.v_off) && aligned16(P.v) &&
              (P.mu_mode == 0 || (ok4(P.mu_rs) && ok4(P.mu_off) && aligned16(P.mu))) &&
              (P.sg_mode == 3 || (ok4(P.sg_rs) && ok4(P.sg_off) && aligned16(P.sigma))) &&
              (!P.v_hat || (ok4(P.vh_rs) && ok4(P.vh_off) && aligned16(P.v_hat))) &&
              (!P.v_hat_bf16 || (ok4(P.vb_rs) && ok4(P.vb_off) && ((((uintptr_t)P.v_hat_bf16) & 7) == 0))) &&
              (!P.lik || aligned16(P.lik));
  long long items = P.rows * (P.cols / (vec4 ? 4 : 1));
  int grid = (int)((items + kLikThreads - 1) / kLikThreads);
  if (grid > kLikMaxBlocks) grid = kLikMaxBlocks;
  if (grid < 1) grid = 1;
  const bool fast = vec4 && a->form == 0 && P.mu_mode == 2 && P.sg_mode == 2 && a->quant <= 1;
  if (fast) {
    // 5 CTAs of 256 threads per SM measured best on B200 (5.55 TB/s; 8/SM: 4.9, 32/SM: 5.5)
    int gmax = 5 * num_sms();
    if (tuning().lik_grid > 0) { int g = tuning().lik_grid * num_sms(); if (g >= 1 && g <= kLikMaxBlocks) gmax = g; }   // tuning aid
    if (grid > gmax) grid = gmax;
    if (a->quant == 1) {
      if (a->sigma_is_log) k_likelihood_fast<1, true><<<grid, kLikThreads, 0, st>>>(P);
      else k_likelihood_fast<1, false><<<grid, kLikThreads, 0, st>>>(P);
    } else {
      if (a->sigma_is_log) k_likelihood_fast<0, true><<<grid, kLikThreads, 0, st>>>(P);
      else k_likelihood_fast<0, false><<<grid, kLikThreads, 0, st>>>(P);
    }
  } else if (vec4) {
    if (a->form == 0) k_likelihood<4, 0><<<grid, kLikThreads, 0, st>>>(P);
    else k_likelihood<4, 1><<<grid, kLikThreads, 0, st>>>(P);
  } else {
    if (a->form == 0) k_likelihood<1, 0><<<grid, kLikThreads, 0, st>>>(P);
    else k_likelihood<1, 1><<<grid, kLikThreads, 0, st>>>(P);
  }
  return check_launch("k_likelihood");
}

// ------------------------------------------------------------------------------------
// a11: MSE on 8-bit levels (exact integer accumulation)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mse_sum(const float* __restrict__ x, const float* __restrict__ xt,
                                                 long long chw, int clamp_pm1, unsigned long long* __restrict__ out) {
  const int b = blockIdx.y;
  const float* xb = x + (long long)b * chw;
  const float* tb = xt + (long long)b * chw;
  unsigned long long acc = 0;
  long long n4 = ((((uintptr_t)xb | (uintptr_t)tb) & 15) == 0) ? chw / 4 : 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 a = __ldg(reinterpret_cast<const float4*>(xb) + i);
    float4 t = __ldg(reinterpret_cast<const float4*>(tb) + i);
    acc += sq_level_err(a.x, t.x, clamp_pm1) + sq_level_err(a.y, t.y, clamp_pm1) +
           sq_level_err(a.z, t.z, clamp_pm1) + sq_level_err(a.w, t.w, clamp_pm1);
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < chw; i += (long long)gridDim.x * 256)
    acc += sq_level_err(xb[i], tb[i], clamp_pm1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int k = 0; k < 8; ++k) s += ws[k];
    atomicAdd(out + b, s);
  }
}
// ------------------------------------------------------------------------------------
// a9 / a11 scalar tail: the per-batch bpp / PSNR arithmetic of model/net.py:856-869 on the three sum(ln L)
// and the B exact squared-error sums.  One block; fixed summation order (deterministic).
//   packed5 = [sum ln L_z, sum ln L_y, sum ln L_syn, sum_i 20 log10(255 / sqrt(mse_i)), B]   (double: what the
//             multi-GPU all-reduce carries), v_mse[i] = sq_err[i] / chw.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rd_pack(const float* __restrict__ bits3, const unsigned long long* __restrict__ sq_err,
                                                 int B, double chw, double* __restrict__ packed5, float* __restrict__ v_mse) {
  __shared__ double ws[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) {
    const double mse = (double)sq_err[i] / chw;
    if (v_mse) v_mse[i] = (float)mse;
    acc += 20.0 * log10(255.0 / sqrt(mse));
  }
  ws[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < 256; ++k) s += ws[k];
    packed5[0] = (double)bits3[0]; packed5[1] = (double)bits3[1]; packed5[2] = (double)bits3[2];
    packed5[3] = s; packed5[4] = (double)B;
  }
}
// bpp = sum(ln L) / (-ln 2 * n * th * tw), v_psnr = sum_i psnr_i / n   (n images in the possibly all-reduced packed5)
__global__ void k_rd_finish(const double* __restrict__ packed5, double pixels_per_image, float* __restrict__ bpp_psnr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double n = packed5[4];
    bpp_psnr[0] = (float)((packed5[0] + packed5[1] + packed5[2]) / (-0.6931471805599453 * n * pixels_per_image));
    bpp_psnr[1] = (float)(packed5[3] / n);
  }
}
extern "C" int ldic_rd_pack_metrics(const float* bits3, const unsigned long long* sq_err, int B, long long chw, double* packed5,
                                    float* v_mse, void* stream) {
  if (!bits3 || !sq_err || !packed5 || B <= 0 || chw <= 0) return fail(LDIC_EINVAL, "rd_pack_metrics: bad argument");
  k_rd_pack<<<1, 256, 0, (cudaStream_t)stream>>>(bits3, sq_err, B, (double)chw, packed5, v_mse);
  return check_launch("k_rd_pack");
}
extern "C" int ldic_rd_finish_metrics(const double* packed5, double pixels_per_image, float* bpp_psnr, void* stream) {
  if (!packed5 || !bpp_psnr || !(pixels_per_image > 0)) return fail(LDIC_EINVAL, "rd_finish_metrics: bad argument");
  k_rd_finish<<<1, 32, 0, (cudaStream_t)stream>>>(packed5, pixels_per_image, bpp_psnr);
  return check_launch("k_rd_finish");
}

extern "C" int ldic_mse_sum(const float* x, const float* x_tilde, int B, long long chw, int clamp_pm1,
                            unsigned long long* sq_err, void* stream) {
  if (B < 0 || chw < 0) return fail(LDIC_EINVAL, "mse: bad shape");
  if (B == 0 || chw == 0) return LDIC_OK;
  long long blocks = (chw / 4 + 255) / 256;
  int per_img = (int)(blocks < 1 ? 1 : (blocks > num_sms() * 4 ? num_sms() * 4 : blocks));
  k_mse_sum<<<dim3(per_img, B), 256, 0, (cudaStream_t)stream>>>(x, x_tilde, chw, clamp_pm1, sq_err);
  return check_launch("k_mse_sum");
}

// batch_conv (per-image 1x1, M -> 3) + the a11 arithmetic.  One thread per pixel.
template <int M>
__global__ void __launch_bounds__(256) k_syntax_conv_mse(const float* __restrict__ x, const float* __restrict__ xt,
                                                         const float* __restrict__ w, long long HW, int tanh_out,
                                                         float* __restrict__ xo, unsigned long long* __restrict__ out) {
  const int b = blockIdx.y;
  __shared__ float ws[3 * M];
  __shared__ unsigned long long red[8];
  for (int i = threadIdx.x; i < 3 * M; i += 256) ws[i] = w[(long long)b * 3 * M + i];
  __syncthreads();
  unsigned long long acc = 0;
  for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < HW; p += (long long)gridDim.x * 256) {
    const float4* src = reinterpret_cast<const float4*>(xt + ((long long)b * HW + p) * M);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
    for (int q = 0; q < M / 4; ++q) {
      float4 t = __ldg(src + q);
      float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o0 = fmaf(tv[k], ws[q * 4 + k], o0);
        o1 = fmaf(tv[k], ws[M + q * 4 + k], o1);
        o2 = fmaf(tv[k], ws[2 * M + q * 4 + k], o2);
      }
    }
    float o[3] = {o0, o1, o2};
    if (tanh_out) {                          // U-Net family: x~ = tanh(batch_conv(...)) (model/net_unet_ha_hs.py:980)
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = tanhf(o[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      long long idx = ((long long)b * 3 + c) * HW + p;
      if (xo) xo[idx] = o[c];
      acc += sq_level_err(__ldg(x + idx), o[c], 0);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(out + b, s);
  }
}
extern "C" int ldic_syntax_conv_mse(const float* x_nchw, const float* xt_nhwc, const float* w, int B, int M, int H, int W,
                                    int tanh_out, float* x_tilde_nchw, unsigned long long* sq_err, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return (B == 0 || H == 0 || W == 0) ? LDIC_OK : fail(LDIC_EINVAL, "syntax_conv_mse: bad shape");
  long long HW = (long long)H * W;
  long long blocks = (HW + 255) / 256;
  int per_img = (int)(blocks > num_sms() * 4 ? num_sms() * 4 : blocks);
  dim3 grid(per_img, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 16) k_syntax_conv_mse<16><<<grid, 256, 0, st>>>(x_nchw, xt_nhwc, w, HW, tanh_out, x_tilde_nchw, sq_err);
  else if (M == 32) k_syntax_conv_mse<32><<<grid, 256, 0, st>>>(x_nchw, xt_nhwc, w, HW, tanh_out, x_tilde_nchw, sq_err);
  else return fail(LDIC_EINVAL, "syntax_conv_mse: M=%d unsupported (16 or 32)", M);
  return check_launch("k_syntax_conv_mse");
}

// ------------------------------------------------------------------------------------
// layout / dtype glue
// ------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 through a 32-pixel x 32-channel smem tile.
__global__ void __launch_bounds__(256) k_nchw_to_nhwc_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           int C, long long HW, int Cp, int apply_abs) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r;
    long long p = p0 + tx;
    float v = (c < C && p < HW) ? x[((long long)b * C + c) * HW + p] : 0.f;
    t[r][tx] = apply_abs ? fabsf(v) : v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    long long p = p0 + r;
    int c = c0 + tx;
    if (p < HW && c < Cp) y[((long long)b * HW + p) * Cp + c] = __float2bfloat16_rn(t[tx][r]);
  }
}
extern "C" int ldic_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, int Cp, int apply_abs,
                                          void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return LDIC_OK;
  if (Cp < C) return fail(LDIC_EINVAL, "nchw_to_nhwc: Cp < C");
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (Cp + 31) / 32, B);
  k_nchw_to_nhwc_bf16<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, C, HW, Cp, apply_abs);
  return check_launch("k_nchw_to_nhwc_bf16");
}

template <typename TIn>
__global__ void __launch_bounds__(256) k_nhwc_to_nchw(const TIn* __restrict__ x, float* __restrict__ y, int C,
                                                      long long HW, int Cp) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    long long p = p0 + r;
    int c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < Cp) v = (float)x[((long long)b * HW + p) * Cp + c];
    t[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r;
    long long p = p0 + tx;
    if (c < C && p < HW) y[((long long)b * C + c) * HW + p] = t[tx][r];
  }
}
extern "C" int ldic_nhwc_to_nchw_f32(const void* x, int x_is_bf16, float* y, int B, int C, int H, int W, int Cp,
                                     void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return LDIC_OK;
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, B);
  if (x_is_bf16) k_nhwc_to_nchw<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, y, C, HW, Cp);
  else k_nhwc_to_nchw<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, y, C, HW, Cp);
  return check_launch("k_nhwc_to_nchw");
}

__global__ void k_latent_prep(const float* __restrict__ y, size_t n, __nv_bfloat16* __restrict__ yr,
                              __nv_bfloat16* __restrict__ ya, float* __restrict__ yrf) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) {
    float v = y[i];
    float r = rintf(v);
    if (yr) yr[i] = __float2bfloat16_rn(r);
    if (ya) ya[i] = __float2bfloat16_rn(fabsf(v));
    if (yrf) yrf[i] = r;
  }
}
// uint8 levels -> fp32 (u/255)*2-1: torchvision ToTensor followed by eval_net.py:84, the same two correctly rounded
// fp32 operations per sample (division by 255, then *2 and -1).
__global__ void k_u8_to_f32_pm1(const unsigned char* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = __fsub_rn(__fmul_rn(__fdiv_rn((float)x[i], 255.f), 2.f), 1.f);
}
extern "C" int ldic_u8_to_f32_pm1(const unsigned char* x, float* y, size_t n, void* stream) {
  if (n == 0) return LDIC_OK;
  if (!x || !y) return fail(LDIC_EINVAL, "u8_to_f32_pm1: null tensor");
  k_u8_to_f32_pm1<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return check_launch("k_u8_to_f32_pm1");
}

extern "C" int ldic_latent_prep(const float* y, size_t n, void* y_round_bf16, void* y_abs_bf16, float* y_round_f32,
                                void* stream) {
  if (n == 0) return LDIC_OK;
  k_latent_prep<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, n, (__nv_bfloat16*)y_round_bf16,
                                                                (__nv_bfloat16*)y_abs_bf16, y_round_f32);
  return check_launch("k_latent_prep");
}

// First-layer patch matrix.  One CTA = 64 output pixels of one output row: the 5 input rows x Cin
// channels it needs are staged in shared memory with coalesced loads, then every thread assembles
// 16-byte (8 x bf16) chunks of A[p][k], k = (ky*5+kx)*Cin + ci, so the stores are fully coalesced.
constexpr int kI2cPix = 64;
constexpr int kI2cSpan = 2 * kI2cPix + 4;      // input columns 2*ox0-1 .. 2*ox0+2*63+3 (+1 pad)
template <typename TO>
__global__ void __launch_bounds__(256) k_im2col_5x5s2(const float* __restrict__ x, TO* __restrict__ a,
                                                      int Cin, int H, int W, int Ho, int Wo, int Kp) {
  extern __shared__ float sx[];                 // [5*Cin][kI2cSpan]
  const int ox0 = blockIdx.x * kI2cPix, oy = blockIdx.y, b = blockIdx.z;
  const int rows = 5 * Cin;
  for (int i = threadIdx.x; i < rows * kI2cSpan; i += 256) {
    const int r = i / kI2cSpan, c = i - r * kI2cSpan;
    const int ky = r / Cin, ci = r - ky * Cin;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox0 - 1 + c;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)b * Cin + ci) * H + iy) * W + ix);
    sx[i] = v;
  }
  // k -> staged offset (ky*Cin+ci)*span + kx, or -1 for the K padding (no per-element divisions below)
  __shared__ int lut[256];
  for (int k = threadIdx.x; k < Kp && k < 256; k += 256) {
    int off = -1;
    if (k < 25 * Cin) {
      const int tap = k / Cin, ci = k - tap * Cin;
      const int ky = tap / 5, kx = tap - ky * 5;
      off = (ky * Cin + ci) * kI2cSpan + kx;
    }
    lut[k] = off;
  }
  __syncthreads();
  const int chunks = Kp / 8;
  const int cshift = 31 - __clz(chunks);                       // chunks is a power of two (checked on the host)
  const long long p_row = ((long long)b * Ho + oy) * Wo;
  for (int i = threadIdx.x; i < kI2cPix * chunks; i += 256) {
    const int px = i >> cshift, q = i & (chunks - 1);
    if (ox0 + px >= Wo) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int off = lut[q * 8 + e];
      v[e] = off >= 0 ? sx[off + 2 * px] : 0.f;
    }
    if constexpr (std::is_same<TO, float>::value) {             // TF32 parity mode: fp32 patch matrix
      float4* dst = reinterpret_cast<float4*>(a + (p_row + ox0 + px) * Kp + q * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {                            // round to tf32 (the tensor core would truncate)
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v[e]));
        v[e] = __uint_as_float(r);
      }
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(a + (p_row + ox0 + px) * Kp + q * 8) = o;
    }
  }
}
static int im2col_impl(const float* x, void* a, int out_f32, int B, int Cin, int H, int W, int Kp, void* stream);
extern "C" int ldic_im2col_5x5s2(const float* x, void* a, int B, int Cin, int H, int W, int Kp, void* stream) {
  return im2col_impl(x, a, 0, B, Cin, H, W, Kp, stream);
}
extern "C" int ldic_im2col_5x5s2_f32(const float* x, float* a, int B, int Cin, int H, int W, int Kp, void* stream) {
  return im2col_impl(x, a, 1, B, Cin, H, W, Kp, stream);
}
static int im2col_impl(const float* x, void* a, int out_f32, int B, int Cin, int H, int W, int Kp, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return LDIC_OK;
  if ((H & 1) || (W & 1) || Kp % 8 || Kp < 25 * Cin) return fail(LDIC_EINVAL, "im2col: H,W must be even and Kp>=25*Cin, Kp%%8==0");
  if (Cin > 8 || B > 65535 || H / 2 > 65535) return fail(LDIC_EINVAL, "im2col: Cin <= 8 and B, H/2 <= 65535");
  if (Kp > 256 || ((Kp / 8) & (Kp / 8 - 1))) return fail(LDIC_EINVAL, "im2col: Kp must be 64, 128 or 256");
  int Ho = H / 2, Wo = W / 2;
  dim3 grid((Wo + kI2cPix - 1) / kI2cPix, Ho, B);
  size_t smem = (size_t)5 * Cin * kI2cSpan * sizeof(float);
  if (out_f32) k_im2col_5x5s2<float><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (float*)a, Cin, H, W, Ho, Wo, Kp);
  else k_im2col_5x5s2<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)a, Cin, H, W, Ho, Wo, Kp);
  return check_launch("k_im2col_5x5s2");
}

// Context-model input image: x[p][0:N] = round(y) (bf16, already rounded), x[p][N:2N] = bf16(h2).
// (model/net.py:307-309 concatenates the sampled y and h patches; here the concat happens once on
// the latent images and the patches are gathered by TMA inside the first context conv.)
__global__ void k_ctx_pack_input(const uint4* __restrict__ yr, const float4* __restrict__ h2, uint4* __restrict__ x,
                                 long long P, int N8) {
  const long long total = P * N8 * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / (2 * N8);
    const int c8 = (int)(i - p * 2 * N8);
    uint4 o;
    if (c8 < N8) {
      o = __ldg(yr + p * N8 + c8);
    } else {
      const float4 a = __ldg(h2 + (p * N8 + (c8 - N8)) * 2), b = __ldg(h2 + (p * N8 + (c8 - N8)) * 2 + 1);
      o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    x[i] = o;
  }
}
extern "C" int ldic_ctx_pack_input(const void* y_round_bf16, const float* h2, void* x, long long P, int N, void* stream) {
  if (P <= 0) return LDIC_OK;
  if (N <= 0 || N % 8) return fail(LDIC_EINVAL, "ctx_pack_input: N must be a multiple of 8");
  long long total = P * (N / 8) * 2;
  int grid = (int)((total + 255) / 256 > num_sms() * 16 ? num_sms() * 16 : (total + 255) / 256);
  k_ctx_pack_input<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)y_round_bf16, (const float4*)h2, (uint4*)x, P, N / 8);
  return check_launch("k_ctx_pack_input");
}

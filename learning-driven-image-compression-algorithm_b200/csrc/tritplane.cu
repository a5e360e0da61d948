// Progressive trit-plane quantisation + per-plane likelihood (BASELINE configs[4]).
//
// NOT a port: the reference file model/Trit_Plane.py:25-57 is a stand-alone script that crashes (cv2 on an
// absolute path) and contains neither trit planes nor a likelihood, so there is no reference behaviour to be
// identical to ("parity unpinned").  This is the builder-defined extension SURVEY 8 a12 describes, built on the
// reference's own quantiser and Gaussian mass (a6/a7):
//   q   = round(v - mu)   (half-to-even, the CompressAI "dequantize" symbol, model/net_unet_ha_hs.py:937),
//         clamped to [-H, H], H = (3^L - 1) / 2
//   u   = q + H  = sum_l t_l 3^l,  t_l in {0,1,2}          (offset ternary, plane L-1 = most significant)
//   L_l = P(t_l | t_{l+1..L-1}) = mass of N(mu, sigma) over the third of the current interval selected by t_l
//         divided by the mass over the whole current interval (DPICT-style interval thirds); the interval of plane
//         L-1 is [-H - 1/2, H + 1/2] around mu, so prod_l L_l = P(q) / P(|q| <= H).
// Outputs: the L trit planes (int8, plane-major), the symbols q (int32) and sum(ln L_l) per plane.
// HBM bound: 12 B read + (L + 4) B written per element.  Bit-exact property: sum_l t_l 3^l - H == q == round(v - mu).
#include "common.cuh"

using namespace ldic;

namespace {

constexpr int kTpThreads = 256;
constexpr int kTpMaxPlanes = 8;

struct TpWs { double partial[kTpMaxPlanes][kMaxSMs * 8]; unsigned int ticket; };

// erfc with a fractional error below 1.2e-7 everywhere (Chebyshev fit of Numerical Recipes' erfcc): one MUFU.RCP, one
// MUFU.EX2 and ten FMAs instead of the ~45 instructions of erfcf -- the kernel evaluates ten of them per element and
// is ALU bound, not HBM bound.
__device__ __forceinline__ float erfc_fit(float x) {
  const float z = fabsf(x);
  const float t = __frcp_rn(fmaf(0.5f, z, 1.0f));
  float p = fmaf(t, 0.17087277f, -0.82215223f);
  p = fmaf(t, p, 1.48851587f);
  p = fmaf(t, p, -1.13520398f);
  p = fmaf(t, p, 0.27886807f);
  p = fmaf(t, p, -0.18628806f);
  p = fmaf(t, p, 0.09678418f);
  p = fmaf(t, p, 0.37409196f);
  p = fmaf(t, p, 1.00002368f);
  p = fmaf(t, p, -1.26551223f);
  const float r = t * __expf(fmaf(-z, z, p));
  return x >= 0.f ? r : 2.0f - r;
}

// mass of N(0, s) over [a, b] (a <= b), erfc form (keeps the tails), floor 1e-30 so that ratios stay finite
__device__ __forceinline__ float gauss_mass(float a, float b, float inv_s_sqrt2) {
  if (a + b < 0.f) { const float t = a; a = -b; b = -t; }      // mirror to the upper tail: erfc stays small and precise
  const float m = 0.5f * (erfc_fit(a * inv_s_sqrt2) - erfc_fit(b * inv_s_sqrt2));
  return fmaxf(m, 1e-30f);
}

__global__ void __launch_bounds__(kTpThreads) k_tritplane(const float* __restrict__ v, const float* __restrict__ mu,
                                                          const float* __restrict__ sigma, long long n, int L,
                                                          float scale_bound, float lik_bound, signed char* __restrict__ planes,
                                                          int* __restrict__ q_out, float* __restrict__ sum_ln, TpWs* ws) {
  __shared__ double red[kTpMaxPlanes][kTpThreads / 32];
  int pow3[kTpMaxPlanes + 1];
  pow3[0] = 1;
#pragma unroll
  for (int l = 0; l < kTpMaxPlanes; ++l) pow3[l + 1] = pow3[l] * 3;
  const int H = (pow3[L] - 1) / 2;
  float acc[kTpMaxPlanes];
#pragma unroll
  for (int l = 0; l < kTpMaxPlanes; ++l) acc[l] = 0.f;
  for (long long i = blockIdx.x * (long long)kTpThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kTpThreads) {
    const float m = mu ? __ldg(mu + i) : 0.f;
    const float s = fmaxf(__ldg(sigma + i), scale_bound);
    float qf = rintf(__fsub_rn(__ldg(v + i), m));
    qf = fminf(fmaxf(qf, (float)-H), (float)H);
    const int q = (int)qf;
    if (q_out) q_out[i] = q;
    int u = q + H;
    const float inv = 1.0f / (s * 1.41421356237309515f);
    // current interval [lo, lo + width) in units of u, as offsets from mu: u -> u - H.  The child interval of plane l IS
    // the parent interval of plane l-1, so its mass (and logarithm) is computed once and carried down: L + 1 interval
    // masses per element instead of 2 L, and ln(child / parent) = ln child - ln parent without a division.
    int lo = 0;
    const float a_top = (float)(-H) - 0.5f;
    float ln_parent = __logf(gauss_mass(a_top, a_top + (float)pow3[L], inv));
    const float ln_bound = __logf(lik_bound);
#pragma unroll
    for (int l = kTpMaxPlanes - 1; l >= 0; --l) {
      if (l < L) {
        const int w = pow3[l];                       // width of one third
        const int t = (u / w) % 3;
        if (planes) planes[(long long)l * n + i] = (signed char)t;
        const float a = (float)(lo - H) - 0.5f;
        const float ln_child = __logf(gauss_mass(a + (float)(t * w), a + (float)((t + 1) * w), inv));
        acc[l] += fmaxf(ln_child - ln_parent, ln_bound);
        ln_parent = ln_child;
        lo += t * w;
      }
    }
  }
  // deterministic reduction: warp shuffle -> smem -> one double per CTA and plane -> last CTA adds in order
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int l = 0; l < kTpMaxPlanes; ++l) {
    const float w = warp_sum(acc[l]);
    if (lane == 0) red[l][wid] = (double)w;
  }
  __syncthreads();
  if (threadIdx.x < kTpMaxPlanes) {
    double s = 0.0;
    for (int k = 0; k < kTpThreads / 32; ++k) s += red[threadIdx.x][k];
    ws->partial[threadIdx.x][blockIdx.x] = s;
  }
  __threadfence();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < kTpMaxPlanes) {
      double s = 0.0;
      for (unsigned k = 0; k < gridDim.x; ++k) s += ws->partial[threadIdx.x][k];
      if (threadIdx.x < L) sum_ln[threadIdx.x] = (float)s;
    }
    if (threadIdx.x == 0) ws->ticket = 0;          // leave the workspace reusable
  }
}

}  // namespace

extern "C" size_t ldic_tritplane_workspace_bytes(void) { return sizeof(TpWs); }

extern "C" int ldic_tritplane_likelihood(const float* v, const float* mu, const float* sigma, long long n, int L,
                                         float scale_bound, float lik_bound, signed char* planes, int* q_out,
                                         float* sum_ln_per_plane, void* workspace, void* stream) {
  if (n < 0 || L < 1 || L > kTpMaxPlanes) return fail(LDIC_EINVAL, "tritplane: 1 <= L <= %d planes", kTpMaxPlanes);
  if (!v || !sigma || !sum_ln_per_plane || !workspace) return fail(LDIC_EINVAL, "tritplane: null argument");
  if (n == 0) {
    LDIC_CUDA(cudaMemsetAsync(sum_ln_per_plane, 0, sizeof(float) * L, (cudaStream_t)stream));
    return LDIC_OK;
  }
  long long blocks = (n + kTpThreads - 1) / kTpThreads;
  const int grid = (int)(blocks > num_sms() * 8 ? num_sms() * 8 : blocks);
  k_tritplane<<<grid, kTpThreads, 0, (cudaStream_t)stream>>>(v, mu, sigma, n, L, scale_bound, lik_bound, planes, q_out,
                                                             sum_ln_per_plane, (TpWs*)workspace);
  return check_launch("k_tritplane");
}

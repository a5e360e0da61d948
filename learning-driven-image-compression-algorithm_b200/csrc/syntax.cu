// The "syntax" side branch of Net.forward (model/net.py:349-413, 322-343, call sites :712-719, :753,
// :789, :805) as five small fp32 CUDA-core launches instead of ~40 cuDNN / ATen launches:
//   Syntax_Model            pool(y[:, :M]) | relu(down0) -> pool | relu(down1) -> pool -> 1x1 conv -> z3_syntax
//   PredictionModel_Syntax  relu(down0(h2)) -> relu(down1) ; pools of h2, ds0, ds1 -> fc -> (mu, exp(.))
//   conv_generator          Linear(M,128) LeakyReLU Linear(128,256) LeakyReLU Linear(256,3M) on round(z3_syntax)
// All tensors are NHWC fp32 (the latent y and the h_s output h2 are produced in that layout by the
// transforms).  < 0.1 % of the forward's FLOPs: the point is launch count and the slow generic
// cuDNN engines these tiny channels-last convolutions hit, not arithmetic throughput.
#include "common.cuh"

using namespace ldic;

namespace {

constexpr int kScPix = 32;       // output pixels per CTA = lanes of a warp
constexpr int kScGroups = 8;     // warps per CTA; warp g owns every 8th (tap, channel chunk) pair
constexpr int kScThreads = kScPix * kScGroups;
constexpr int kScChunk = 32;     // input channels per (tap, chunk) pair

// Re-pack of the four 3x3 weight tensors into wp[tap][ci][Cout] (one launch).
struct PackDesc { const float* w; float* wp; int Cin, Cout, total; };
struct PackArgs { PackDesc d[4]; };
__global__ void k_pack_small(PackArgs a) {
  const PackDesc d = a.d[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.total; i += gridDim.x * blockDim.x) {
    const int co = i % d.Cout, r = i / d.Cout, ci = r % d.Cin, tap = r / d.Cin;
    d.wp[i] = __ldg(d.w + ((long long)co * d.Cin + ci) * 9 + tap);
  }
}

// y[b][oy][ox][co] = relu(bias[co] + sum_{ky,kx,ci} w[co][ci][ky][kx] * x[b][2oy+ky-1][2ox+kx-1][ci])
// (Conv2d(k3, s2, p1), zero padding).  x pixels are `xs` floats apart (a channel slice of a wider tensor).
// Lane = output pixel; the (tap, 32-channel chunk) pairs of the reduction are dealt round-robin to the 8
// warps of the CTA, so there is no barrier inside the reduction: every lane reads its own pixel's 32 channels
// (one 128-byte line) with 256-bit loads -- whole 32-byte sectors per request; with 128-bit loads every sector
// crossed the L2 -> L1 path twice (124 MB for the 19 MB input of PredictionModel_Syntax.down0).  The packed weights
// of the whole layer (9 x Cin x COUT floats, 9 .. 110 KB) are brought into shared memory once per CTA with cp.async
// and read from there as broadcasts: fetched from global memory inside the reduction every weight row was a first
// touch on its SM and the kernels ran at L2 latency (ncu: 12-18 long-scoreboard stall cycles per issued instruction,
// 48 us for the 56 MFLOP of Syntax_Model.down1).  The eight partial sums are added in a fixed order at the end.
__device__ __forceinline__ void ldg_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}

template <int COUT>
__global__ void __launch_bounds__(kScThreads) k_small_conv3x3s2_relu(const float* __restrict__ x, long long xs,
                                                                     const float* __restrict__ wp,
                                                                     const float* __restrict__ bias, float* __restrict__ y,
                                                                     int B, int H, int W, int Ho, int Wo, int Cin,
                                                                     int cslice) {
  extern __shared__ __align__(16) float s_w[];                         // [9][cslice][COUT]: one channel slice of the weights
  __shared__ float s_red[kScGroups / 2][kScPix][COUT + 1];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  // the layer's weights are staged one slice of `cslice` input channels at a time (all of them when they fit: N = 192;
  // PredictionModel_Syntax.down0 of the N = 384 model needs three slices of 128)
  auto stage_weights = [&](int s0, int cs) {
    const int per_tap = cs * COUT / 4;                                 // 16-byte chunks per tap
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_w);
    for (int i = threadIdx.x; i < 9 * per_tap; i += kScThreads) {
      const int tap = i / per_tap, r = i - tap * per_tap;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u),
                   "l"(wp + ((size_t)tap * Cin + s0) * COUT + (size_t)r * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_weights(0, Cin < cslice ? Cin : cslice);
  const long long npix = (long long)B * Ho * Wo;
  const long long p = (long long)blockIdx.x * kScPix + lane;
  const bool live = p < npix;
  int ox = 0, oy = 0, b = 0;
  if (live) { ox = (int)(p % Wo); long long r = p / Wo; oy = (int)(r % Ho); b = (int)(r / Ho); }
  float acc[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
  const int ck = Cin < kScChunk ? Cin : kScChunk;          // channels per chunk (16 or 32)
#pragma unroll 1
  for (int s0 = 0; s0 < Cin; s0 += cslice) {
  const int cs = Cin - s0 < cslice ? Cin - s0 : cslice;
  if (s0 > 0) { __syncthreads(); stage_weights(s0, cs); }   // every warp is done with the previous slice
  const int nchunk = cs / ck, npairs = 9 * nchunk;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
#pragma unroll 1
  for (int q = g; q < npairs; q += kScGroups) {
    const int tap = q / nchunk, c0 = (q - tap * nchunk) * ck;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int iy = 2 * oy + ky - 1, ix = 2 * ox + kx - 1;
    const bool in = live && iy >= 0 && iy < H && ix >= 0 && ix < W;
    const float* xp = x + (((long long)b * H + iy) * W + ix) * xs + s0 + c0;
    float xv[kScChunk / 8][8];
#pragma unroll
    for (int j8 = 0; j8 < kScChunk / 8; ++j8) {
      if (in && 8 * j8 < ck) {
        ldg_v8(xp + 8 * j8, xv[j8]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[j8][e] = 0.f;
      }
    }
    const float4* wr = reinterpret_cast<const float4*>(s_w + ((size_t)tap * cs + c0) * COUT);
#pragma unroll
    for (int j = 0; j < kScChunk; ++j) {
      if (j < ck) {                                          // warp-uniform
        const float4* wj = wr + j * (COUT / 4);
        const float xu = xv[j >> 3][j & 7];
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 w4 = wj[c4];
          acc[4 * c4] = fmaf(xu, w4.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(xu, w4.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(xu, w4.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(xu, w4.w, acc[4 * c4 + 3]);
        }
      }
    }
  }
  }
  // fixed-order tree over the 8 warps: (g) += (g + 4), then warp 0 adds warps 1, 2, 3
  if (g >= kScGroups / 2) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) s_red[g - kScGroups / 2][lane][c] = acc[c];
  }
  __syncthreads();
  if (g < kScGroups / 2) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] += s_red[g][lane][c];
  }
  __syncthreads();
  if (g > 0 && g < kScGroups / 2) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) s_red[g][lane][c] = acc[c];
  }
  __syncthreads();
  if (g == 0 && live) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      float v = acc[c];
#pragma unroll
      for (int gg = 1; gg < kScGroups / 2; ++gg) v += s_red[gg][lane][c];
      acc[c] = fmaxf(v + __ldg(bias + c), 0.f);
    }
    float4* dst = reinterpret_cast<float4*>(y + p * COUT);
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) dst[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
  }
}

template <int COUT>
int launch_small_conv(const float* x, long long xs, const float* wp, const float* bias, float* y, int B, int H, int W,
                      int Cin, cudaStream_t st) {
  if (Cin % 16 || xs % 8 || (((uintptr_t)x) & 31) || (((uintptr_t)wp) & 15) || (((uintptr_t)y) & 15))
    return fail(LDIC_EINVAL, "syntax branch: small conv needs Cin %% 16 == 0, pixel stride %% 8 == 0, a 32-byte aligned input and 16-byte aligned weights / output");
  if (Cin > kScChunk && Cin % kScChunk) return fail(LDIC_EINVAL, "syntax branch: small conv needs Cin <= 32 or a multiple of 32");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long npix = (long long)B * Ho * Wo;
  const unsigned grid = (unsigned)((npix + kScPix - 1) / kScPix);
  const int ck = Cin < kScChunk ? Cin : kScChunk;
  int cslice = (int)((160 * 1024) / (9 * COUT * sizeof(float))) / ck * ck;     // input channels whose weights fit at once
  if (cslice > Cin) cslice = Cin;
  if (cslice < ck) return fail(LDIC_EINVAL, "syntax branch: 3x3 weights of %d x %d channels do not fit in shared memory", Cin, COUT);
  const size_t smem = (size_t)9 * cslice * COUT * sizeof(float);
  {
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return fail(LDIC_ECUDA, "syntax branch: bad current device");
    std::lock_guard<std::mutex> init_lock(g_init_mu);
    static bool attr_set[kMaxDevices] = {};
    if (!attr_set[dev]) {
      LDIC_CUDA(cudaFuncSetAttribute(k_small_conv3x3s2_relu<COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set[dev] = true;
    }
  }
  k_small_conv3x3s2_relu<COUT><<<grid, kScThreads, smem, st>>>(x, xs, wp, bias, y, B, H, W, Ho, Wo, Cin, cslice);
  return check_launch("k_small_conv3x3s2_relu");
}

int small_conv(const float* x, long long xs, const float* wp, const float* bias, float* y, int B, int H, int W, int Cin,
               int Cout, cudaStream_t st) {
  switch (Cout) {
    case 16: return launch_small_conv<16>(x, xs, wp, bias, y, B, H, W, Cin, st);
    case 32: return launch_small_conv<32>(x, xs, wp, bias, y, B, H, W, Cin, st);
    case 64: return launch_small_conv<64>(x, xs, wp, bias, y, B, H, W, Cin, st);
  }
  return fail(LDIC_EINVAL, "syntax branch: small conv supports 16, 32 or 64 output channels, got %d", Cout);
}

// Partial sums for the six mean pools (AdaptiveAvgPool2d(1), model/net.py:362-367,401-404): CTA (chunk, b, t)
// adds kPoolChunk pixels of all channels of tensor t; the head adds the chunks in order (deterministic).
constexpr int kPoolChunk = 32;
constexpr int kPoolTensors = 6;
struct PoolDesc { const float* x; long long xs; int C, npix, nchunk; long long part_off; };
struct PoolArgs { PoolDesc d[kPoolTensors]; float* part; };

__global__ void __launch_bounds__(256) k_pool_partial(PoolArgs a) {
  __shared__ float s[256];
  const PoolDesc d = a.d[blockIdx.z];
  const int b = blockIdx.y, chunk = blockIdx.x;
  if (chunk >= d.nchunk) return;
  const int p0 = chunk * kPoolChunk, p1 = min(p0 + kPoolChunk, d.npix);
  for (int c0 = 0; c0 < d.C; c0 += 256) {               // channel tiles of 256 (C <= 256: one pass, 256/C pixel groups)
    const int Ct = min(d.C - c0, 256);
    const int Gt = 256 / Ct;
    const int t = threadIdx.x;
    float v = 0.f;
    const int c = t % Ct, g = t / Ct;
    if (g < Gt)
      for (int p = p0 + g; p < p1; p += Gt) v += __ldg(d.x + ((long long)b * d.npix + p) * d.xs + c0 + c);
    s[t] = v;
    __syncthreads();
    if (t < Ct) {
      float acc = 0.f;
      for (int gg = 0; gg < Gt; ++gg) acc += s[gg * Ct + t];
      a.part[d.part_off + ((long long)b * d.nchunk + chunk) * d.C + c0 + t] = acc;
    }
    __syncthreads();
  }
}

inline int pool_chunks(int npix) { return (npix + kPoolChunk - 1) / kPoolChunk; }

// ---- head: one CTA per image ------------------------------------------------------------------
constexpr int kHeadThreads = 512;

// mean pool from the per-chunk partial sums: out[c] = (sum_k part[b][k][c]) / npix, chunk order fixed
__device__ void pool_finish(const float* __restrict__ part, long long off, int b, int C, int npix, float* out) {
  const int nchunk = (npix + kPoolChunk - 1) / kPoolChunk;
  for (int c = threadIdx.x; c < C; c += kHeadThreads) {
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += __ldg(part + off + ((long long)b * nchunk + k) * C + c);
    out[c] = s / (float)npix;
  }
}

// out[j] = bias[j] + sum_i w[j][i] * in[i]   (nn.Linear / 1x1 conv on a pooled vector): a warp computes four
// output rows at a time (coalesced weight reads, independent load chains, fixed-order butterfly)
__device__ void dense(const float* __restrict__ w, const float* __restrict__ bias, const float* in, int n_in, int n_out,
                      float* out, int act /*0 none, 1 leaky 0.2*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j0 = warp * 4; j0 < n_out; j0 += (kHeadThreads / 32) * 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < n_in; i += 32) {
      const float v = in[i];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < n_out) s[u] = fmaf(__ldg(w + (long long)(j0 + u) * n_in + i), v, s[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(s[u]);
      if (lane == 0 && j0 + u < n_out) {
        float r = t + __ldg(bias + j0 + u);
        if (act == 1) r = r > 0.f ? r : 0.2f * r;
        out[j0 + u] = r;
      }
    }
  }
  __syncthreads();
}

struct HeadOffsets { long long off[kPoolTensors]; };

__global__ void __launch_bounds__(kHeadThreads) k_syntax_head(LdicSyntaxArgs a, HeadOffsets po) {
  __shared__ float cat_s[512];          // pooled vectors (M + 32 + 64  and  N + M + M)
  __shared__ float v0[256], v1[256];
  const int b = blockIdx.x, M = a.M, N = a.N;
  const int h = a.h, w = a.w;
  const int h1 = (h - 1) / 2 + 1, w1 = (w - 1) / 2 + 1, h2s = (h1 - 1) / 2 + 1, w2s = (w1 - 1) / 2 + 1;
  // ---- Syntax_Model (model/net.py:361-375) ----
  pool_finish(a.pool_part, po.off[0], b, M, h * w, cat_s);
  pool_finish(a.pool_part, po.off[1], b, 32, h1 * w1, cat_s + M);
  pool_finish(a.pool_part, po.off[2], b, 64, h2s * w2s, cat_s + M + 32);
  __syncthreads();
  dense(a.sm_conv_w, a.sm_conv_b, cat_s, M + 96, M, v0, 0);
  if (threadIdx.x < M) {
    const float z = v0[threadIdx.x];
    a.z3[(long long)b * M + threadIdx.x] = z;
    const float zr = rintf(z);                                      // torch.round, model/net.py:753
    a.z3_round[(long long)b * M + threadIdx.x] = zr;
    v1[threadIdx.x] = a.z3_round_in ? a.z3_round_in[(long long)b * M + threadIdx.x] : zr;
  }
  __syncthreads();
  // ---- PredictionModel_Syntax (model/net.py:393-413): pooled h2 | ds0 | ds1 -> fc ----
  pool_finish(a.pool_part, po.off[3], b, N, h * w, cat_s);
  pool_finish(a.pool_part, po.off[4], b, M, h1 * w1, cat_s + N);
  pool_finish(a.pool_part, po.off[5], b, M, h2s * w2s, cat_s + N + M);
  // ---- conv_generator (model/net.py:331-343) on the rounded syntax vector ----
  dense(a.cg_w0, a.cg_b0, v1, M, 128, v0, 1);
  dense(a.cg_w1, a.cg_b1, v0, 128, 256, v1, 1);
  dense(a.cg_w2, a.cg_b2, v1, 256, 3 * M, v0, 0);
  if (threadIdx.x < 3 * M) a.conv_w[(long long)b * 3 * M + threadIdx.x] = v0[threadIdx.x];
  __syncthreads();
  dense(a.ps_fc_w, a.ps_fc_b, cat_s, N + 2 * M, 2 * M, v0, 0);
  if (threadIdx.x < M) {
    a.mu[(long long)b * M + threadIdx.x] = v0[threadIdx.x];
    a.sigma[(long long)b * M + threadIdx.x] = expf(v0[M + threadIdx.x]);
  }
}

struct SyntaxDims { int h1, w1, h2, w2; long long off[kPoolTensors + 1]; long long wp_off[5]; };
SyntaxDims syntax_dims(int B, int h, int w, int N, int M) {
  SyntaxDims d;
  d.h1 = (h - 1) / 2 + 1; d.w1 = (w - 1) / 2 + 1; d.h2 = (d.h1 - 1) / 2 + 1; d.w2 = (d.w1 - 1) / 2 + 1;
  const int C[kPoolTensors] = {M, 32, 64, N, M, M};
  const int np[kPoolTensors] = {h * w, d.h1 * d.w1, d.h2 * d.w2, h * w, d.h1 * d.w1, d.h2 * d.w2};
  d.off[0] = 0;
  for (int t = 0; t < kPoolTensors; ++t) d.off[t + 1] = d.off[t] + (long long)B * pool_chunks(np[t]) * C[t];
  // packed 3x3 weights of the four small convs follow the pool partials (16-byte aligned offsets)
  const long long wsz[4] = {9LL * M * 32, 9LL * 32 * 64, 9LL * N * M, 9LL * M * M};
  d.wp_off[0] = (d.off[kPoolTensors] + 3) / 4 * 4;
  for (int i = 0; i < 4; ++i) d.wp_off[i + 1] = d.wp_off[i] + (wsz[i] + 3) / 4 * 4;
  return d;
}

}  // namespace

extern "C" long long ldic_syntax_workspace_elems(int B, int h, int w, int N, int M) {
  return syntax_dims(B, h, w, N, M).wp_off[4];
}

extern "C" int ldic_syntax_branch(const LdicSyntaxArgs* a, void* stream) {
  if (!a) return fail(LDIC_EINVAL, "syntax branch: null args");
  if (a->B <= 0) return LDIC_OK;
  const int M = a->M, N = a->N, h = a->h, w = a->w, B = a->B;
  if ((M != 16 && M != 32) || N < M || N > 384 || h <= 0 || w <= 0)
    return fail(LDIC_EINVAL, "syntax branch: M must be 16 or 32 and N <= 384 (got M=%d N=%d)", M, N);
  if (N + 2 * M > 512 || M + 96 > 512) return fail(LDIC_EINVAL, "syntax branch: pooled vector too long");
  if (B > 65535) return fail(LDIC_EINVAL, "syntax branch: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  const SyntaxDims d = syntax_dims(B, h, w, N, M);
  float* ws = a->pool_part;
  int rc;
  {
    PackArgs pk;
    const float* wsrc[4] = {a->sm_down0_w, a->sm_down1_w, a->ps_down0_w, a->ps_down1_w};
    const int cin[4] = {M, 32, N, M}, cout[4] = {32, 64, M, M};
    int maxtotal = 0;
    for (int i = 0; i < 4; ++i) {
      pk.d[i].w = wsrc[i]; pk.d[i].wp = ws + d.wp_off[i]; pk.d[i].Cin = cin[i]; pk.d[i].Cout = cout[i];
      pk.d[i].total = 9 * cin[i] * cout[i];
      if (pk.d[i].total > maxtotal) maxtotal = pk.d[i].total;
    }
    k_pack_small<<<dim3((maxtotal + 255) / 256, 4), 256, 0, st>>>(pk);
    if ((rc = check_launch("k_pack_small"))) return rc;
  }
  // Syntax_Model.down0 / down1 on the first M channels of y (model/net.py:364,366)
  if ((rc = small_conv(a->y, N, ws + d.wp_off[0], a->sm_down0_b, a->sm_ds1, B, h, w, M, 32, st))) return rc;
  if ((rc = small_conv(a->sm_ds1, 32, ws + d.wp_off[1], a->sm_down1_b, a->sm_ds2, B, d.h1, d.w1, 32, 64, st))) return rc;
  // PredictionModel_Syntax.down0 / down1 on the h_s output (model/net.py:399-400)
  if ((rc = small_conv(a->h2, N, ws + d.wp_off[2], a->ps_down0_b, a->ps_ds0, B, h, w, N, M, st))) return rc;
  if ((rc = small_conv(a->ps_ds0, M, ws + d.wp_off[3], a->ps_down1_b, a->ps_ds1, B, d.h1, d.w1, M, M, st))) return rc;
  PoolArgs pa;
  const float* xs_[kPoolTensors] = {a->y, a->sm_ds1, a->sm_ds2, a->h2, a->ps_ds0, a->ps_ds1};
  const long long st_[kPoolTensors] = {N, 32, 64, N, M, M};
  const int C_[kPoolTensors] = {M, 32, 64, N, M, M};
  const int np_[kPoolTensors] = {h * w, d.h1 * d.w1, d.h2 * d.w2, h * w, d.h1 * d.w1, d.h2 * d.w2};
  int maxchunk = 0;
  HeadOffsets ho;
  for (int t = 0; t < kPoolTensors; ++t) {
    pa.d[t].x = xs_[t]; pa.d[t].xs = st_[t]; pa.d[t].C = C_[t]; pa.d[t].npix = np_[t]; pa.d[t].nchunk = pool_chunks(np_[t]);
    pa.d[t].part_off = d.off[t]; ho.off[t] = d.off[t];
    if (pa.d[t].nchunk > maxchunk) maxchunk = pa.d[t].nchunk;
  }
  pa.part = ws;
  k_pool_partial<<<dim3(maxchunk, B, kPoolTensors), 256, 0, st>>>(pa);
  if ((rc = check_launch("k_pool_partial"))) return rc;
  k_syntax_head<<<B, kHeadThreads, 0, st>>>(*a, ho);
  return check_launch("k_syntax_head");
}

// rANS entropy coder over the quantised latents and their Gaussian parameters.
//
// NOT a port: the reference has no entropy coder or bitstream anywhere (it only *estimates* the rate,
// model/net.py:856-861), so there is no reference behaviour to be identical to ("parity unpinned" against the
// reference).  What pins this file instead: (1) decode(encode(symbols)) == symbols bit for bit, for every input incl.
// symbols far outside the model's window, (2) the byte stream equals the one the CPU restatement oracle/rans_ref.py
// writes, bit for bit (all model arithmetic below is IEEE fp32 add / sub / mul / div and integers, nothing the two
// sides could round differently), (3) 8 * bytes stays within a fraction of a percent + a fixed header of the rate
// sum(-log2 L) the a7 / a8 likelihood kernels estimate for the same symbols.
//
// Symbols: quant 1: k = round(v) coded with N(mu, sigma)           (GaussianModel, model/net.py:272-286, :741)
//          quant 2: k = round(v - mu) coded with N(0, sigma), v^ = k + mu   (GaussianConditional "dequantize",
//                   model/net_unet_ha_hs.py:937)
// Integer model of one symbol (16-bit probabilities): window of Nsym = 2R+1 integers around m = rint(mu),
//   R = min(1023, max(15, 2 + ceil(6 sigma)));   C(j) = ((Phi24(t_j) * (65536 - Nsym)) >> 24) + j,  t_j = (k_j - 1/2 - mu) * (1 / sigma)  [fp32, each op rounded],
//   C(0) = 0, C(Nsym) = 65536; Phi24 = 24-bit table of the normal CDF (step 1/128, linear interpolation in integers).
// Every symbol of the window has frequency >= 1; the two edge symbols double as escape markers ("at or beyond the
// edge") and the escaped values travel out of band as (index, value) pairs.
// Coder: 32-bit rANS states in [2^16, 2^32), 16-bit words, `streams` independent states per segment (stream s codes the
// run of symbols [s Ls, (s+1) Ls), Ls = ceil(n / S)), one independent bitstream per segment (image).  Encoding: one
// fully parallel pass turns (v, mu, sigma) into (start, freq, floor(2^32 / freq)) triples, one thread per stream runs
// the sequential state recurrence backwards over its run (the division is a multiply-high + one correction), a scan +
// pack pass concatenates the streams.  Decoding: one WARP per stream -- the lanes fetch the parameters of 32 symbols at
// a time and, per symbol, each lane owns one sub-interval of the window (one integer for windows of up to 32, i.e.
// sigma <= 2.2; wider windows continue with a 32-ary search inside the winning sub-interval, at most two more rounds).
// Bitstream (little endian), per segment:
//   u32 magic 'LRA1' | u32 n | u32 S | u32 E | u32 W | u32 quant | u32 column groups (0: none) | u32 0 | u32 state[S] | u16 words_of_stream[S] (+pad to 4)
//   | {u32 index, i32 value} escape[E] | u16 word[W]   (stream 0's words first, each in decoding order)
#include "common.cuh"
#include "rans_phi_table.h"

using namespace ldic;

namespace {

constexpr uint32_t kMagic = 0x3141524Cu;       // 'LRA1'
constexpr int kHeaderBytes = 32;
constexpr uint32_t kRansL = 1u << 16;          // lower bound of the state interval
constexpr int kOpsThreads = 256;
constexpr int kOpsPerThread = 4;
constexpr int kOpsChunk = kOpsThreads * kOpsPerThread;   // symbols per CTA of the parallel passes
constexpr int kStreamThreads = 32;

__device__ const uint32_t c_phi[LDIC_RANS_PHI_N + 1] = {LDIC_RANS_PHI_TABLE};
const uint32_t h_phi[LDIC_RANS_PHI_N + 1] = {LDIC_RANS_PHI_TABLE};

// status bits (per segment)
enum { ST_SYMBOL_RANGE = 1, ST_CAPACITY = 2, ST_HEADER = 4, ST_CORRUPT = 8, ST_ORDER = 16 };

struct Addr {
  const float* v; long long v_rs, v_off;
  const float* mu; long long mu_rs, mu_off; int mu_mode;
  const float* sigma; long long sg_rs, sg_off; int sg_mode; int sg_period;
  unsigned cols; long long rows_per_seg;
  int quant, sigma_is_log; float scale_bound;
  const int* prow;      // optional: per-element (mode 2) mu / sigma of row r are read from row prow[r] (incremental decoding)
  // column groups: the segment is coded group by group -- all rows' columns [0, cg), then [cg, 2 cg), ... -- so that the
  // symbols of one row lie in `groups` different streams and a wavefront decoder advances them in parallel
  unsigned groups, cg, grp_elems;
};

struct Model { int m, R; float mu, inv; };     // inv = 1 / sigma, correctly rounded

__device__ __forceinline__ Model make_model(float mu, float sigma) {
  Model M;
  if (!(fabsf(mu) <= 2097152.f)) mu = mu > 0.f ? 2097152.f : (mu < 0.f ? -2097152.f : 0.f);   // NaN -> 0
  if (!(sigma >= 1e-6f)) sigma = 1e-6f;                                                        // NaN, <= 0 -> 1e-6
  if (sigma > 1e6f) sigma = 1e6f;
  M.mu = mu; M.inv = __frcp_rn(sigma);
  M.m = (int)rintf(mu);
  const float r = ceilf(__fmul_rn(6.f, sigma));
  M.R = r >= 1021.f ? 1023 : max(15, 2 + (int)r);
  return M;
}

__device__ __forceinline__ uint32_t phi24(float t) {
  float tq = __fmaf_rn(t, 128.f, 1024.f);                 // t * 128 is exact: the same value as mul-then-add
  tq = fminf(fmaxf(tq, 0.f), 2048.f);
  const int i = min((int)tq, LDIC_RANS_PHI_N - 1);
  const uint32_t f = (uint32_t)__fmul_rn(__fsub_rn(tq, (float)i), 4096.f);
  const uint32_t a = __ldg(c_phi + i), b = __ldg(c_phi + i + 1);
  return a + (((b - a) * f) >> 12);
}

// C(j), j in [0, Nsym]
__device__ __forceinline__ uint32_t cdf_at(const Model& M, int j) {
  const int nsym = 2 * M.R + 1;
  if (j <= 0) return 0u;
  if (j >= nsym) return 65536u;
  const float k = (float)(M.m - M.R + j);
  const float t = __fmul_rn(__fsub_rn(__fsub_rn(k, 0.5f), M.mu), M.inv);
  return (uint32_t)(((unsigned long long)phi24(t) * (uint32_t)(65536 - nsym)) >> 24) + (uint32_t)j;
}

// (mu, sigma) of element i of segment seg as the coder sees them; also the raw mean (for quant 2) and the address row / col
__device__ __forceinline__ void load_params(const Addr& A, long long seg, unsigned i, long long& row, unsigned& col,
                                            float& mu_raw, float& sigma) {
  unsigned r;
  if (A.groups > 1) {
    const unsigned g = i / A.grp_elems, rem = i - g * A.grp_elems;
    r = rem / A.cg;
    col = g * A.cg + (rem - r * A.cg);
  } else {
    r = i / A.cols;
    col = i - r * A.cols;
  }
  row = seg * A.rows_per_seg + r;
  mu_raw = 0.f;
  const long long prow = A.prow ? (long long)__ldg(A.prow + row) : row;
  if (A.mu_mode == 2) mu_raw = __ldg(A.mu + prow * A.mu_rs + A.mu_off + col);
  else if (A.mu_mode == 1) mu_raw = __ldg(A.mu + col);
  else if (A.mu_mode == 3) mu_raw = __ldg(A.mu + row % A.sg_period);
  float s;
  if (A.sg_mode == 2) s = __ldg(A.sigma + prow * A.sg_rs + A.sg_off + col);
  else if (A.sg_mode == 1) s = __ldg(A.sigma + col);
  else s = __ldg(A.sigma + row % A.sg_period);
  if (A.sigma_is_log) s = expf(s);
  if (A.scale_bound > 0.f) s = fmaxf(s, A.scale_bound);
  sigma = s;
}

// integer symbol of element (row, col); flags symbols the coder cannot represent (NaN, |k| > 2^30)
__device__ __forceinline__ int load_symbol(const Addr& A, long long row, unsigned col, float mu_raw, bool& bad) {
  const float v = __ldg(A.v + row * A.v_rs + A.v_off + col);
  const float kf = A.quant == 2 ? rintf(__fsub_rn(v, mu_raw)) : rintf(v);
  bad = !(fabsf(kf) <= 1073741824.f);
  return bad ? 0 : (int)kf;
}

// ---- encoder pass 1: (v, mu, sigma) -> packed (start | freq << 16); escapes counted per CTA -------------------------
__global__ void __launch_bounds__(kOpsThreads) k_rans_ops(Addr A, unsigned n, uint2* __restrict__ ops,
                                                          uint32_t* __restrict__ esc_count, uint32_t* __restrict__ status) {
  const long long seg = blockIdx.y;
  const unsigned base = blockIdx.x * kOpsChunk;
  int esc = 0;
  bool any_bad = false;
#pragma unroll
  for (int u = 0; u < kOpsPerThread; ++u) {
    const unsigned i = base + u * kOpsThreads + threadIdx.x;
    if (i < n) {
      long long row; unsigned col; float mu_raw, sigma;
      load_params(A, seg, i, row, col, mu_raw, sigma);
      bool bad;
      const int k = load_symbol(A, row, col, mu_raw, bad);
      any_bad |= bad;
      const Model M = make_model(A.quant == 2 ? 0.f : mu_raw, sigma);
      const int nsym = 2 * M.R + 1;
      long long jl = (long long)k - ((long long)M.m - M.R);
      const int j = jl <= 0 ? 0 : (jl >= nsym - 1 ? nsym - 1 : (int)jl);
      esc += (j == 0 || j == nsym - 1);
      const uint32_t start = cdf_at(M, j), freq = cdf_at(M, j + 1) - start;
      // floor(2^32 / freq) (2^32 - 1 for freq = 1): umulhi(x, rcp) is x / freq or one less
      const uint32_t rcp = freq == 1u ? 0xffffffffu : 0xffffffffu / freq + (0xffffffffu % freq == freq - 1u ? 1u : 0u);
      ops[seg * n + i] = make_uint2(start | (freq << 16), rcp);
    }
  }
  __shared__ int s_esc[kOpsThreads / 32];
  for (int o = 16; o > 0; o >>= 1) esc += __shfl_xor_sync(0xffffffffu, esc, o);
  if ((threadIdx.x & 31) == 0) s_esc[threadIdx.x >> 5] = esc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kOpsThreads / 32; ++w) t += s_esc[w];
    esc_count[seg * gridDim.x + blockIdx.x] = (uint32_t)t;
  }
  if (any_bad) atomicOr(status + seg, (uint32_t)ST_SYMBOL_RANGE);
}

// exclusive scan of `cnt[0..m)` by one CTA (any m); returns the total to every thread
__device__ uint32_t block_exclusive_scan(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ out, unsigned m) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (unsigned b = 0; b < m; b += blockDim.x) {
    const unsigned i = b + threadIdx.x;
    const uint32_t x = i < m ? cnt[i] : 0u;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < nw ? s_warp[lane] : 0u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      s_warp[lane] = w;                                    // inclusive over warps
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    const uint32_t wbase = warp ? s_warp[warp - 1] : 0u;
    if (i < m) out[i] = carry + wbase + inc - x;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_warp[nw - 1];
    __syncthreads();
  }
  return s_carry;
}

struct Layout { unsigned states_off, counts_off, esc_off; };
__host__ __device__ inline Layout layout_of(unsigned S) {
  Layout L;
  L.states_off = kHeaderBytes;
  L.counts_off = L.states_off + 4u * S;
  L.esc_off = L.counts_off + ((2u * S + 3u) & ~3u);
  return L;
}

// ---- encoder pass 2: scan the per-CTA escape counts (one CTA per segment) ------------------------------------------
__global__ void __launch_bounds__(1024) k_rans_esc_scan(const uint32_t* __restrict__ esc_count, uint32_t* __restrict__ esc_base,
                                                        uint32_t* __restrict__ esc_total, unsigned chunks) {
  const long long seg = blockIdx.x;
  const uint32_t t = block_exclusive_scan(esc_count + seg * chunks, esc_base + seg * chunks, chunks);
  if (threadIdx.x == 0) esc_total[seg] = t;
}

// ---- encoder pass 3: write the escape list in index order (escape <=> the op is one of the two window edges) -------
__global__ void __launch_bounds__(kOpsThreads) k_rans_escapes(Addr A, unsigned n, unsigned S, const uint2* __restrict__ ops,
                                                              const uint32_t* __restrict__ esc_count,
                                                              const uint32_t* __restrict__ esc_base, unsigned char* __restrict__ out,
                                                              long long out_stride, uint32_t* __restrict__ status) {
  const long long seg = blockIdx.y;
  if (esc_count[seg * gridDim.x + blockIdx.x] == 0) return;
  __shared__ uint32_t s_run;
  __shared__ uint32_t s_w[kOpsThreads / 32];
  if (threadIdx.x == 0) s_run = esc_base[seg * gridDim.x + blockIdx.x];
  __syncthreads();
  const Layout L = layout_of(S);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = 0; u < kOpsPerThread; ++u) {
    const unsigned i = blockIdx.x * kOpsChunk + u * kOpsThreads + threadIdx.x;
    bool is_esc = false;
    if (i < n) {
      const uint32_t op = ops[seg * n + i].x;
      const uint32_t start = op & 0xffffu, freq = op >> 16;
      is_esc = start == 0u || start + freq == 65536u;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, is_esc);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_run;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (is_esc) {
      const uint32_t e = before + __popc(bal & ((1u << lane) - 1u));
      long long row; unsigned col; float mu_raw, sigma;
      load_params(A, seg, i, row, col, mu_raw, sigma);
      bool bad;
      const int k = load_symbol(A, row, col, mu_raw, bad);
      const unsigned long long off = (unsigned long long)L.esc_off + 8ull * e;
      if ((long long)(off + 8) <= out_stride) {
        uint32_t* p = reinterpret_cast<uint32_t*>(out + seg * out_stride + off);
        p[0] = i; p[1] = (uint32_t)k;
      } else {
        atomicOr(status + seg, (uint32_t)ST_CAPACITY);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kOpsThreads / 32; ++w) t += s_w[w]; s_run += t; }
    __syncthreads();
  }
}

// ---- encoder pass 4: one thread per stream, the state recurrence backwards over the stream's run ---------------------
__device__ __forceinline__ unsigned run_length(unsigned n, unsigned S) { return S ? (n + S - 1) / S : 0u; }
__device__ __forceinline__ unsigned run_count(unsigned n, unsigned Ls, unsigned s) {
  const unsigned long long b = (unsigned long long)s * Ls;
  return b >= n ? 0u : min(Ls, (unsigned)(n - b));
}

// One thread per stream.  With one warp per scheduler every instruction costs its full latency (ncu: ~5.5 cycles per
// instruction, "wait" stalls), so the kernel is written for instruction count: each lane reads its own run straight from
// global memory in blocks of 16 ops (two register buffers, the next block in flight under the dependent chain of the
// current one; a 32-byte sector serves four consecutive ops of a lane) and stores its words straight back (2-byte
// stores, merged in L2).  Per symbol the chain is compare -> select -> multiply-high -> multiply-subtract -> compare ->
// add, without a branch: a full block is a straight line of 16 steps, only the ragged top block of a run is guarded.
constexpr int kEncBlk = 16;

__device__ __forceinline__ void rans_put(uint32_t& x, uint16_t*& wp, const uint2 op) {
  const uint32_t start = op.x & 0xffffu, freq = op.x >> 16;
  const bool emit = x >= (freq << 16);
  if (emit) *wp = (uint16_t)x;
  wp += emit ? 1 : 0;
  x = emit ? x >> 16 : x;
  const uint32_t q = __umulhi(x, op.y);
  const uint32_t r = x - q * freq;                       // q is floor(x / freq) or one less
  x = (q << 16) + r + start + (r >= freq ? 65536u - freq : 0u);
}

__device__ __forceinline__ void rans_load_block(uint2 (&op)[kEncBlk], const uint2* __restrict__ o, int b, unsigned cnt) {
#pragma unroll
  for (int u = 0; u < kEncBlk; ++u) {
    const unsigned p = (unsigned)b * kEncBlk + u;
    op[u] = (b >= 0 && p < cnt) ? __ldg(o + p) : make_uint2(1u << 16, 0u);
  }
}

__device__ __forceinline__ void rans_chain_block(uint32_t& x, uint16_t*& wp, const uint2 (&op)[kEncBlk], int b, unsigned cnt) {
  if ((unsigned)(b + 1) * kEncBlk <= cnt) {
#pragma unroll
    for (int u = kEncBlk - 1; u >= 0; --u) rans_put(x, wp, op[u]);
  } else {
#pragma unroll
    for (int u = kEncBlk - 1; u >= 0; --u)
      if ((unsigned)b * kEncBlk + u < cnt) rans_put(x, wp, op[u]);
  }
}

__global__ void __launch_bounds__(kStreamThreads) k_rans_enc_streams(const uint2* __restrict__ ops, unsigned n, unsigned S,
                                                                     uint16_t* __restrict__ slab, uint32_t* __restrict__ states,
                                                                     uint32_t* __restrict__ wcount) {
  const long long seg = blockIdx.y;
  const unsigned s = blockIdx.x * kStreamThreads + threadIdx.x;
  if (s >= S) return;
  const unsigned Ls = run_length(n, S), cnt = run_count(n, Ls, s);
  const uint2* o = ops + seg * n + (size_t)s * Ls;
  uint16_t* wrow = slab + seg * n + (size_t)s * Ls;
  uint32_t x = kRansL;
  uint16_t* wp = wrow;
  uint2 A[kEncBlk], B[kEncBlk];
  int b = (int)((cnt + kEncBlk - 1) / kEncBlk) - 1;      // top block (may be ragged)
  rans_load_block(A, o, b, cnt);
  for (; b >= 0; b -= 2) {
    rans_load_block(B, o, b - 1, cnt);
    rans_chain_block(x, wp, A, b, cnt);
    if (b - 1 < 0) break;
    rans_load_block(A, o, b - 2, cnt);
    rans_chain_block(x, wp, B, b - 1, cnt);
  }
  states[seg * S + s] = x;
  wcount[seg * S + s] = (uint32_t)(wp - wrow);
}

// ---- encoder pass 5: scan the word counts, write header / states / counts (one CTA per segment) --------------------
__global__ void __launch_bounds__(1024) k_rans_enc_scan(unsigned n, unsigned S, int quant, unsigned groups,
                                                        const uint32_t* __restrict__ states,
                                                        const uint32_t* __restrict__ wcount, uint32_t* __restrict__ woffs,
                                                        const uint32_t* __restrict__ esc_total, unsigned char* __restrict__ out,
                                                        long long out_stride, uint32_t* __restrict__ sizes,
                                                        uint32_t* __restrict__ status) {
  const long long seg = blockIdx.x;
  const uint32_t W = block_exclusive_scan(wcount + seg * S, woffs + seg * S, S);
  const Layout L = layout_of(S);
  const uint32_t E = esc_total[seg];
  const unsigned long long total = (unsigned long long)L.esc_off + 8ull * E + 2ull * W;
  unsigned char* o = out + seg * out_stride;
  const bool fits = (long long)total <= out_stride && total <= 0xffffffffull;
  if (threadIdx.x == 0) {
    if (!fits) atomicOr(status + seg, (uint32_t)ST_CAPACITY);
    sizes[seg] = fits ? (uint32_t)total : 0u;
    if (out_stride >= kHeaderBytes) {
      uint32_t* h = reinterpret_cast<uint32_t*>(o);
      h[0] = kMagic; h[1] = n; h[2] = S; h[3] = E; h[4] = W; h[5] = (uint32_t)quant; h[6] = groups > 1 ? groups : 0; h[7] = 0;
    }
  }
  if ((long long)L.esc_off > out_stride) return;
  uint32_t* st = reinterpret_cast<uint32_t*>(o + L.states_off);
  uint16_t* ct = reinterpret_cast<uint16_t*>(o + L.counts_off);
  for (unsigned s = threadIdx.x; s < S; s += blockDim.x) { st[s] = states[seg * S + s]; ct[s] = (uint16_t)wcount[seg * S + s]; }
  if ((S & 1u) && threadIdx.x == 0) ct[S] = 0;            // pad
}

// ---- encoder pass 6: concatenate the streams' words, each reversed into decoding order ------------------------------
__global__ void __launch_bounds__(kOpsThreads) k_rans_pack(unsigned n, unsigned S, const uint16_t* __restrict__ slab,
                                                           const uint32_t* __restrict__ wcount, const uint32_t* __restrict__ woffs,
                                                           const uint32_t* __restrict__ esc_total, const uint32_t* __restrict__ sizes,
                                                           unsigned char* __restrict__ out, long long out_stride) {
  const long long seg = blockIdx.y;
  if (sizes[seg] == 0) return;                            // did not fit
  const Layout L = layout_of(S);
  uint16_t* words = reinterpret_cast<uint16_t*>(out + seg * out_stride + L.esc_off + 8ull * esc_total[seg]);
  const unsigned Ls = run_length(n, S);
#pragma unroll
  for (int u = 0; u < kOpsPerThread; ++u) {
    const unsigned i = blockIdx.x * kOpsChunk + u * kOpsThreads + threadIdx.x;
    if (i >= n) continue;
    const unsigned s = i / Ls, w = i - s * Ls;
    const uint32_t c = wcount[seg * S + s];
    if (w < c) words[woffs[seg * S + s] + (c - 1 - w)] = slab[seg * n + i];
  }
}

// ---- decoder pass 1: validate the header, scan the word counts (one CTA per segment) -------------------------------
__global__ void __launch_bounds__(1024) k_rans_dec_scan(unsigned n, unsigned S, int quant, unsigned groups,
                                                        const unsigned char* __restrict__ in,
                                                        long long in_stride, const uint32_t* __restrict__ sizes,
                                                        uint32_t* __restrict__ wcount, uint32_t* __restrict__ woffs,
                                                        uint32_t* __restrict__ status) {
  const long long seg = blockIdx.x;
  const unsigned char* src = in + seg * in_stride;
  const uint32_t size = sizes[seg];
  const Layout L = layout_of(S);
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    int ok = size >= (uint32_t)kHeaderBytes && (long long)size <= in_stride;
    if (ok) {
      const uint32_t* h = reinterpret_cast<const uint32_t*>(src);
      ok = h[0] == kMagic && h[1] == n && h[2] == S && h[5] == (uint32_t)quant && h[6] == (groups > 1 ? groups : 0u) &&
           h[3] <= n && h[4] <= n &&
           (unsigned long long)L.esc_off + 8ull * h[3] + 2ull * h[4] == size;
    }
    s_ok = ok;
    if (!ok) atomicOr(status + seg, (uint32_t)ST_HEADER);
  }
  __syncthreads();
  if (!s_ok) { for (unsigned s = threadIdx.x; s < S; s += blockDim.x) wcount[seg * S + s] = 0xffffffffu; return; }
  const uint16_t* ct = reinterpret_cast<const uint16_t*>(src + L.counts_off);
  for (unsigned s = threadIdx.x; s < S; s += blockDim.x) wcount[seg * S + s] = ct[s];
  __syncthreads();
  const uint32_t W = block_exclusive_scan(wcount + seg * S, woffs + seg * S, S);
  if (threadIdx.x == 0 && W != reinterpret_cast<const uint32_t*>(src)[4]) {
    atomicOr(status + seg, (uint32_t)ST_CORRUPT);
    wcount[seg * S] = 0xffffffffu;                        // poisons stream 0; the others stay inside W by the check below
  }
}

// ---- decoder pass 2: one warp per stream ---------------------------------------------------------------------------
constexpr int kDecWarps = 4;

// C(j) for 0 < j < Nsym from base = (float)(m - R) - 1/2 (exact) and scale = 65536 - Nsym: the same bits as cdf_at
__device__ __forceinline__ uint32_t cdf_from(float base, float mu, float inv, uint32_t scale, int j) {
  const float t = __fmul_rn(__fsub_rn(__fadd_rn(base, (float)j), mu), inv);
  return (uint32_t)(((unsigned long long)phi24(t) * scale) >> 24) + (uint32_t)j;
}

// Per symbol the 32 lanes own the 32 sub-intervals [l step, (l+1) step) of the window (step = ceil(Nsym / 32): 1 for
// windows of up to 32 integers); a lane's cumulative bounds do not depend on the coder state, so they are computed one
// symbol ahead, off the dependent chain  slot -> "is it mine" ballot -> winner's state update -> shuffle -> next slot.
// Wider windows (sigma > 2.2) continue with a 32-ary search inside the winner's interval.
struct Cand { float base, mu, inv; uint32_t scale; int step, nsym; uint32_t c_lo, c_hi; };

// where the decoded symbols go: fp32 (always) and optionally a bf16 copy (the rounded latent image the context model
// and the synthesis transform read), both with row addressing
struct DecOut { float* v; long long rs, off; __nv_bfloat16* vb; long long vb_rs, vb_off; const int* vb_map; };
__device__ __forceinline__ void dec_store(const DecOut& O, long long row, unsigned col, float vv) {
  O.v[row * O.rs + O.off + col] = vv;
  if (O.vb) O.vb[(O.vb_map ? (long long)__ldg(O.vb_map + row) : row) * O.vb_rs + O.vb_off + col] = __float2bfloat16_rn(vv);
}

struct DecCursor { uint32_t x, wpos, wend, wnext; bool corrupt; };

// Decodes positions [p_begin, p_end) of stream s (symbols base_i + p of segment seg); warp-collective.
__device__ __forceinline__ void decode_run(const Addr& A, long long seg, unsigned base_i, unsigned p_begin, unsigned p_end,
                                           const uint16_t* __restrict__ words, DecCursor& cur_, const DecOut& O,
                                           float4* s_par, int lane) {
  uint32_t x = cur_.x, wpos = cur_.wpos, wnext = cur_.wnext;
  const uint32_t wend = cur_.wend;
  bool corrupt = cur_.corrupt;
  for (unsigned p0 = p_begin; p0 < p_end; p0 += 32) {
    // each lane fetches the parameters of one of the next 32 symbols (coalesced), then the warp decodes them in order
    const bool valid = p0 + lane < p_end;
    long long row = 0; unsigned col = 0; float mu_raw = 0.f, sigma = 1.f;
    if (valid) load_params(A, seg, base_i + p0 + lane, row, col, mu_raw, sigma);
    const Model Mm = make_model(A.quant == 2 ? 0.f : mu_raw, sigma);
    const int my_k0 = Mm.m - Mm.R;
    __syncwarp();
    s_par[lane] = make_float4(Mm.mu, Mm.inv, (float)my_k0 - 0.5f, __int_as_float(Mm.R));
    __syncwarp();
    int my_j = 0;
    const int todo = min(32u, p_end - p0);
    auto prep = [&](int u) {
      Cand q;
      const float4 P = s_par[u];
      q.mu = P.x; q.inv = P.y; q.base = P.z;
      q.nsym = 2 * __float_as_int(P.w) + 1;
      q.scale = (uint32_t)(65536 - q.nsym);
      q.step = (q.nsym + 31) >> 5;
      const int jh = (lane + 1) * q.step;                  // this lane owns [jh - step, min(jh, nsym))
      q.c_hi = jh < q.nsym ? cdf_from(q.base, q.mu, q.inv, q.scale, jh) : 65536u;
      q.c_lo = __shfl_up_sync(0xffffffffu, q.c_hi, 1);
      if (lane == 0) q.c_lo = 0u;
      return q;
    };
    Cand cur = prep(0);
    for (int u = 0; u < todo; ++u) {
      Cand nxt = cur;
      if (u + 1 < todo) nxt = prep(u + 1);
      const uint32_t slot = x & 0xffffu;
      // C is strictly increasing, so exactly one lane has c_lo <= slot < c_hi (empty tail intervals are [65536, 65536))
      const int w = __ffs(__ballot_sync(0xffffffffu, cur.c_lo <= slot && slot < cur.c_hi)) - 1;
      int j;
      if (cur.step == 1) {
        const uint32_t xw = (cur.c_hi - cur.c_lo) * (x >> 16) + slot - cur.c_lo;
        x = __shfl_sync(0xffffffffu, xw, w);
        j = w;
      } else {
        int lo = w * cur.step, hi = min(lo + cur.step, cur.nsym);
        uint32_t c_lo = __shfl_sync(0xffffffffu, cur.c_lo, w), c_hi = __shfl_sync(0xffffffffu, cur.c_hi, w);
        while (hi - lo > 1) {                              // 32-ary search inside [lo, hi): C(lo) <= slot < C(hi)
          const int step = (hi - lo + 31) >> 5;
          const int jl = lo + (lane + 1) * step;
          const uint32_t cj = jl < hi ? cdf_from(cur.base, cur.mu, cur.inv, cur.scale, jl) : 0xffffffffu;
          const int t = __popc(__ballot_sync(0xffffffffu, cj <= slot));
          const uint32_t c_prev = __shfl_sync(0xffffffffu, cj, max(t - 1, 0));
          const uint32_t c_next = __shfl_sync(0xffffffffu, cj, min(t, 31));
          const int jn = lo + (t + 1) * step;
          if (t > 0) c_lo = c_prev;
          if (t < 32 && jn < hi) { hi = jn; c_hi = c_next; }
          lo += t * step;
        }
        x = (c_hi - c_lo) * (x >> 16) + slot - c_lo;
        j = lo;
      }
      if (x < kRansL) {
        if (wpos < wend) {
          x = (x << 16) | wnext;
          ++wpos;
          wnext = wpos < wend ? words[wpos] : 0u;
        } else { corrupt = true; x |= kRansL; }
      }
      if (lane == u) my_j = j;
      cur = nxt;
    }
    if (valid) {
      const float kf = (float)(my_k0 + my_j);
      dec_store(O, row, col, A.quant == 2 ? __fadd_rn(kf, mu_raw) : kf);
    }
  }
  cur_.x = x; cur_.wpos = wpos; cur_.wnext = wnext; cur_.corrupt = corrupt;
}

__global__ void __launch_bounds__(kDecWarps * 32) k_rans_dec_streams(Addr A, unsigned n, unsigned S,
                                                                     const unsigned char* __restrict__ in, long long in_stride,
                                                                     const uint32_t* __restrict__ wcount,
                                                                     const uint32_t* __restrict__ woffs, DecOut O,
                                                                     uint32_t* __restrict__ status) {
  __shared__ float4 s_par[kDecWarps][32];
  const long long seg = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned s = blockIdx.x * kDecWarps + wib;
  if (s >= S) return;
  const uint32_t c = wcount[seg * S + s];
  if (c == 0xffffffffu) return;                           // header rejected
  const unsigned char* src = in + seg * in_stride;
  const uint32_t* h = reinterpret_cast<const uint32_t*>(src);
  const Layout L = layout_of(S);
  const uint32_t W = h[4];
  const uint16_t* words = reinterpret_cast<const uint16_t*>(src + L.esc_off + 8ull * h[3]);
  DecCursor cur;
  cur.wpos = woffs[seg * S + s];
  cur.wend = cur.wpos + c;
  if (cur.wend > W) { if (lane == 0) atomicOr(status + seg, (uint32_t)ST_CORRUPT); return; }
  cur.x = reinterpret_cast<const uint32_t*>(src + L.states_off)[s];
  cur.wnext = cur.wpos < cur.wend ? words[cur.wpos] : 0u; // the next word, fetched before it is needed
  cur.corrupt = false;
  const unsigned Ls = run_length(n, S), cnt = run_count(n, Ls, s);
  decode_run(A, seg, s * Ls, 0u, cnt, words, cur, O, s_par[wib], lane);
  if (lane == 0 && (cur.corrupt || cur.x != kRansL || cur.wpos != cur.wend)) atomicOr(status + seg, (uint32_t)ST_CORRUPT);
}

// ---- incremental decoding (a decoder whose (mu, sigma) depend on symbols decoded earlier, e.g. the causal context
// model of model/net.py:289-319 walked along wavefronts): the streams' cursors live in a caller-owned state buffer
// (uint4 {x, next word, end word, next symbol of the run} per stream), k_rans_dec_init fills it from the bitstream,
// k_rans_dec_ranges decodes one [first, first + count) symbol range per warp (inside ONE stream, continuing exactly
// where that stream stopped) and applies the escapes that fall into the range.
__global__ void __launch_bounds__(256) k_rans_dec_init(unsigned S, const unsigned char* __restrict__ in, long long in_stride,
                                                       const uint32_t* __restrict__ wcount, const uint32_t* __restrict__ woffs,
                                                       uint4* __restrict__ state, uint32_t* __restrict__ status) {
  const long long seg = blockIdx.y;
  const unsigned s = blockIdx.x * 256 + threadIdx.x;
  if (s >= S) return;
  const uint32_t c = wcount[seg * S + s];
  const unsigned char* src = in + seg * in_stride;
  uint4 st = make_uint4(0u, 0u, 0u, 0xffffffffu);         // next = 2^32-1: unusable
  if (c != 0xffffffffu) {
    const uint32_t W = reinterpret_cast<const uint32_t*>(src)[4];
    const uint32_t wpos = woffs[seg * S + s];
    if (wpos + c <= W) st = make_uint4(reinterpret_cast<const uint32_t*>(src + layout_of(S).states_off)[s], wpos, wpos + c, 0u);
    else atomicOr(status + seg, (uint32_t)ST_CORRUPT);
  }
  state[seg * S + s] = st;
}

__global__ void __launch_bounds__(kDecWarps * 32) k_rans_dec_ranges(Addr A, unsigned n, unsigned S,
                                                                    const unsigned char* __restrict__ in, long long in_stride,
                                                                    uint4* __restrict__ state, const int* __restrict__ ranges,
                                                                    int nranges, DecOut O, uint32_t* __restrict__ status) {
  __shared__ float4 s_par[kDecWarps][32];
  const long long seg = blockIdx.y;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int j = blockIdx.x * kDecWarps + wib;
  if (j >= nranges) return;
  const int first = ranges[2 * j], count = ranges[2 * j + 1];
  if (count <= 0) return;
  const unsigned Ls = run_length(n, S);
  const unsigned s = (unsigned)first / Ls, p = (unsigned)first - s * Ls;
  const unsigned cnt = s < S ? run_count(n, Ls, s) : 0u;
  if (first < 0 || s >= S || p + (unsigned)count > cnt) { if (lane == 0) atomicOr(status + seg, (uint32_t)ST_ORDER); return; }
  const uint4 st = state[seg * S + s];
  if (st.w != p) { if (lane == 0) atomicOr(status + seg, (uint32_t)ST_ORDER); return; }   // not where this stream stopped
  const unsigned char* src = in + seg * in_stride;
  const uint32_t* h = reinterpret_cast<const uint32_t*>(src);
  const uint32_t E = h[3];
  const Layout L = layout_of(S);
  const uint16_t* words = reinterpret_cast<const uint16_t*>(src + L.esc_off + 8ull * E);
  DecCursor cur;
  cur.x = st.x; cur.wpos = st.y; cur.wend = st.z; cur.corrupt = false;
  cur.wnext = cur.wpos < cur.wend ? words[cur.wpos] : 0u;
  decode_run(A, seg, s * Ls, p, p + (unsigned)count, words, cur, O, s_par[wib], lane);
  const unsigned next = p + (unsigned)count;
  if (lane == 0) {
    state[seg * S + s] = make_uint4(cur.x, cur.wpos, cur.wend, next);
    if (cur.corrupt || (next == cnt && (cur.x != kRansL || cur.wpos != cur.wend))) atomicOr(status + seg, (uint32_t)ST_CORRUPT);
  }
  // escapes inside [first, first + count): the list is sorted by index
  if (E) {
    const uint32_t* esc = reinterpret_cast<const uint32_t*>(src + L.esc_off);
    uint32_t lo = 0, hi = E;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (esc[2 * mid] < (uint32_t)first) lo = mid + 1; else hi = mid; }
    __syncwarp();
    for (uint32_t e = lo + lane; e < E; e += 32) {
      const uint32_t i = esc[2 * e];
      if (i >= (uint32_t)(first + count)) break;
      long long row; unsigned col; float mu_raw, sigma;
      load_params(A, seg, i, row, col, mu_raw, sigma);
      const float kf = (float)(int)esc[2 * e + 1];
      dec_store(O, row, col, A.quant == 2 ? __fadd_rn(kf, mu_raw) : kf);
    }
  }
}

// ---- decoder pass 3: overwrite the escaped symbols ----------------------------------------------------------------
__global__ void __launch_bounds__(kOpsThreads) k_rans_dec_escapes(Addr A, unsigned n, unsigned S, const unsigned char* __restrict__ in,
                                                                  long long in_stride, const uint32_t* __restrict__ wcount,
                                                                  float* __restrict__ v_hat, long long vh_rs, long long vh_off,
                                                                  uint32_t* __restrict__ status) {
  const long long seg = blockIdx.y;
  if (S == 0 || wcount[seg * S] == 0xffffffffu) return;
  const unsigned char* src = in + seg * in_stride;
  const uint32_t E = reinterpret_cast<const uint32_t*>(src)[3];
  const uint32_t* esc = reinterpret_cast<const uint32_t*>(src + layout_of(S).esc_off);
  for (unsigned e = blockIdx.x * kOpsThreads + threadIdx.x; e < E; e += gridDim.x * kOpsThreads) {
    const uint32_t i = esc[2 * e];
    const int k = (int)esc[2 * e + 1];
    if (i >= n) { atomicOr(status + seg, (uint32_t)ST_CORRUPT); continue; }
    long long row; unsigned col; float mu_raw, sigma;
    load_params(A, seg, i, row, col, mu_raw, sigma);
    const float kf = (float)k;
    v_hat[row * vh_rs + vh_off + col] = A.quant == 2 ? __fadd_rn(kf, mu_raw) : kf;
  }
}

struct Ws {                       // carved out of the caller's workspace
  uint2* ops;
  uint32_t *esc_count, *esc_base, *esc_total, *states, *wcount, *woffs;
  uint16_t* slab;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(void* base, long long segs, long long n, unsigned S, Ws* w) {
  const size_t chunks = (size_t)((n + kOpsChunk - 1) / kOpsChunk);
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += align_up(bytes); return p; };
  uint2* ops = (uint2*)take(8ull * segs * n);
  uint16_t* slab = (uint16_t*)take(2ull * segs * n);
  uint32_t* esc_count = (uint32_t*)take(4ull * segs * chunks);
  uint32_t* esc_base = (uint32_t*)take(4ull * segs * chunks);
  uint32_t* esc_total = (uint32_t*)take(4ull * segs);
  uint32_t* states = (uint32_t*)take(4ull * segs * S);
  uint32_t* wcount = (uint32_t*)take(4ull * segs * S);
  uint32_t* woffs = (uint32_t*)take(4ull * segs * S);
  if (w) { w->ops = ops; w->slab = slab; w->esc_count = esc_count; w->esc_base = esc_base; w->esc_total = esc_total;
           w->states = states; w->wcount = wcount; w->woffs = woffs; }
  return off;
}

int make_addr(const LdicRansArgs* a, bool need_v, Addr* A, long long* segs, long long* n) {
  if (!a) return fail(LDIC_EINVAL, "rans: null args");
  if (a->rows < 0 || a->cols <= 0 || a->cols > 0x7fffffffLL) return fail(LDIC_EINVAL, "rans: rows / cols");
  if (a->rows_per_segment <= 0 || a->rows % a->rows_per_segment) return fail(LDIC_EINVAL, "rans: rows must be a multiple of rows_per_segment");
  const long long ne = a->rows_per_segment * a->cols;
  if (ne > 0x7fffffffLL) return fail(LDIC_EINVAL, "rans: a segment holds at most 2^31-1 symbols");
  if (a->streams <= 0 || a->streams > (1 << 20)) return fail(LDIC_EINVAL, "rans: streams must be in [1, 2^20]");
  if ((ne + a->streams - 1) / a->streams > 65535) return fail(LDIC_EINVAL, "rans: more than 65535 symbols per stream; raise `streams`");
  if (a->quant != 1 && a->quant != 2) return fail(LDIC_EINVAL, "rans: quant must be 1 (round v) or 2 (round(v - mu) + mu)");
  if (a->mu_mode < 0 || a->mu_mode > 3 || a->sigma_mode < 1 || a->sigma_mode > 3) return fail(LDIC_EINVAL, "rans: bad broadcast mode");
  if ((a->mu_mode != 0 && !a->mu) || !a->sigma || (need_v && !a->v)) return fail(LDIC_EINVAL, "rans: null tensor");
  if ((a->sigma_mode == 3 || a->mu_mode == 3) && a->sigma_period <= 0) return fail(LDIC_EINVAL, "rans: sigma_period");
  A->v = a->v; A->v_rs = a->v_rs; A->v_off = a->v_off;
  A->mu = a->mu; A->mu_rs = a->mu_rs; A->mu_off = a->mu_off; A->mu_mode = a->mu_mode;
  A->sigma = a->sigma; A->sg_rs = a->sigma_rs; A->sg_off = a->sigma_off; A->sg_mode = a->sigma_mode;
  A->sg_period = a->sigma_period > 0 ? a->sigma_period : 1;
  A->cols = (unsigned)a->cols; A->rows_per_seg = a->rows_per_segment;
  A->quant = a->quant; A->sigma_is_log = a->sigma_is_log; A->scale_bound = a->scale_bound;
  A->prow = nullptr;
  A->groups = a->col_groups > 1 ? (unsigned)a->col_groups : 1u;
  if (a->col_groups < 0 || a->cols % A->groups) return fail(LDIC_EINVAL, "rans: cols must be a multiple of col_groups");
  A->cg = (unsigned)(a->cols / A->groups);
  A->grp_elems = (unsigned)(a->rows_per_segment * A->cg);
  *segs = a->rows / a->rows_per_segment;
  *n = ne;
  if (*segs > 65535) return fail(LDIC_EINVAL, "rans: at most 65535 segments per call");
  return LDIC_OK;
}

}  // namespace

extern "C" {

LDIC_API const unsigned int* ldic_rans_phi_table(int* entries) {
  if (entries) *entries = LDIC_RANS_PHI_N + 1;
  return h_phi;
}

LDIC_API size_t ldic_rans_max_bytes(long long seg_elems, int streams) {
  if (seg_elems < 0 || streams <= 0) return 0;
  const Layout L = layout_of((unsigned)streams);
  return (((size_t)L.esc_off + 10ull * (size_t)seg_elems) + 15) & ~(size_t)15;   // every symbol escaped and one word each
}

LDIC_API size_t ldic_rans_workspace_bytes(long long segments, long long seg_elems, int streams) {
  if (segments < 0 || seg_elems < 0 || streams <= 0) return 0;
  return carve(nullptr, segments, seg_elems, (unsigned)streams, nullptr);
}

LDIC_API int ldic_rans_encode(const LdicRansArgs* a, unsigned char* out, long long out_stride, unsigned int* sizes,
                              unsigned int* status, void* workspace, void* stream) {
  Addr A; long long segs, n;
  if (int r = make_addr(a, true, &A, &segs, &n)) return r;
  if (!out || !sizes || !status || !workspace) return fail(LDIC_EINVAL, "rans encode: null buffer");
  if (out_stride < kHeaderBytes || (out_stride & 3)) return fail(LDIC_EINVAL, "rans encode: out_stride must be a multiple of 4 and hold the header");
  if (((uintptr_t)out & 3) || ((uintptr_t)workspace & 255)) return fail(LDIC_EINVAL, "rans encode: out must be 4-byte, workspace 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (segs == 0) return LDIC_OK;
  const unsigned S = (unsigned)a->streams;
  Ws w; carve(workspace, segs, n, S, &w);
  LDIC_CUDA(cudaMemsetAsync(status, 0, 4 * segs, st));
  const unsigned chunks = (unsigned)((n + kOpsChunk - 1) / kOpsChunk);
  if (n > 0) {
    k_rans_ops<<<dim3(chunks, (unsigned)segs), kOpsThreads, 0, st>>>(A, (unsigned)n, w.ops, w.esc_count, status);
    if (int r = check_launch("k_rans_ops")) return r;
    k_rans_esc_scan<<<(unsigned)segs, 1024, 0, st>>>(w.esc_count, w.esc_base, w.esc_total, chunks);
    if (int r = check_launch("k_rans_esc_scan")) return r;
    k_rans_escapes<<<dim3(chunks, (unsigned)segs), kOpsThreads, 0, st>>>(A, (unsigned)n, S, w.ops, w.esc_count, w.esc_base, out,
                                                                        out_stride, status);
    if (int r = check_launch("k_rans_escapes")) return r;
  } else {
    LDIC_CUDA(cudaMemsetAsync(w.esc_total, 0, 4 * segs, st));
  }
  k_rans_enc_streams<<<dim3((S + kStreamThreads - 1) / kStreamThreads, (unsigned)segs), kStreamThreads, 0, st>>>(
      w.ops, (unsigned)n, S, w.slab, w.states, w.wcount);
  if (int r = check_launch("k_rans_enc_streams")) return r;
  k_rans_enc_scan<<<(unsigned)segs, 1024, 0, st>>>((unsigned)n, S, a->quant, A.groups, w.states, w.wcount, w.woffs, w.esc_total, out, out_stride,
                                                   sizes, status);
  if (int r = check_launch("k_rans_enc_scan")) return r;
  if (n > 0) {
    k_rans_pack<<<dim3(chunks, (unsigned)segs), kOpsThreads, 0, st>>>((unsigned)n, S, w.slab, w.wcount, w.woffs, w.esc_total, sizes, out,
                                                                     out_stride);
    if (int r = check_launch("k_rans_pack")) return r;
  }
  return LDIC_OK;
}

LDIC_API int ldic_rans_decode(const LdicRansArgs* a, const unsigned char* in, long long in_stride, const unsigned int* sizes,
                              float* v_hat, long long v_hat_rs, long long v_hat_off, unsigned int* status, void* workspace,
                              void* stream) {
  Addr A; long long segs, n;
  if (int r = make_addr(a, false, &A, &segs, &n)) return r;
  if (!in || !sizes || !status || !workspace || !v_hat) return fail(LDIC_EINVAL, "rans decode: null buffer");
  if ((in_stride & 3) || ((uintptr_t)in & 3) || ((uintptr_t)workspace & 255)) return fail(LDIC_EINVAL, "rans decode: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  if (segs == 0) return LDIC_OK;
  const unsigned S = (unsigned)a->streams;
  Ws w; carve(workspace, segs, n, S, &w);
  LDIC_CUDA(cudaMemsetAsync(status, 0, 4 * segs, st));
  k_rans_dec_scan<<<(unsigned)segs, 1024, 0, st>>>((unsigned)n, S, a->quant, A.groups, in, in_stride, sizes, w.wcount, w.woffs, status);
  if (int r = check_launch("k_rans_dec_scan")) return r;
  if (n == 0) return LDIC_OK;
  DecOut O; O.v = v_hat; O.rs = v_hat_rs; O.off = v_hat_off; O.vb = nullptr; O.vb_rs = 0; O.vb_off = 0; O.vb_map = nullptr;
  k_rans_dec_streams<<<dim3((S + kDecWarps - 1) / kDecWarps, (unsigned)segs), kDecWarps * 32, 0, st>>>(
      A, (unsigned)n, S, in, in_stride, w.wcount, w.woffs, O, status);
  if (int r = check_launch("k_rans_dec_streams")) return r;
  k_rans_dec_escapes<<<dim3(8, (unsigned)segs), kOpsThreads, 0, st>>>(A, (unsigned)n, S, in, in_stride, w.wcount, v_hat, v_hat_rs,
                                                                     v_hat_off, status);
  if (int r = check_launch("k_rans_dec_escapes")) return r;
  return LDIC_OK;
}

LDIC_API int ldic_rans_decode_begin(const LdicRansArgs* a, const unsigned char* in, long long in_stride, const unsigned int* sizes,
                                    void* state, unsigned int* status, void* workspace, void* stream) {
  Addr A; long long segs, n;
  if (int r = make_addr(a, false, &A, &segs, &n)) return r;
  if (!in || !sizes || !status || !workspace || !state) return fail(LDIC_EINVAL, "rans decode_begin: null buffer");
  if ((in_stride & 3) || ((uintptr_t)in & 3) || ((uintptr_t)workspace & 255) || ((uintptr_t)state & 15))
    return fail(LDIC_EINVAL, "rans decode_begin: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  if (segs == 0) return LDIC_OK;
  const unsigned S = (unsigned)a->streams;
  Ws w; carve(workspace, segs, n, S, &w);
  LDIC_CUDA(cudaMemsetAsync(status, 0, 4 * segs, st));
  k_rans_dec_scan<<<(unsigned)segs, 1024, 0, st>>>((unsigned)n, S, a->quant, A.groups, in, in_stride, sizes, w.wcount, w.woffs, status);
  if (int r = check_launch("k_rans_dec_scan")) return r;
  k_rans_dec_init<<<dim3((S + 255) / 256, (unsigned)segs), 256, 0, st>>>(S, in, in_stride, w.wcount, w.woffs, (uint4*)state, status);
  return check_launch("k_rans_dec_init");
}

LDIC_API int ldic_rans_decode_ranges(const LdicRansArgs* a, const unsigned char* in, long long in_stride, void* state,
                                     const int* ranges, int nranges, float* v_hat, long long v_hat_rs, long long v_hat_off,
                                     void* v_hat_bf16, long long vb_rs, long long vb_off, const int* param_row_map,
                                     const int* bf16_row_map, unsigned int* status, void* stream) {
  Addr A; long long segs, n;
  if (int r = make_addr(a, false, &A, &segs, &n)) return r;
  if (!in || !status || !state || !v_hat || (nranges > 0 && !ranges)) return fail(LDIC_EINVAL, "rans decode_ranges: null buffer");
  if (nranges < 0 || nranges > 65535 * kDecWarps) return fail(LDIC_EINVAL, "rans decode_ranges: nranges");
  if (segs == 0 || nranges == 0 || n == 0) return LDIC_OK;
  DecOut O; O.v = v_hat; O.rs = v_hat_rs; O.off = v_hat_off;
  O.vb = reinterpret_cast<__nv_bfloat16*>(v_hat_bf16); O.vb_rs = vb_rs; O.vb_off = vb_off; O.vb_map = bf16_row_map;
  A.prow = param_row_map;
  k_rans_dec_ranges<<<dim3((nranges + kDecWarps - 1) / kDecWarps, (unsigned)segs), kDecWarps * 32, 0, (cudaStream_t)stream>>>(
      A, (unsigned)n, (unsigned)a->streams, in, in_stride, (uint4*)state, ranges, nranges, O, status);
  return check_launch("k_rans_dec_ranges");
}

}  // extern "C"

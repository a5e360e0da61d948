// Window attention core: softmax(q k^T + rel-pos bias + shift mask) v for one 8x8 (or 4x4) window
// per CTA, one warp per head, bf16 tensor-core MMAs with fp32 softmax.
//   reference: layers/win_attention.py:96-126 (WindowAttention.forward between the qkv and proj Linears),
//              :6-36 (window_partition / window_reverse), :160-193 (shift mask, cyclic shift).
// The qkv and proj projections run as 1x1 convs on the tcgen05 kernel (conv_tc.cu); this kernel does the part in
// between.  Window partition, cyclic shift and its reverse are index arithmetic on the NHWC token image: token
// (i, j) of window (wy, wx) is pixel ((wy*ws + i + shift) mod H, (wx*ws + j + shift) mod W), read and written in
// place (no rolled / partitioned copies).  The 0 / -100 shift mask is recomputed from the region ids of the two
// tokens (:160-177) instead of being built on the host every call.
//
// Per CTA: the window's q | k | v rows (N tokens x 3*C bf16) are staged in shared memory with coalesced 16-byte
// loads (row pitch padded by 16 B so that the 8 fragment rows hit different banks).
// Per warp (= head): S = Q K^T with mma.sync.m16n8k16 (head_dim padded to a multiple of 16 with zeros, K fragments
// held in registers), softmax on the accumulator fragments (row max / sum across the 4 lanes of a quad), P (bf16) is
// re-used as the A fragment of O = P V whose B operand is read from the row-major V rows with ldmatrix.trans.
// FLOPs are 14 % of the block; the kernel streams 96 KB per window.
#include "common.cuh"

using namespace ldic;

namespace {

constexpr int kWaMaxTokens = 64;
constexpr int kWaMaxHeads = 16, kWaMaxTab = 15 * 15;      // (2 * 8 - 1)^2 table entries per head

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WaParams {
  const __nv_bfloat16 *q, *k, *v;      // [B*H*W][C] token images (q already scaled by head_dim^-0.5)
  __nv_bfloat16* out;                  // [B*H*W][C]
  const float* bias;                   // [heads][N][N] relative position bias (already gathered), or, with bias_table, the
                                       // module's relative_position_bias_table [(2ws-1)^2][heads] itself
  int B, H, W, C, heads, ws, shift;
  int bias_table;
};

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}

// HD = head_dim padded to a multiple of 16 (24 -> 32, 16 -> 16); N = ws*ws tokens (64 or 16).
// Shared memory: the window's q | k | v rows, [N][3C bf16 + 16 B] (the pad spreads the 8 rows of a fragment over the
// banks).  Per warp = head: K fragments are loaded once (registers) and reused by every 16-row query tile; V is read
// as the col-major B operand of O = P V straight from its row-major rows with ldmatrix.trans (no transposed copy);
// the output of a query tile is written over that tile's own q columns of this head, and the whole [N][C] result
// leaves with coalesced 16-byte stores.
template <int HD, int N>
__global__ void __launch_bounds__(256, 2) k_window_attention(WaParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int C = P.C, hd = C / P.heads;
  const int pitch = 3 * C * 2 + 16;                      // bytes per token row (q | k | v)
  uint8_t* s_qkv = smem;                                 // [N][pitch]
  __shared__ int s_pix[kWaMaxTokens];                    // token -> pixel index in the (B,H,W) image
  __shared__ int s_reg[kWaMaxTokens];                    // token -> region id of the shift mask
  // relative position bias table, [head][(2ws-1)^2]: 7.2 KB at 8 heads / 8x8 windows.  The gathered [heads][N][N] form is
  // 128 KB per window -- more than the window's q | k | v rows -- and does not stay in L1 next to two CTAs' shared memory:
  // its loads ran at L2 latency in front of every softmax (ncu: the kernel's top stall).  bias(i, j) =
  // table[(yi - yj + ws-1) * (2ws-1) + (xi - xj + ws-1)] (layers/win_attention.py:64-78, 101-104) is index arithmetic.
  __shared__ float s_tab[kWaMaxHeads * kWaMaxTab];
  const int ws_ = P.ws, Lt = (2 * ws_ - 1) * (2 * ws_ - 1);
  if (P.bias_table)
    for (int i = threadIdx.x; i < P.heads * Lt; i += blockDim.x) {
      const int hh = i / Lt, idx = i - hh * Lt;
      s_tab[i] = __ldg(P.bias + idx * P.heads + hh);
    }

  const int ws = P.ws, nwx = P.W / ws, nwy = P.H / ws;
  const int win = blockIdx.x;
  const int b = win / (nwx * nwy), wr = win % (nwx * nwy), wy = wr / nwx, wx = wr % nwx;
  if (threadIdx.x < N) {
    const int i = threadIdx.x / ws, j = threadIdx.x % ws;
    const int ys = wy * ws + i, xs = wx * ws + j;        // position in the shifted image
    const int y = (ys + P.shift) % P.H, x = (xs + P.shift) % P.W;      // torch.roll(x, -shift): shifted[p] = x[p + shift]
    s_pix[threadIdx.x] = (b * P.H + y) * P.W + x;
    int ry = 0, rx = 0;
    if (P.shift > 0) {
      ry = ys < P.H - ws ? 0 : (ys < P.H - P.shift ? 1 : 2);
      rx = xs < P.W - ws ? 0 : (xs < P.W - P.shift ? 1 : 2);
    }
    s_reg[threadIdx.x] = ry * 3 + rx;
  }
  __syncthreads();
  // only the windows that straddle the wrap-around of the cyclic shift (the last window row / column) hold tokens of
  // different regions: everywhere else the 0 / -100 mask is all zeros and its loads and compares are skipped
  bool masked = false;
  if (P.shift > 0) {
    const int ra = s_reg[0];
    masked = (s_reg[ws - 1] != ra) || (s_reg[N - ws] != ra) || (s_reg[N - 1] != ra);   // regions are monotone in i and j
  }
  // ---- stage q | k | v rows with cp.async (16-byte chunks, a warp per token row): every copy of the window is in
  // flight at once and nothing passes through registers.  (A flat loop with index divisions and a load -> store
  // dependency per iteration spent 42 % of the kernel's samples on its shared-memory store.)
  const int cpt = C / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_qkv);
#pragma unroll
  for (int t = warp; t < N; t += 8) {
    const long long pix = s_pix[t];
    const uint32_t drow = s_base + (uint32_t)(t * pitch);
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      const __nv_bfloat16* src = (which == 0 ? P.q : which == 1 ? P.k : P.v) + pix * C;
      for (int c8 = lane; c8 < cpt; c8 += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(drow + (uint32_t)((which * C + c8 * 8) * 2)), "l"(src + c8 * 8)
                     : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int h = warp; h < P.heads; h += blockDim.x / 32) {
    // ---- K fragments of this head (B operand of S = Q K^T), zero beyond head_dim ----
    uint32_t kf[N / 8][HD / 16][2];
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
      const uint8_t* krow = s_qkv + (size_t)(nt * 8 + g) * pitch + (C + h * hd) * 2;
#pragma unroll
      for (int kt = 0; kt < HD / 16; ++kt) {
        const int c0 = kt * 16 + t4 * 2, c1 = c0 + 8;
        kf[nt][kt][0] = c0 < hd ? *reinterpret_cast<const uint32_t*>(krow + c0 * 2) : 0u;
        kf[nt][kt][1] = c1 < hd ? *reinterpret_cast<const uint32_t*>(krow + c1 * 2) : 0u;
      }
    }
    const float* bias_h = P.bias + (size_t)h * N * N;
    // ldmatrix.trans row addresses of V: lane l (0..15) -> token (l & 15) of a 16-token k-tile, this head's dims
    const uint32_t v_lane = s_base + (uint32_t)((lane & 15) * pitch + (2 * C + h * hd) * 2);
#pragma unroll 1
    for (int mt = 0; mt < N / 16; ++mt) {                // 16 query rows at a time
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      uint32_t qa[HD / 16][4];
#pragma unroll
      for (int kt = 0; kt < HD / 16; ++kt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e & 1) ? r1 : r0, col = kt * 16 + t4 * 2 + ((e & 2) ? 8 : 0);
          qa[kt][e] = col < hd ? *reinterpret_cast<const uint32_t*>(s_qkv + (size_t)row * pitch + (h * hd + col) * 2) : 0u;
        }
      }
      float s[N / 8][4];
#pragma unroll
      for (int nt = 0; nt < N / 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < HD / 16; ++kt) mma_bf16_16816(s[nt], qa[kt], kf[nt][kt][0], kf[nt][kt][1]);
      }
      // + bias + mask, softmax over the N keys of rows r0 (elements 0,1) and r1 (elements 2,3)
      const int reg0 = s_reg[r0], reg1 = s_reg[r1];
      float m0 = -INFINITY, m1 = -INFINITY;
      const int wsh = ws == 8 ? 3 : 2, tw = 2 * ws - 1;
      const float* tab_h = s_tab + h * Lt;
      const int t0 = ((r0 >> wsh) + ws - 1) * tw + (r0 & (ws - 1)) + ws - 1;      // table index of (row r0, column 0)
      const int t1 = ((r1 >> wsh) + ws - 1) * tw + (r1 & (ws - 1)) + ws - 1;
#pragma unroll
      for (int nt = 0; nt < N / 8; ++nt) {
        const int col = nt * 8 + t4 * 2;
        float2 b0, b1;
        if (P.bias_table) {                              // column (ci, cj): index drops by ci * (2ws-1) + cj; col + 1 is cj + 1
          const int dc = (col >> wsh) * tw + (col & (ws - 1));
          b0.x = tab_h[t0 - dc]; b0.y = tab_h[t0 - dc - 1];
          b1.x = tab_h[t1 - dc]; b1.y = tab_h[t1 - dc - 1];
        } else {
          b0 = __ldg(reinterpret_cast<const float2*>(bias_h + r0 * N + col));
          b1 = __ldg(reinterpret_cast<const float2*>(bias_h + r1 * N + col));
        }
        float v0 = s[nt][0] + b0.x, v1 = s[nt][1] + b0.y, v2 = s[nt][2] + b1.x, v3 = s[nt][3] + b1.y;
        if (masked) {
          const int rc0 = s_reg[col], rc1 = s_reg[col + 1];
          if (rc0 != reg0) v0 += -100.f;
          if (rc1 != reg0) v1 += -100.f;
          if (rc0 != reg1) v2 += -100.f;
          if (rc1 != reg1) v3 += -100.f;
        }
        s[nt][0] = v0; s[nt][1] = v1; s[nt][2] = v2; s[nt][3] = v3;
        m0 = fmaxf(m0, fmaxf(v0, v1)); m1 = fmaxf(m1, fmaxf(v2, v3));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < N / 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = __expf(s[nt][e] - ((e & 2) ? m1 : m0));
          s[nt][e] = p;
          if (e & 2) l1 += p; else l0 += p;
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // O = P V : the S accumulator fragments of key tiles (2kt, 2kt+1) are the A fragment of k-tile kt; the B
      // fragment (V^T) comes from the row-major V rows through ldmatrix.trans.  Dim groups beyond head_dim are skipped.
      float o[HD / 8][4];
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < N / 16; ++kt) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
        pa[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
        pa[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
        for (int dt = 0; dt < HD / 8; ++dt) {
          if (dt * 8 < hd) {                             // warp-uniform
            uint32_t vb0, vb1;
            ldmatrix_x2_trans(vb0, vb1, v_lane + (uint32_t)(kt * 16 * pitch + dt * 16));
            mma_bf16_16816(o[dt], pa, vb0, vb1);
          }
        }
      }
      // normalise; rows r0, r1 of this head overwrite the tile's own q columns (no longer needed)
      const float i0 = 1.f / l0, i1 = 1.f / l1;
      uint8_t* d0 = s_qkv + (size_t)r0 * pitch + (h * hd) * 2;
      uint8_t* d1 = s_qkv + (size_t)r1 * pitch + (h * hd) * 2;
      __syncwarp();                                      // every lane has read its q fragments of this tile
#pragma unroll
      for (int dt = 0; dt < HD / 8; ++dt) {
        const int col = dt * 8 + t4 * 2;
        if (col < hd) {
          *reinterpret_cast<uint32_t*>(d0 + col * 2) = pack_bf16x2(o[dt][0] * i0, o[dt][1] * i0);
          *reinterpret_cast<uint32_t*>(d1 + col * 2) = pack_bf16x2(o[dt][2] * i1, o[dt][3] * i1);
        }
      }
    }
  }
  __syncthreads();
  // ---- window_reverse + reverse shift: the [N][C] result back to its pixels, 16 bytes per thread and store ----
#pragma unroll
  for (int t = warp; t < N; t += 8) {
    __nv_bfloat16* dst = P.out + (long long)s_pix[t] * C;
    for (int c8 = lane; c8 < cpt; c8 += 32)
      *reinterpret_cast<uint4*>(dst + c8 * 8) = *reinterpret_cast<const uint4*>(s_qkv + (size_t)t * pitch + c8 * 16);
  }
}

template <int HD, int N>
int launch_wa(const WaParams& P, cudaStream_t st) {
  const size_t smem = (size_t)N * (3 * P.C * 2 + 16);
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return fail(LDIC_ECUDA, "window attention: bad current device");
  std::lock_guard<std::mutex> init_lock(g_init_mu);
  static bool attr_set[kMaxDevices] = {};
  if (!attr_set[dev]) {
    LDIC_CUDA(cudaFuncSetAttribute(k_window_attention<HD, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev] = true;
  }
  if (smem > 200 * 1024) return fail(LDIC_EINVAL, "window attention: shared memory budget exceeded (%zu)", smem);
  const int nwin = P.B * (P.H / P.ws) * (P.W / P.ws);
  k_window_attention<HD, N><<<nwin, 256, smem, st>>>(P);
  return check_launch("k_window_attention");
}

// bias[h][i][j] = table[index[i][j]][h]   (layers/win_attention.py:101-104)
__global__ void k_gather_rel_bias(const float* __restrict__ table, const long long* __restrict__ index, float* __restrict__ bias,
                                  int heads, int NN) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < heads * NN) {
    const int h = i / NN, ij = i - h * NN;
    bias[i] = table[index[ij] * heads + h];
  }
}

// y (NCHW fp32) = shortcut (NCHW fp32) + o (NHWC fp32): the block's residual (layers/win_attention.py:204-205),
// 32x32 tiles through shared memory so both sides stay coalesced.
__global__ void __launch_bounds__(256) k_residual_nhwc_to_nchw(const float* __restrict__ o, const float* __restrict__ sc,
                                                               float* __restrict__ y, int C, long long HW, int Cp) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long p = p0 + r;
    const int c = c0 + tx;
    t[r][tx] = (p < HW && c < Cp) ? o[((long long)b * HW + p) * Cp + c] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long p = p0 + tx;
    if (c < C && p < HW) {
      const long long i = ((long long)b * C + c) * HW + p;
      y[i] = sc[i] + t[tx][r];
    }
  }
}


// y (NCHW fp32) = x (NCHW fp32) + a * sigmoid(b), a / b NHWC bf16: the gate + residual that closes Win_noShift_Attention
// (layers/layers.py:104-111) fused with the way back to the module surface's layout.
__global__ void __launch_bounds__(256) k_gate_residual_nhwc_to_nchw(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                                    const float* __restrict__ x, float* __restrict__ y, int C,
                                                                    long long HW, int Cp) {
  __shared__ float t[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const long long p = p0 + r;
    const int c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < Cp) {
      const long long i = ((long long)n * HW + p) * Cp + c;
      const float av = __bfloat162float(a[i]), bv = __bfloat162float(b[i]);
      v = av / (1.f + __expf(-bv));
    }
    t[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const long long p = p0 + tx;
    if (c < C && p < HW) {
      const long long i = ((long long)n * C + c) * HW + p;
      y[i] = x[i] + t[tx][r];
    }
  }
}

}  // namespace

extern "C" int ldic_window_attention_bias(const float* table, const long long* index, float* bias, int heads, int ws, void* stream) {
  if (!table || !index || !bias || heads <= 0 || ws <= 0) return fail(LDIC_EINVAL, "window attention bias: bad argument");
  const int NN = ws * ws * ws * ws;
  k_gather_rel_bias<<<(heads * NN + 255) / 256, 256, 0, (cudaStream_t)stream>>>(table, index, bias, heads, NN);
  return check_launch("k_gather_rel_bias");
}

static int window_attention_impl(const void* q, const void* k, const void* v, const float* bias, int bias_table, void* out,
                                 int B, int H, int W, int C, int heads, int ws, int shift, void* stream);
extern "C" int ldic_window_attention_core(const void* q, const void* k, const void* v, const float* bias, void* out, int B, int H,
                                          int W, int C, int heads, int ws, int shift, void* stream) {
  return window_attention_impl(q, k, v, bias, 0, out, B, H, W, C, heads, ws, shift, stream);
}
extern "C" int ldic_window_attention_core_table(const void* q, const void* k, const void* v, const float* table, void* out,
                                                int B, int H, int W, int C, int heads, int ws, int shift, void* stream) {
  return window_attention_impl(q, k, v, table, 1, out, B, H, W, C, heads, ws, shift, stream);
}
static int window_attention_impl(const void* q, const void* k, const void* v, const float* bias, int bias_table, void* out,
                                 int B, int H, int W, int C, int heads, int ws, int shift, void* stream) {
  if (!q || !k || !v || !bias || !out) return fail(LDIC_EINVAL, "window attention: null tensor");
  if (B <= 0) return LDIC_OK;
  if (heads <= 0 || C % heads || C % 8 || (ws != 8 && ws != 4) || H % ws || W % ws || shift < 0 || shift >= ws)
    return fail(LDIC_EINVAL, "window attention: need C %% heads == 0, C %% 8 == 0, window 8 or 4 dividing H and W, 0 <= shift < window");
  const int hd = C / heads;
  if (hd % 2 || hd > 32) return fail(LDIC_EINVAL, "window attention: head_dim must be even and <= 32 (got %d)", hd);
  if (heads > 16) return fail(LDIC_EINVAL, "window attention: at most 16 heads");
  WaParams P{(const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, bias, B, H, W, C, heads, ws, shift,
             bias_table};
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = hd > 16;
  if (ws == 8) return wide ? launch_wa<32, 64>(P, st) : launch_wa<16, 64>(P, st);
  return wide ? launch_wa<32, 16>(P, st) : launch_wa<16, 16>(P, st);
}

extern "C" int ldic_residual_nhwc_to_nchw_f32(const float* o_nhwc, const float* shortcut_nchw, float* y_nchw, int B, int C, int H,
                                              int W, int Cp, void* stream) {
  if (!o_nhwc || !shortcut_nchw || !y_nchw) return fail(LDIC_EINVAL, "residual: null tensor");
  if (B <= 0 || H <= 0 || W <= 0) return LDIC_OK;
  if (Cp < C) return fail(LDIC_EINVAL, "residual: Cp < C");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, B);
  k_residual_nhwc_to_nchw<<<grid, 256, 0, (cudaStream_t)stream>>>(o_nhwc, shortcut_nchw, y_nchw, C, HW, Cp);
  return check_launch("k_residual_nhwc_to_nchw");
}

extern "C" int ldic_gate_residual_nhwc_to_nchw_f32(const void* a_nhwc_bf16, const void* b_nhwc_bf16, const float* x_nchw, float* y_nchw,
                                                   int B, int C, int H, int W, int Cp, void* stream) {
  if (!a_nhwc_bf16 || !b_nhwc_bf16 || !x_nchw || !y_nchw) return fail(LDIC_EINVAL, "gate: null tensor");
  if (B <= 0 || H <= 0 || W <= 0) return LDIC_OK;
  if (Cp < C) return fail(LDIC_EINVAL, "gate: Cp < C");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, B);
  k_gate_residual_nhwc_to_nchw<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_nhwc_bf16, (const __nv_bfloat16*)b_nhwc_bf16,
                                                                       x_nchw, y_nchw, C, HW, Cp);
  return check_launch("k_gate_residual_nhwc_to_nchw");
}

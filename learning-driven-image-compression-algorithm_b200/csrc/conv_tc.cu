// Implicit-GEMM convolution / transposed convolution on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM), operands staged by TMA, with the GDN / IGDN
// normalisation fused into the epilogue as a second tensor-core contraction.
//
// One persistent CTA per SM, 384 threads:
//   warps 0..7   epilogue       (TMEM -> registers -> global), 2 warps per TMEM lane quadrant
//   warp 8       TMA producer   (activation tiles / regions; warp-uniform loop, one elected lane issues)
//   warp 9       MMA issuer     (one elected lane) + TMEM allocator
//   warp 10      weight / gamma tile producer (CTA-pair and halo kernels); first layer: patch builders (+ warp 11)
// (setmaxnreg moves registers from warpgroup 2 to the two epilogue warpgroups)
//
// Kernels: conv_tc2_kernel (CTA pairs, cta_group::2; W3 = wide-N form of the merged last deconv), conv_wide_kernel
// (384 output channels: single-buffered accumulator, gamma contraction tiled over its output channels),
// conv_first_kernel (first analysis layer straight from the NCHW fp32 or uint8 image).
//
// GEMM view: M = 128 output (or, for transposed convs, input-grid) pixels per tile,
// N = Np accumulator columns (= output channels, or 4 sub-pixel phases x channels for the
// merged small-Cout deconv), K = taps x Cin_pad walked in 64-channel blocks.
// A tile  : 128 pixels x 64 channels bf16, gathered by TMA from the NHWC activation
//           (zero fill outside the image = the reference's ZeroPad2d / conv padding);
//           stride-2 convs use one box with element strides (1,2,2,1).
// B tile  : Np x 64 slice of the packed weights [tap][Np][Cin_pad].
// GDN     : acc(+bias) -> x (registers) ; x^2 -> bf16 -> smem A-slot of the SAME pipeline
//           ring whose B-slot receives a 64-wide K block of gamma; norm = x^2 . gamma^T is
//           accumulated in place over the tile's accumulator; out = x * rsqrt|sqrt(norm + beta).
//
// Reference semantics: model/net.py:96-114 (g_a), :126-144 (g_s), :188-216 (h_a/h_s),
// model/gdn.py:69-92,134-156 (GDN/IGDN).
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include <memory>
#include <unordered_map>
#include <string.h>
#include "common.cuh"

using namespace ldic;

namespace {

#ifndef LDIC_EPI_WARPS
#define LDIC_EPI_WARPS 8                  // 8: two warps per TMEM lane quadrant (96 columns each at Np=192); 16: four (48 columns)
#endif
constexpr int kEpiWarps = LDIC_EPI_WARPS; // warps 0..kEpiWarps-1: epilogue (TMEM -> registers -> global)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 128;   // + one warpgroup: TMA producer, MMA issuer, B producer / patch builders
constexpr int kProdWarp = kEpiWarps;      // TMA producer (A tiles / A regions)
constexpr int kMmaWarp = kEpiWarps + 1;   // tcgen05.mma issuer + TMEM allocator
constexpr int kProdBWarp = kEpiWarps + 2; // halo kernels: weight / gamma tile producer; first layer: patch builders (+3)
constexpr int kEpiRegs = kEpiWarps == 8 ? 216 : 104;   // setmaxnreg of the epilogue warpgroups (the last warpgroup drops to 64)
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // bf16 elements = 128 B = one swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kMaxTaps = 112;
constexpr int kMaxJobs = 16;
constexpr int kTmemCols = 512;
constexpr int kBufCols = 256;           // two TMEM accumulator buffers at columns 0 and 256
#ifndef LDIC_PAIR_STORE
#define LDIC_PAIR_STORE 1                 // build switch: lane-pair transposed epilogue stores (0 = one 32-byte chunk per lane)
#endif
constexpr int kGdnInsertDefault = 4;     // conv stages of the next tile issued before the previous tile's GDN stages

// one filter tap = one spatial offset of the gather + a K range: nkc 64-wide blocks starting at
// channel a_c0 of the activation and column b_c0 of the tap's packed weight block
struct Tap { short dx, dy, px, nkc; int a_c0, b_c0; };
struct Job { int ntaps, tap_begin, nkb, oy_off, ox_off, out_off; };

struct ConvParams {
  int mode;                 // 0: stride-1 gather, 2: stride-2 gather (one TMA box with element strides (1,2,2,1))
  int TW, TH, TN;
  int tw_shift, th_shift, cg_shift;
  int tiles_x, tiles_y, tiles_n, tiles_per_job, njobs, total_tiles;
  int Wg, Hg, B;            // pixel grid the M tiles walk over
  int gdn_kblocks;          // Np / 64 when act is GDN/IGDN, else 0
  int act, out_f32;
  int stages;
  int SA, SB;               // conv_wide_kernel: slots of the activation / weight rings
  int x_ovl;                // tiles overlap by x_ovl pixels in x (wide-N merged deconv: 2 = one halo pixel per side), else 0
  int tma_store;            // first layer: the bf16 output tile leaves through shared memory and TMA stores
  int gdn_insert;           // streaming kernels: conv stages of tile it+1 issued before the GDN stages of tile it
  int super_per_job;        // CTA-pair kernel: (tiles_per_job + 1) / 2 pairs of adjacent tiles per job
  int ngroups, Cg;          // accumulator columns = ngroups x Cg
  int sy, sx;               // output pixel = grid pixel * (sy,sx) + job offset (+ sub-pixel group offset)
  int nbias;                // 1, or njobs when every job has its own bias vector
  long long out_sN, out_sY, out_sX;  // output strides in elements
  Job jobs[kMaxJobs];
  Tap taps[kMaxTaps];
  const float* bias;
  const float* beta;
  void* out;
  unsigned long long* dbg;   // optional cycle counters of CTA 0 (LDIC_DEBUG_TIMING=1), else null
  int tf_round_out;          // TF32 mode: round the fp32 outputs to tf32 (layers that feed another tf32 layer)
  const __nv_bfloat16* residual;   // optional NHWC bf16 tensor of the output's shape added after the activation (no GDN)
  int dbg_nostore;           // experiments: bit 0 skips the epilogue's global stores (LDIC_DEBUG_NOSTORE=1), bit 1 the first
                             // layer's patch assembly (ldic_set_tuning("debug_nostore", 2)): which role bounds the kernel?
  // fused tail of Net.forward on the merged last deconv: per-image 1x1 conv (batch_conv) + 8-bit-level squared error
  const void* tail_x;        // NCHW input image [B,3,tail_H,tail_W] (fp32 in [-1,1], or uint8 levels when tail_u8), or null
  int tail_u8;
  int tail_tanh;             // x~ = tanh(batch_conv(...)) before the level error (U-Net family, model/net_unet_ha_hs.py:980)
  const float* tail_w;       // [B][3][Cg] per-image filters
  float* tail_xo;            // optional NCHW fp32 reconstruction
  unsigned long long* tail_sq;  // [B] exact sums (accumulated into)
  int tail_H, tail_W;
};

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.  Before trapping, the waiter
// records {block, thread, barrier offset in dynamic smem, parity} in host-mapped memory so that the host can
// still read which barrier starved after the context is lost (ldic_debug_last_timeout).
__device__ unsigned long long* g_timeout_report = nullptr;
constexpr long long kWatchdogCycles = 60000000000LL;   // ~30 s at 2 GHz: a protocol bug, not a time-sliced or profiled GPU
__device__ __forceinline__ void mbar_timeout(uint32_t bar_addr, uint32_t parity) {
  extern __shared__ uint8_t smem_raw[];
  unsigned long long* r = g_timeout_report;
  if (r && atomicCAS(r, 0ull, 1ull) == 0ull) {
    r[1] = blockIdx.x; r[2] = threadIdx.x; r[3] = bar_addr - smem_u32(smem_raw); r[4] = parity;
    __threadfence_system();
  }
  printf("ldic conv_tc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar_addr, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > kWatchdogCycles) mbar_timeout(smem_u32(bar), parity);
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// TMA store of a shared-memory box (same 128-byte swizzle as the operand tiles) to global memory; bulk-group completion
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {     // true on exactly one lane of a converged warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T, bf16 inputs, fp32 accumulate, M=128, N from idesc, K=16.  The two smem descriptors
// (K-major, 128-byte swizzled operand tiles: rows of 128 B, 8-row groups `SBO` bytes apart) are given as 32-bit halves
// (lo = start address >> 4 | LBO, hi = SBO >> 4 | descriptor version 1 << 14 | SWIZZLE_128B 2 << 29): the issuing thread
// only does one 32-bit add per operand per K step.
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ void st_global_v8(void* p, const T* v) {   // 32-byte store, p 32-byte aligned
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(u[0]), "r"(u[1]), "r"(u[2]),
               "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}

__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}


// ---- thread-block cluster / CTA-pair (cta_group::2) wrappers -----------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on a barrier of any CTA of the cluster.  Default semantics (.release at .cta scope, the form CUTLASS's
// ClusterBarrier::arrive uses): what the waiter consumes is shared-memory operand data already made visible to the
// async proxy by fence.proxy.async, and TMEM reads ordered by tcgen05.fence::before_thread_sync.  The
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, which also waits for every global store of the thread
// to be acknowledged by L2: 17 % of all warp stall samples of the deconv-3 launch (profiles/r01_ncu_d3pair_stalls.txt).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {  // acquire at cluster scope
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cl(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > kWatchdogCycles) mbar_timeout(smem_u32(bar), parity);
  }
}
// TMA loads of a CTA pair: the data lands in the executing CTA, the completion may be signalled on a barrier of
// either CTA of the pair (mbar = shared::cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t mbar_cl, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(mbar_cl), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t mbar_cl, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(mbar_cl), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA of a CTA pair: each CTA contributes 128 rows of A and half of the B rows; issued by the leader only
__device__ __forceinline__ void umma_bf16_lh_cg2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::tf32 form of the same (fp32 operands in shared memory, K = 8 per instruction = the same 32 bytes per row)
__device__ __forceinline__ void umma_tf32_lh_cg2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                 uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {        // round to nearest (the tensor core would truncate)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// completion of the pair's MMAs arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// ---------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------
struct TileCoord { int job, n0, y0, x0; };

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& P, int t) {
  TileCoord c;
  c.job = t / P.tiles_per_job;
  int r = t - c.job * P.tiles_per_job;
  int per_n = P.tiles_y * P.tiles_x;
  int tn = r / per_n;
  r -= tn * per_n;
  int ty = r / P.tiles_x;
  int tx = r - ty * P.tiles_x;
  c.n0 = tn * P.TN;
  c.y0 = ty * P.TH;
  c.x0 = tx * (P.TW - P.x_ovl) - (P.x_ovl >> 1);
  return c;
}

// CTA-pair kernels walk "super tiles": super tile s of a job = tiles 2r and 2r+1 of that job, one per CTA of the pair.
// A tile index past the end of the job (odd tile count) decodes to image index B: every TMA read of it is
// zero-filled and every store masked.
__device__ __forceinline__ TileCoord decode_tile2(const ConvParams& P, int s, int rank) {
  const int job = s / P.super_per_job;
  const int t = 2 * (s - job * P.super_per_job) + rank;
  if (t >= P.tiles_per_job) { TileCoord c; c.job = job; c.n0 = P.B; c.y0 = 0; c.x0 = 0; return c; }
  return decode_tile(P, job * P.tiles_per_job + t);
}

// ---------------------------------------------------------------------------------
// Epilogue role (shared by the streaming and the halo kernels): warps 4..11.
// ---------------------------------------------------------------------------------
struct EpiRing {
  uint32_t ring_base;       // smem address of slot 0 of the ring that carries the x^2 operand tiles
  uint32_t slot_bytes;
  uint32_t nslots;
  uint64_t* empty_bar;      // [nslots] "slot free" barriers of that ring
  uint64_t *acc_full, *buf_free, *x2_ready, *norm_full;   // [2] each
  const float *s_bias, *s_beta;
  int insert_after;         // GDN items of tile it-1 ride after this many stream items of tile it
  // CTA-pair kernels: tiles are super tiles t_first + it * t_stride decoded for `rank`; buf_free / x2_ready live in
  // the leader CTA and are reached through their shared::cluster addresses
  int t_first, t_stride, rank;
  uint32_t buf_free_cl, x2_ready_cl;
  int pair_store;           // lane-pair transposed bf16 stores (off for the first layer: measured 6 % slower there)
  // first layer: the normalised bf16 tile is written over the tile's own x^2 operand slots (same rows, same swizzle) and
  // leaves with one TMA store per 64-channel block instead of 32-byte-per-lane global stores
  const CUtensorMap* tma_out;
  int epi_threads;
};

// Fused tail on the CPT accumulator columns of one thread (CPT / CG output pixels of CG channels each); every index
// into xr is a compile-time constant so the accumulators stay in registers.
// a11 on one sample of the fused tail: the image is fp32 in [-1,1] (model/net.py:864) or its uint8 levels -- for an
// 8-bit image x = (u/255)*2-1 (eval_net.py:84 after ToTensor) the reference's gt = round((x+1)*127.5) is u itself
// (checked exhaustively for u = 0..255 in tests/test_oracle_golden.py), so the level is used directly.
__device__ __forceinline__ unsigned int tail_sq_err(const ConvParams& P, long long idx, float xt) {
  if (P.tail_u8) {
    const float gt = (float)__ldg(reinterpret_cast<const unsigned char*>(P.tail_x) + idx);
    float xh = __fmul_rn(__fadd_rn(xt, 1.f), 127.5f);
    xh = rintf(fminf(fmaxf(xh, 0.f), 255.f));
    const float d = xh - gt;
    return (unsigned int)(d * d);
  }
  return sq_level_err(__ldg(reinterpret_cast<const float*>(P.tail_x) + idx), xt, 0);
}

template <int CPT, int CG>
__device__ __forceinline__ unsigned long long fused_tail_pixels(const ConvParams& P, const float (&xr)[CPT], int col0, int n,
                                                                int oy0, int ox0) {
  unsigned long long acc = 0;
  if constexpr (CPT % CG == 0) {
    const float* wn = P.tail_w + (long long)n * 3 * CG;
#pragma unroll
    for (int jj = 0; jj < CPT; jj += CG) {
      const int g = (col0 + jj) / CG;
      const int oy = oy0 + (g >> 1), ox = ox0 + (g & 1);
      float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < CG; ++m) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = fmaf(xr[jj + m], __ldg(wn + c * CG + m), o[c]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long idx = (((long long)n * 3 + c) * P.tail_H + oy) * P.tail_W + ox;
        if (P.tail_tanh) o[c] = tanhf(o[c]);
        if (P.tail_xo) P.tail_xo[idx] = o[c];
        acc += tail_sq_err(P, idx, o[c]);
      }
    }
  }
  return acc;
}

// EW = number of epilogue warps (EW / 4 per TMEM lane quadrant, NP / (EW / 4) accumulator columns per thread): 8 in the
// streaming kernels; the first layer, which IS its epilogue (5 conv MMAs per tile), runs 12 at NP = 192.
template <int NP, bool CL = false, int EW = kEpiWarps, bool RES = false, bool TF = false>
__device__ __forceinline__ void epilogue_role(const ConvParams& P, const EpiRing& R, uint32_t tmem_base, int gk,
                                              int ntiles_cta, int warp, int lane) {
  static_assert(EW % 4 == 0 && NP % (EW / 4) == 0 && (NP / (EW / 4)) % 16 == 0, "epilogue warps must split the columns in 16s");
  constexpr int CPT = NP / (EW / 4);             // accumulator columns per thread
  constexpr int LDW = (EW == 8 && CPT % 32 == 0) ? 32 : 16;   // columns per tcgen05.ld (16: register budget of 12 / 16 warps)
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int h = warp >> 2;                // column slice (EW / 4 slices of CPT columns)
    const int r = q * 32 + lane;            // tile row = TMEM lane
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int col0 = h * CPT;
    // row -> pixel of the M grid (TW, TH are powers of two)
    const int xi = r & (P.TW - 1), yi = (r >> P.tw_shift) & (P.TH - 1), ni = r >> (P.tw_shift + P.th_shift);
    const bool igdn = (P.act == LDIC_ACT_IGDN);
    uint32_t sbase = 0;                     // ring position of the first stage of tile `it`'s stream
    const bool edbg = P.dbg != nullptr && blockIdx.x == 0 && warp == 0;
    long long e_acc = 0, e_norm = 0, e_slot = 0, e_t0 = 0, e_p1 = 0, e_p2 = 0, e_t1 = 0;
    const long long e_begin = edbg ? clock64() : 0;
    // tile coordinates are decoded two tiles ahead, in the shadow of the wait for the gamma contraction (five integer
    // divisions: 6 % of the epilogue warps' samples when done at the top of the loop)
    auto decode_it = [&](int i) {
      return CL ? decode_tile2(P, R.t_first + i * R.t_stride, R.rank) : decode_tile(P, blockIdx.x + i * gridDim.x);
    };
    TileCoord tc = decode_it(0), tc_n1 = decode_it(1), tc_n2 = tc_n1;
    for (int it = 0; it < ntiles_cta; ++it, tc = tc_n1, tc_n1 = tc_n2) {
      const Job jb = P.jobs[tc.job];
      const int bsel = it & 1;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      const uint32_t tbuf = tmem_base + bsel * kBufCols + lane_sel + col0;
      const int gx_ = tc.x0 + xi, gy_ = tc.y0 + yi, gn_ = tc.n0 + ni;
      const bool valid = (gx_ < P.Wg) && (gy_ < P.Hg) && (gn_ < P.B) && !(P.dbg_nostore & 1);
      const long long pix_base = (long long)gn_ * P.out_sN + (long long)(gy_ * P.sy + jb.oy_off) * P.out_sY +
                                 (long long)(gx_ * P.sx + jb.ox_off) * P.out_sX + jb.out_off;
      // ring position of this tile's gamma items: after the first min(insert_after, len) items of the
      // NEXT tile's stream, or right after this tile's stream when it is the CTA's last tile
      const int len_this = jb.nkb;
      sbase += (uint32_t)len_this + ((gk && it > 0) ? (uint32_t)gk : 0u);
      uint32_t gpos = sbase;
      if (gk && it + 1 < ntiles_cta) {
        const Job jn = P.jobs[tc_n1.job];
        const int len_next = jn.nkb;
        gpos += (uint32_t)(len_next < R.insert_after ? len_next : R.insert_after);
      }

      if (edbg) e_t0 = clock64();
      mbar_wait(&R.acc_full[bsel], par);
      if (edbg) { e_t1 = clock64(); e_acc += e_t1 - e_t0; }
      tc_fence_after();

      // ---- pass 1: accumulator -> registers (+bias) ----
      const float* sb = R.s_bias + (P.nbias > 1 ? tc.job * NP : 0);
      float xr[CPT];
      {
        uint32_t(&xu)[CPT] = reinterpret_cast<uint32_t(&)[CPT]>(xr);
#pragma unroll
        for (int c = 0; c < CPT; c += LDW) tmem_ldn<LDW>(tbuf + c, reinterpret_cast<uint32_t(&)[LDW]>(xu[c]));
        tmem_ld_wait();
      }
      tc_fence_before();
      if (!gk) {                                 // no GDN: the buffer can take the tile after next right away
        if (CL) mbar_arrive_cluster(R.buf_free_cl + 8u * bsel); else mbar_arrive(&R.buf_free[bsel]);
        tc_n2 = decode_it(it + 2);
      }
#pragma unroll
      for (int c = 0; c < CPT; c += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(&sb[col0 + c]);
        xr[c] += b4.x; xr[c + 1] += b4.y; xr[c + 2] += b4.z; xr[c + 3] += b4.w;
      }

      // stores columns [c, c+n) of this thread's pixel (ngroups == 1: contiguous channels): whole 32-byte sectors
      bool stored = false;
      auto store_cols = [&](int c, int n) {
        if (P.out_f32) {
          float* dst = reinterpret_cast<float*>(P.out) + pix_base + col0 + c;
          if constexpr (TF) {
            // an activation that feeds another kind::tf32 layer is rounded to tf32 here (round to nearest): the tensor
            // core would truncate the fp32 value it reads, a one-sided error that adds up coherently over K
            if (P.tf_round_out) {
#pragma unroll
              for (int k = 0; k < LDW; ++k) if (k < n) xr[c + k] = __uint_as_float(to_tf32(xr[c + k]));
            }
          }
#pragma unroll
          for (int j = 0; j < LDW / 8; ++j) if (8 * j < n) st_global_v8(dst + 8 * j, &xr[c + 8 * j]);
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + pix_base + col0 + c;
#pragma unroll
          for (int j = 0; j < LDW / 16; ++j) {
            if (16 * j < n) {
              const float* x16 = &xr[c + j * 16];
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = pack_bf16x2(x16[2 * e], x16[2 * e + 1]);
              st_global_v8(dst + 16 * j, pk);
            }
          }
        }
      };
      // bf16 output, 32-column chunks: the two lanes of a pair swap half a chunk so that every store instruction
      // writes 64 contiguous bytes per pixel with two lanes (16 distinct 128-byte lines per warp instruction instead
      // of 32 -- the LSU cost of these stores is per line, LDIC_DEBUG_NOSTORE: 2.7 k of deconv 3's 11 k cycles per tile)
      const bool pair_store = (LDW == 32) && !P.out_f32 && P.ngroups == 1 && (LDIC_PAIR_STORE != 0) && R.pair_store;
      long long pb_other = 0;
      bool valid_other = false;
      if (pair_store) {
        pb_other = __shfl_xor_sync(0xffffffffu, pix_base, 1);
        valid_other = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
      }
      auto store_cols_paired = [&](int c) {          // all lanes call it (shuffles); stores are predicated
        uint32_t pk[16], rv[8];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(xr[c + 2 * e], xr[c + 2 * e + 1]);
        const bool odd = lane & 1;
#pragma unroll
        for (int e = 0; e < 8; ++e) rv[e] = __shfl_xor_sync(0xffffffffu, odd ? pk[e] : pk[8 + e], 1);
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(P.out) + col0 + c + (odd ? 16 : 0);
        // even lane's pixel: even lane writes its columns [c, c+16), odd lane the even lane's [c+16, c+32)
        uint32_t d1[8], d2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { d1[e] = odd ? rv[e] : pk[e]; d2[e] = odd ? pk[8 + e] : rv[e]; }
        if (odd ? valid_other : valid) st_global_v8(out + (odd ? pb_other : pix_base), d1);
        // odd lane's pixel: even lane writes the odd lane's [c, c+16), odd lane its own [c+16, c+32)
        if (odd ? valid : valid_other) st_global_v8(out + (odd ? pix_base : pb_other), d2);
      };
      if (gk) {
        // x^2 -> bf16 -> A slots (K-major, 128B swizzle: 16-byte chunk index XOR (row & 7))
        if (edbg) { e_t0 = clock64(); e_p1 += e_t0 - e_t1; }
        {
          // the gk operand slots are released in ring order by the MMA warp's commits: waiting for the LAST one
          // covers all of them (one barrier round trip instead of gk)
          const uint32_t kc2 = gpos + (uint32_t)gk - 1u;
          mbar_wait(&R.empty_bar[kc2 % R.nslots], ((kc2 / R.nslots) & 1) ^ 1);
        }
        if (edbg) { e_t1 = clock64(); e_slot += e_t1 - e_t0; }
        const uint32_t row_off = (uint32_t)r * 128u, rx = (uint32_t)(r & 7);
        if (R.tma_out && it > 0) {
          // the slots still feed the TMA store of the previous tile's output: its issuer waits for the reads, then everyone
          if (warp == 0 && lane == 0) tma_store_wait_read();
          asm volatile("bar.sync 2, %0;" ::"r"(R.epi_threads) : "memory");
        }
        if constexpr (TF) {
          // TF32 parity mode: the x^2 operand stays fp32 (rounded to tf32): 32 columns per 128-byte row, NP / 32 slots
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) {
            const int col = col0 + j * 4;
            const uint32_t kc2 = gpos + (col >> 5);
            const uint32_t a_addr = R.ring_base + (kc2 % R.nslots) * R.slot_bytes;
            const uint32_t chunk = (uint32_t)((col & 31) >> 2);
            const float* x4 = &xr[j * 4];
            st_shared_v4(a_addr + row_off + ((chunk ^ rx) << 4), to_tf32(x4[0] * x4[0]), to_tf32(x4[1] * x4[1]),
                         to_tf32(x4[2] * x4[2]), to_tf32(x4[3] * x4[3]));
          }
        } else {
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j) {
          const int col = col0 + j * 8;
          const uint32_t kc2 = gpos + (col >> 6);
          const uint32_t a_addr = R.ring_base + (kc2 % R.nslots) * R.slot_bytes;
          const uint32_t chunk = (uint32_t)((col & 63) >> 3);
          const float* x8 = &xr[j * 8];
          st_shared_v4(a_addr + row_off + ((chunk ^ rx) << 4), pack_bf16x2(x8[0] * x8[0], x8[1] * x8[1]),
                       pack_bf16x2(x8[2] * x8[2], x8[3] * x8[3]), pack_bf16x2(x8[4] * x8[4], x8[5] * x8[5]),
                       pack_bf16x2(x8[6] * x8[6], x8[7] * x8[7]));
        }
        }
        fence_async_smem();                  // generic-proxy writes -> visible to the tensor-core (async) proxy
        if (CL) mbar_arrive_cluster(R.x2_ready_cl + 8u * bsel); else mbar_arrive(&R.x2_ready[bsel]);
        if (edbg) { e_t0 = clock64(); e_p1 += e_t0 - e_t1; }
        tc_n2 = decode_it(it + 2);
        mbar_wait(&R.norm_full[bsel], par);
        if (edbg) { e_t1 = clock64(); e_norm += e_t1 - e_t0; }
        tc_fence_after();
        // ---- pass 2: out = x * rsqrt(norm + beta)   (IGDN: x * sqrt = x * n * rsqrt(n)) ----
        // The TMEM read of chunk c+1 is in flight while chunk c is normalised, and every finished chunk is
        // stored at once (its stores drain while the next chunks are computed).
        constexpr int NCH = CPT / LDW;
        uint32_t tr[2][LDW];
        tmem_ldn<LDW>(tbuf, tr[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c = ch * LDW;
          float bb[LDW];                         // this chunk's beta: the shared loads are in flight during the TMEM wait
#pragma unroll
          for (int k = 0; k < LDW; k += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&R.s_beta[col0 + c + k]);
            bb[k] = b4.x; bb[k + 1] = b4.y; bb[k + 2] = b4.z; bb[k + 3] = b4.w;
          }
          tmem_ld_wait();
          if (ch + 1 < NCH) {
            tmem_ldn<LDW>(tbuf + c + LDW, tr[(ch + 1) & 1]);
          } else {
            // the last norm chunk is in registers: the accumulator buffer is free for tile it+2 now, not after the
            // remaining arithmetic and stores
            tc_fence_before();
            if (CL) mbar_arrive_cluster(R.buf_free_cl + 8u * bsel); else mbar_arrive(&R.buf_free[bsel]);
          }
#pragma unroll
          for (int k = 0; k < LDW; ++k) {
            const float nrm = __uint_as_float(tr[ch & 1][k]) + bb[k];
            const float rs = rsqrt_approx(nrm);
            xr[c + k] *= igdn ? nrm * rs : rs;
          }
          if (R.tma_out) {
#pragma unroll
            for (int j = 0; j < LDW / 8; ++j) {
              const int col = col0 + c + j * 8;
              const uint32_t a_addr = R.ring_base + ((gpos + (uint32_t)(col >> 6)) % R.nslots) * R.slot_bytes;
              const float* x8 = &xr[c + j * 8];
              st_shared_v4(a_addr + row_off + (((uint32_t)((col & 63) >> 3) ^ rx) << 4), pack_bf16x2(x8[0], x8[1]),
                           pack_bf16x2(x8[2], x8[3]), pack_bf16x2(x8[4], x8[5]), pack_bf16x2(x8[6], x8[7]));
            }
          } else if (pair_store) store_cols_paired(c);
          else if (P.ngroups == 1 && valid) store_cols(c, LDW);
        }
        if (R.tma_out) {
          fence_async_smem();                // generic-proxy writes -> visible to the TMA (async proxy)
          asm volatile("bar.sync 2, %0;" ::"r"(R.epi_threads) : "memory");
          if (warp == 0 && lane == 0 && !(P.dbg_nostore & 1)) {
            for (int kb = 0; kb < gk; ++kb)
              tma_store_4d(R.tma_out, R.ring_base + ((gpos + (uint32_t)kb) % R.nslots) * R.slot_bytes, kb * 64, tc.x0, tc.y0, tc.n0);
            tma_store_commit();
          }
        }
        stored = (P.ngroups == 1);
      } else if (P.act == LDIC_ACT_RELU) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) xr[c] = fmaxf(xr[c], 0.f);
      } else if (P.act == LDIC_ACT_LEAKY02 || P.act == LDIC_ACT_LEAKY001) {
        const float slope = P.act == LDIC_ACT_LEAKY02 ? 0.2f : 0.01f;
#pragma unroll
        for (int c = 0; c < CPT; ++c) xr[c] = xr[c] > 0.f ? xr[c] : slope * xr[c];
      }
      if constexpr (RES) {
        // residual connection fused in: out = act(conv(x) + bias) + r, r an NHWC bf16 tensor of the output's shape
        // (ResidualBlock: layers/layers.py:87-102 via CompressAI; WinBasedAttention: layers/win_attention.py:204-205)
        if (valid && P.residual) {
          const __nv_bfloat16* rp = P.residual + pix_base + col0;
#pragma unroll
          for (int c = 0; c < CPT; c += 16) {
            uint32_t u[8];
            asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                         : "l"(rp + c));
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              xr[c + 2 * e] += __uint_as_float(u[e] << 16);
              xr[c + 2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u);
            }
          }
        }
      }

      unsigned long long tail_acc = 0;       // fused tail: this thread's squared level error on this tile
      if (valid) {
        // 256-bit stores: every lane writes whole 32-byte sectors of its own pixel row
        if (P.ngroups == 1) {
          // this thread's CPT columns are contiguous channels of one output pixel
          if (!stored) {
#pragma unroll
            for (int c = 0; c < CPT; c += LDW) store_cols(c, LDW);
          }
        } else {
          // merged sub-pixel phases: Cg (a power of two >= 8) channels per output pixel, group g = col / Cg
          if (P.out) {
#pragma unroll
            for (int j = 0; j < CPT / 8; ++j) {
              const int col = col0 + j * 8;
              const int g = col >> P.cg_shift, cc = col & (P.Cg - 1);
              const long long off = pix_base + (long long)(g >> 1) * P.out_sY + (long long)(g & 1) * P.out_sX + cc;
              const float* x8 = &xr[j * 8];
              if (P.out_f32) {
                st_global_v8(reinterpret_cast<float*>(P.out) + off, x8);
              } else {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.out) + off);
                *dst = make_uint4(pack_bf16x2(x8[0], x8[1]), pack_bf16x2(x8[2], x8[3]), pack_bf16x2(x8[4], x8[5]),
                                  pack_bf16x2(x8[6], x8[7]));
              }
            }
          }
          if constexpr (NP <= 128) if (P.tail_x) {   // merged deconv: Np = 4 x Cg = 64 or 128 accumulator columns
            // batch_conv (model/net.py:527-537, :811): x~[c] = sum_m w[n][c][m] * v[m] on this thread's output pixels,
            // then the a11 squared level error against the input image (model/net.py:864-868)
            if (P.Cg == 16) tail_acc = fused_tail_pixels<CPT, 16>(P, xr, col0, gn_, gy_ * P.sy + jb.oy_off, gx_ * P.sx + jb.ox_off);
            else if (P.Cg == 32) tail_acc = fused_tail_pixels<CPT, 32>(P, xr, col0, gn_, gy_ * P.sy + jb.oy_off, gx_ * P.sx + jb.ox_off);
          }
        }
      }
      if (edbg) e_p2 += clock64() - e_t1;
      if constexpr (NP <= 128) if (P.tail_x) {
        // exact integer sums: the order of the atomic adds does not matter.  One image per tile (TN == 1): one
        // warp-level sum and one atomic per warp; tiles spanning images: one atomic per contributing thread.
        if (P.TN == 1) {
          unsigned long long v = tail_acc;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && v) atomicAdd(P.tail_sq + tc.n0, v);
        } else if (tail_acc) {
          atomicAdd(P.tail_sq + gn_, tail_acc);
        }
      }
    }
    if (R.tma_out && warp == 0 && lane == 0) tma_store_wait_all();      // the last tile's output has left shared memory
    if (edbg && lane == 0) {
      P.dbg[16] = (unsigned long long)(clock64() - e_begin); P.dbg[17] = (unsigned long long)e_acc;
      P.dbg[18] = (unsigned long long)e_norm; P.dbg[19] = (unsigned long long)e_slot; P.dbg[20] = (unsigned long long)ntiles_cta;
      P.dbg[21] = (unsigned long long)e_p1; P.dbg[22] = (unsigned long long)e_p2;
    }
}

// ---------------------------------------------------------------------------------
// Epilogue of the wide-N merged last deconv (CTA-pair streaming kernel, W3 = true).
// The accumulator holds NP = 3 * NPO columns: block dxi = columns [dxi*NPO, +NPO) is
//   D[p][dxi] = sum_dy sum_k X[p + (dy, 0)][k] * W[(dy, dxi - 1)][k][.]        (NPO = 4 sub-pixel phases x Cg channels)
// i.e. the three dx taps are carried side by side in N instead of as separate N = NPO MMAs (an SS-mode MMA costs
// ~55 cycles at N = 64 but 96 at N = 192: 36 MMAs per tile instead of 108).  The transposed conv output of pixel p is
//   acc[p] = D[p-1][0] + D[p][1] + D[p+1][2],
// and with TW = 32 a warp is one row of the tile (lane = x), so the neighbours are lanes -1 / +1: two shuffles per
// column.  Lanes 0 and 31 are halo pixels (tiles overlap by 2 in x), they contribute but are not written.
// Then the usual IGDN (x^2 tile -> gamma contraction with N = NPO -> x * sqrt(norm)), optional NHWC store of the
// 4 x Cg sub-pixel outputs, and the fused tail (batch_conv + squared level error).
// ---------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void epilogue_w3(const ConvParams& P, const EpiRing& R, uint32_t tmem_base, int gk,
                                            int ntiles_cta, int warp, int lane) {
  constexpr int NPO = NP / 3;
  constexpr int CPT = NPO / (kEpiWarps / 4);             // output columns per thread (32)
  static_assert(CPT == 32, "wide-N merged deconv: 64 logical columns, 8 epilogue warps");
  const int q = warp & 3, h = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const int col0 = h * CPT;
  const int xi = r & (P.TW - 1), yi = (r >> P.tw_shift) & (P.TH - 1), ni = r >> (P.tw_shift + P.th_shift);
  uint32_t sbase = 0;
  const bool edbg = P.dbg != nullptr && blockIdx.x == 0 && warp == 0;
  long long e_acc = 0, e_norm = 0, e_slot = 0, e_t0 = 0, e_p1 = 0, e_p2 = 0, e_t1 = 0;
  const long long e_begin = edbg ? clock64() : 0;
  float wreg[48];                        // the current image's batch_conv filter (3 x 16), reloaded when the image changes
  int w_n = -1;
  auto decode_it = [&](int i) { return decode_tile2(P, R.t_first + i * R.t_stride, R.rank); };
  TileCoord tc = decode_it(0), tc_n1 = decode_it(1), tc_n2 = tc_n1;
  for (int it = 0; it < ntiles_cta; ++it, tc = tc_n1, tc_n1 = tc_n2) {
    const Job jb = P.jobs[tc.job];
    const int bsel = it & 1;
    const uint32_t par = (uint32_t)(it >> 1) & 1u;
    const uint32_t tbuf = tmem_base + bsel * kBufCols + lane_sel;
    const int gx_ = tc.x0 + xi, gy_ = tc.y0 + yi, gn_ = tc.n0 + ni;
    const bool valid = xi >= 1 && xi <= P.TW - 2 && gx_ < P.Wg && gy_ < P.Hg && gn_ < P.B && !(P.dbg_nostore & 1);
    const long long pix_base = (long long)gn_ * P.out_sN + (long long)(gy_ * P.sy + jb.oy_off) * P.out_sY +
                               (long long)(gx_ * P.sx + jb.ox_off) * P.out_sX + jb.out_off;
    sbase += (uint32_t)jb.nkb + (it > 0 ? (uint32_t)gk : 0u);
    uint32_t gpos = sbase;
    if (it + 1 < ntiles_cta) {
      const int len_next = P.jobs[tc_n1.job].nkb;
      gpos += (uint32_t)(len_next < R.insert_after ? len_next : R.insert_after);
    }
    if (edbg) e_t0 = clock64();
    mbar_wait(&R.acc_full[bsel], par);
    if (edbg) { e_t1 = clock64(); e_acc += e_t1 - e_t0; }
    tc_fence_after();
    // ---- pass 1: the three dx blocks of this thread's 32 columns, combined across the x neighbours ----
    float xr[CPT];
    {
      uint32_t d0[CPT], d1[CPT], d2[CPT];
      tmem_ldn<32>(tbuf + 0 * NPO + col0, d0);
      tmem_ldn<32>(tbuf + 1 * NPO + col0, d1);
      tmem_ldn<32>(tbuf + 2 * NPO + col0, d2);
      tmem_ld_wait();
      const float* sb = R.s_bias;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[c]), 1);       // D[p-1][0]
        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[c]), 1);     // D[p+1][2]
        xr[c] = (up + __uint_as_float(d1[c])) + dn + sb[col0 + c];
      }
    }
    tc_fence_before();
    // x^2 -> bf16 -> the one operand slot of the gamma contraction (K = NPO = 64 columns)
    {
      const uint32_t kc2 = gpos;
      if (edbg) { e_t0 = clock64(); e_p1 += e_t0 - e_t1; }
      mbar_wait(&R.empty_bar[kc2 % R.nslots], ((kc2 / R.nslots) & 1) ^ 1);
      if (edbg) { e_t1 = clock64(); e_slot += e_t1 - e_t0; }
      const uint32_t a_addr = R.ring_base + (kc2 % R.nslots) * R.slot_bytes;
      const uint32_t row_off = (uint32_t)r * 128u, rx = (uint32_t)(r & 7);
#pragma unroll
      for (int j = 0; j < CPT / 8; ++j) {
        const uint32_t chunk = (uint32_t)((col0 + j * 8) >> 3);
        const float* x8 = &xr[j * 8];
        st_shared_v4(a_addr + row_off + ((chunk ^ rx) << 4), pack_bf16x2(x8[0] * x8[0], x8[1] * x8[1]),
                     pack_bf16x2(x8[2] * x8[2], x8[3] * x8[3]), pack_bf16x2(x8[4] * x8[4], x8[5] * x8[5]),
                     pack_bf16x2(x8[6] * x8[6], x8[7] * x8[7]));
      }
      fence_async_smem();
      mbar_arrive_cluster(R.x2_ready_cl + 8u * bsel);
    }
    tc_n2 = decode_it(it + 2);
    if (P.tail_x && tc.n0 != w_n && tc.n0 < P.B) {       // warp-uniform: tiles of one image come in runs
      w_n = tc.n0;
#pragma unroll
      for (int i = 0; i < 48; ++i) wreg[i] = __ldg(P.tail_w + (long long)w_n * 48 + i);
    }
    if (edbg) { e_t0 = clock64(); e_p1 += e_t0 - e_t1; }
    mbar_wait(&R.norm_full[bsel], par);
    if (edbg) { e_t1 = clock64(); e_norm += e_t1 - e_t0; }
    tc_fence_after();
    // ---- pass 2: IGDN  out = x * sqrt(norm + beta) = x * n * rsqrt(n)   (GDN: x * rsqrt(n)) ----
    {
      uint32_t tr[CPT];
      tmem_ldn<32>(tbuf + col0, tr);           // the contraction wrote norm over columns [0, NPO) of the buffer
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_cluster(R.buf_free_cl + 8u * bsel);
      const bool igdn = (P.act == LDIC_ACT_IGDN);
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float nrm = __uint_as_float(tr[c]) + R.s_beta[col0 + c];
        const float rs = rsqrt_approx(nrm);
        xr[c] *= igdn ? nrm * rs : rs;
      }
    }
    unsigned long long tail_acc = 0;
    if (valid) {
      if (P.out) {
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j) {
          const int col = col0 + j * 8;
          const int g = col >> P.cg_shift, cc = col & (P.Cg - 1);
          const long long off = pix_base + (long long)(g >> 1) * P.out_sY + (long long)(g & 1) * P.out_sX + cc;
          const float* x8 = &xr[j * 8];
          if (P.out_f32) {
            st_global_v8(reinterpret_cast<float*>(P.out) + off, x8);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(P.out) + off);
            *dst = make_uint4(pack_bf16x2(x8[0], x8[1]), pack_bf16x2(x8[2], x8[3]), pack_bf16x2(x8[4], x8[5]),
                              pack_bf16x2(x8[6], x8[7]));
          }
        }
      }
      if (P.tail_x) {
        // batch_conv (model/net.py:527-537, :811) on this thread's two output pixels with the filter in registers,
        // then the a11 squared level error against the input image (model/net.py:864-868)
        const int oy0 = gy_ * P.sy + jb.oy_off, ox0 = gx_ * P.sx + jb.ox_off;
#pragma unroll
        for (int jj = 0; jj < CPT; jj += 16) {
          const int g = (col0 + jj) >> 4;
          const int oy = oy0 + (g >> 1), ox = ox0 + (g & 1);
          float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int m = 0; m < 16; ++m) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c] = fmaf(xr[jj + m], wreg[c * 16 + m], o[c]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const long long idx = (((long long)gn_ * 3 + c) * P.tail_H + oy) * P.tail_W + ox;
            if (P.tail_tanh) o[c] = tanhf(o[c]);
            if (P.tail_xo) P.tail_xo[idx] = o[c];
            tail_acc += tail_sq_err(P, idx, o[c]);
          }
        }
      }
    }
    if (edbg) e_p2 += clock64() - e_t1;
    if (P.tail_x) {                            // TN == 1: one image per tile, one warp-level sum and one atomic per warp
      unsigned long long v = tail_acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v) atomicAdd(P.tail_sq + tc.n0, v);
    }
  }
  if (edbg && lane == 0) {
    P.dbg[16] = (unsigned long long)(clock64() - e_begin); P.dbg[17] = (unsigned long long)e_acc;
    P.dbg[18] = (unsigned long long)e_norm; P.dbg[19] = (unsigned long long)e_slot; P.dbg[20] = (unsigned long long)ntiles_cta;
    P.dbg[21] = (unsigned long long)e_p1; P.dbg[22] = (unsigned long long)e_p2;
  }
}

// ---------------------------------------------------------------------------------
// CTA-pair variant of the streaming kernel (tcgen05.mma.cta_group::2, M = 256): the two CTAs of a cluster work
// on two adjacent tiles; each stage holds the CTA's own 128-pixel A tile and HALF of the weight rows, so a stage
// is 16 KB + Np*64 B instead of 16 KB + Np*128 B (more stages in flight, 30 % less L2 -> SM traffic, half the
// B-operand shared-memory reads per SM).  Barrier protocol as in conv_halo2_kernel below: TMA completions of
// both CTAs land on the leader's full barriers, tcgen05.commit multicasts to both CTAs, the peer's epilogue
// reaches the leader's x2_ready / buf_free barriers through shared::cluster addresses.
// ---------------------------------------------------------------------------------
// W3 = true: wide-N form of the merged last deconv (see epilogue_w3): NP = 3 * NPO accumulator columns, bias / beta /
// gamma refer to the NPO logical columns and the gamma contraction is an N = NPO MMA.
// TF = true: TF32 parity mode (kind::tf32): activations, weights and gamma are fp32 in memory; the host describes them
// to TMA as bf16 tensors with twice the channels, so every byte count, box and descriptor below is unchanged.
template <int NP, bool W3 = false, bool RES = false, bool TF = false>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmG, const __grid_constant__ ConvParams P) {
  constexpr int NPO = W3 ? NP / 3 : NP;
  constexpr int kBHalfBytes = (NP / 2) * kBlockK * 2;
  constexpr int kGHalfBytes = (NPO / 2) * kBlockK * 2;
  constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  constexpr uint32_t kFmt = TF ? 2u : 1u;      // a / b format: 1 = BF16, 2 = TF32
  constexpr uint32_t kIdesc2 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  constexpr uint32_t kIdescG = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(NPO >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = P.stages;
  uint8_t* aux = smem_al + (size_t)stages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);              // [kMaxStages]  (leader's are the live ones)
  uint64_t* empty_bar = full_bar + kMaxStages;                         // [kMaxStages]
  uint64_t* acc_full = empty_bar + kMaxStages;                         // [2]
  uint64_t* buf_free = acc_full + 2;                                   // leader only
  uint64_t* x2_ready = buf_free + 2;                                   // leader only
  uint64_t* norm_full = x2_ready + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(norm_full + 2);
  float* s_bias = reinterpret_cast<float*>(aux + 256);
  float* s_beta = s_bias + NP * P.nbias;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool gdn = (P.act == LDIC_ACT_GDN || P.act == LDIC_ACT_IGDN);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&buf_free[i], 2 * kEpiThreads);
      mbar_init(&x2_ready[i], 2 * kEpiThreads);
      mbar_init(&norm_full[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    if (gdn) prefetch_tmap(&tmG);
  }
  if (warp == kMmaWarp) tmem_alloc_cg2(tmem_ptr, kTmemCols);
  for (int i = threadIdx.x; i < NPO * P.nbias; i += kThreads) s_bias[i] = P.bias ? P.bias[i] : 0.f;
  for (int i = threadIdx.x; i < NPO; i += kThreads) s_beta[i] = (gdn && P.beta) ? P.beta[i] : 1.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int gk = gdn ? P.gdn_kblocks : 0;
  const int npairs = (int)gridDim.x >> 1, pi = (int)blockIdx.x >> 1;
  const int total_super = P.super_per_job * P.njobs;
  const int nt = (total_super - pi + npairs - 1) / npairs;             // super tiles of this pair
  const uint32_t full_L = mapa_shared(smem_u32(full_bar), 0);

  if (warp >= kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == kProdWarp || warp == kProdBWarp) {
      // ===================== TMA producers (both CTAs): warp kProdWarp loads the activation tiles, warp
      // kProdBWarp the weight / gamma half tiles.  One warp issuing both copies of a stage spent ~420 cycles per
      // stage against the 384 cycles the tensor pipe needs for it (wait_full 133 cycles per stage in the MMA warp);
      // two warps walk the same ring in lockstep order, each waiting on the stage's empty barrier itself.  The A
      // warp's expect_tx carries the bytes of the whole stage (a complete_tx that lands first only drives the
      // transaction count negative while the arrival is still pending); gamma stages belong to the B warp alone.
      auto produce = [&](auto is_a_tag) {
        constexpr bool IS_A = decltype(is_a_tag)::value;
        uint32_t slot = 0, ph = 0;
        int stages_r = stages;
        asm volatile("" : "+r"(stages_r));
        const bool pdbg = IS_A && P.dbg != nullptr && blockIdx.x == 0;
        long long t_empty = 0, tp0 = 0;
        const long long tp_begin = clock64();
        const int row_half = (int)rank * (NP / 2), row_half_g = (int)rank * (NPO / 2);
        auto advance = [&]() { if (++slot == (uint32_t)stages_r) { slot = 0; ph ^= 1; } };
        auto load_gamma = [&]() {
          for (int kb = 0; kb < gk; ++kb) {                     // gamma K-blocks ride the same ring (B half only)
            if constexpr (!IS_A) {
              mbar_wait(&empty_bar[slot], ph ^ 1);
              if (elect_one()) {
                if (leader) mbar_expect_tx(&full_bar[slot], 2 * kGHalfBytes);
                tma_load_2d_cg2(smem_base + slot * kStageBytes + kATileBytes, &tmG, full_L + 8u * slot, kb * kBlockK, row_half_g);
              }
            }
            advance();
          }
        };
        const int cs = P.mode == 2 ? 2 : 1;                     // input pixels per tile pixel (strided TMA box)
        int cur_job = -1, ntaps = 0, tap_begin = 0, nkb = 0;
        uint32_t my_w0 = 0, my_w1 = 0;                          // lane t: tap t of the current job, packed
        for (int it = 0; it < nt; ++it) {
          const TileCoord tc = decode_tile2(P, pi + it * npairs, (int)rank);
          if (tc.job != cur_job) {
            cur_job = tc.job;
            ntaps = P.jobs[cur_job].ntaps; tap_begin = P.jobs[cur_job].tap_begin; nkb = P.jobs[cur_job].nkb;
            if (lane < ntaps) {
              const Tap t = P.taps[tap_begin + lane];
              my_w0 = (uint32_t)(uint8_t)t.dx | ((uint32_t)(uint8_t)t.dy << 8) | ((uint32_t)t.nkc << 24);
              my_w1 = (uint32_t)t.a_c0 | ((uint32_t)t.b_c0 << 16);
            }
          }
          const int jins = (gk && it > 0) ? (nkb < P.gdn_insert ? nkb : P.gdn_insert) : -1;
          const int x0 = tc.x0 * cs, y0 = tc.y0 * cs;
          int cb = 0;
          for (int tp = 0; tp < ntaps; ++tp) {
            const uint32_t w0 = __shfl_sync(0xffffffffu, my_w0, tp), w1 = __shfl_sync(0xffffffffu, my_w1, tp);
            const int dx = (int)(int8_t)(w0 & 0xff), dy = (int)(int8_t)((w0 >> 8) & 0xff);
            const int nkc = (int)(w0 >> 24), a_c0 = (int)(w1 & 0xffff), b_c0 = (int)(w1 >> 16);
            const int brow = (tap_begin + tp) * NP + row_half;
            for (int kc = 0; kc < nkc; ++kc, ++cb) {
              if (cb == jins) load_gamma();
              const uint32_t a_dst = smem_base + slot * kStageBytes;
              if (pdbg) tp0 = clock64();
              mbar_wait(&empty_bar[slot], ph ^ 1);
              if (pdbg) t_empty += clock64() - tp0;
              if (elect_one()) {
                if constexpr (IS_A) {
                  if (leader) mbar_expect_tx(&full_bar[slot], 2 * kStageBytes);
                  tma_load_4d_cg2(a_dst, &tmA, full_L + 8u * slot, a_c0 + kc * kBlockK, x0 + dx, y0 + dy, tc.n0);
                } else {
                  tma_load_2d_cg2(a_dst + kATileBytes, &tmW, full_L + 8u * slot, b_c0 + kc * kBlockK, brow);
                }
              }
              advance();
            }
          }
          if (cb == jins) load_gamma();
        }
        if (gk) load_gamma();                                   // contraction of the last tile
        if (pdbg && lane == 0) { P.dbg[8] = (unsigned long long)(clock64() - tp_begin); P.dbg[9] = (unsigned long long)t_empty; }
      };
      if (warp == kProdWarp) produce(std::true_type{}); else produce(std::false_type{});
    } else if (warp == kMmaWarp && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      uint32_t slot = 0, ph = 0;
      int stages_r = stages;
      asm volatile("" : "+r"(stages_r));
      const uint32_t hi = desc_hi(1024);
      const uint32_t a_lo0 = desc_lo(smem_base), b_lo0 = desc_lo(smem_base + kATileBytes);
      uint32_t slot_lo = 0;
      const bool dbg = P.dbg != nullptr && blockIdx.x == 0;
      long long t_full = 0, t_buf = 0, t_x2 = 0, n_st = 0, t0 = 0;
      const long long t_begin = clock64();
      auto mma_stage = [&](uint32_t d_tmem, bool first, uint32_t idesc) {
        if (dbg) t0 = clock64();
        mbar_wait_cl(&full_bar[slot], ph);
        if (dbg) { t_full += clock64() - t0; ++n_st; }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t alo = a_lo0 + slot_lo, blo = b_lo0 + slot_lo;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            if constexpr (TF) umma_tf32_lh_cg2(d_tmem, alo + 2 * k, hi, blo + 2 * k, hi, idesc, !(first && k == 0));
            else umma_bf16_lh_cg2(d_tmem, alo + 2 * k, hi, blo + 2 * k, hi, idesc, !(first && k == 0));
          tc_commit_mc(&empty_bar[slot]);                    // frees the slot in both CTAs when these MMAs retire
        }
        ++slot; slot_lo += (uint32_t)(kStageBytes >> 4);
        if (slot == (uint32_t)stages_r) { slot = 0; slot_lo = 0; ph ^= 1; }
      };
      auto gdn_of = [&](int j) {                             // norm(j) = x^2 . gamma^T, in place over acc(j)
        const int bsel = j & 1;
        if (dbg) t0 = clock64();
        mbar_wait_cl(&x2_ready[bsel], (j >> 1) & 1);          // x^2 tiles written by the epilogue warps of both CTAs
        if (dbg) t_x2 += clock64() - t0;
        tc_fence_after();
        for (int kb = 0; kb < gk; ++kb) mma_stage(tmem_base + bsel * kBufCols, kb == 0, kIdescG);
        if (elect_one()) tc_commit_mc(&norm_full[bsel]);
      };
      int cur_job = -1, nkb = 0;
      for (int it = 0; it < nt; ++it) {
        const int job = (pi + it * npairs) / P.super_per_job;
        if (job != cur_job) { cur_job = job; nkb = P.jobs[cur_job].nkb; }
        const int jins = (gk && it > 0) ? (nkb < P.gdn_insert ? nkb : P.gdn_insert) : -1;
        const int bsel = it & 1;
        if (dbg) t0 = clock64();
        mbar_wait_cl(&buf_free[bsel], ((it >> 1) & 1) ^ 1);
        if (dbg) t_buf += clock64() - t0;
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          if (kb == jins) gdn_of(it - 1);
          mma_stage(tmem_base + bsel * kBufCols, kb == 0, kIdesc2);
        }
        if (elect_one()) tc_commit_mc(&acc_full[bsel]);
        if (nkb == jins) gdn_of(it - 1);
      }
      if (gk) gdn_of(nt - 1);
      if (dbg && lane == 0) {
        P.dbg[0] = (unsigned long long)(clock64() - t_begin); P.dbg[1] = (unsigned long long)t_full;
        P.dbg[2] = (unsigned long long)t_buf; P.dbg[3] = (unsigned long long)t_x2; P.dbg[4] = (unsigned long long)n_st;
        P.dbg[5] = (unsigned long long)nt;
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own tile) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    EpiRing R;
    R.tma_out = nullptr; R.epi_threads = 0;
    R.pair_store = 1;
    R.ring_base = smem_base; R.slot_bytes = kStageBytes; R.nslots = (uint32_t)stages; R.empty_bar = empty_bar;
    R.acc_full = acc_full; R.buf_free = buf_free; R.x2_ready = x2_ready; R.norm_full = norm_full;
    R.s_bias = s_bias; R.s_beta = s_beta; R.insert_after = P.gdn_insert;
    R.t_first = pi; R.t_stride = npairs; R.rank = (int)rank;
    R.buf_free_cl = mapa_shared(smem_u32(buf_free), 0); R.x2_ready_cl = mapa_shared(smem_u32(x2_ready), 0);
    if constexpr (W3) epilogue_w3<NP>(P, R, tmem_base, gk, nt, warp, lane);
    else epilogue_role<NP, true, kEpiWarps, RES, TF>(P, R, tmem_base, gk, nt, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------
// Wide-accumulator CTA-pair kernel: NP = 384 output channels (the reference's `--high` model, N = 384,
// model/net.py:446-451).  A GDN layer then needs 384 accumulator columns + 384 norm columns > the 512 TMEM columns,
// so the accumulator is single-buffered in columns [0, NP) and the gamma contraction is tiled over its
// OUTPUT channels: norm chunk c (128 columns, TMEM [NP, NP+128)) = x^2[128 px, NP] . gamma[128c..128c+128, :]^T with all
// NP/64 x^2 operand tiles resident in shared memory; pass 2 re-reads x from the accumulator chunk by chunk, so no
// thread ever holds more than 64 + 64 fp32 values.
//   * MMA N <= 256: every K step issues two M256 x N(NP/2) MMAs (column halves 0 and NP/2).
//   * Two rings instead of one: A (16 KB activation tiles; NP/64 of its slots carry the x^2 tiles during the GDN
//     phase) and B (this CTA's half of the weight rows of both column halves = NP/2 rows x 128 B; during the GDN
//     phase one slot carries 3 K blocks of this CTA's 64 gamma rows of the current norm chunk).
//   * Schedule per tile: main loop -> acc_full -> pass 1 (x^2 tiles) -> per norm chunk: [gamma MMAs -> norm_full ->
//     pass 2 of the chunk (the MMAs of chunk c+1 run under the arithmetic and stores of chunk c)] -> acc_free.
//     The tensor pipe idles during passes 1 and 2 (~17 k cycles per tile against 115 k cycles of MMAs for the
//     5x5 conv at N = 384), the TMA producers run ahead into the free ring slots meanwhile.
// Layers without GDN at NP = 384 (conv 4, h_a, h_s, context model) use the same kernel with the plain epilogue.
// ---------------------------------------------------------------------------------
constexpr int kWideNC = 128;            // norm chunk (columns)
constexpr int kWideGPI = 3;             // gamma K blocks per B-ring slot

template <int NP>
__global__ void __launch_bounds__(kThreads, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmG, const __grid_constant__ ConvParams P) {
  constexpr int NH = NP / 2;                               // columns per MMA
  constexpr int kBSlotBytes = (NP / 2) * kBlockK * 2;      // 2 x (NH / 2) weight rows
  constexpr int kBHalfOff = (NH / 2) * kBlockK * 2;        // column half 1 inside a B slot
  constexpr int GK = NP / 64;                              // x^2 operand tiles
  constexpr int NCH = NP / kWideNC;                        // norm chunks
  constexpr int kGTileBytes = (kWideNC / 2) * kBlockK * 2; // this CTA's 64 gamma rows of one K block
  constexpr int GITEMS = GK / kWideGPI;                    // B-ring items per norm chunk
  static_assert(NP % 128 == 0 && NP > 256 && NP + kWideNC <= kTmemCols, "wide kernel: 256 < NP <= 384, multiple of 128");
  static_assert(GK % kWideGPI == 0 && kWideGPI * kGTileBytes <= kBSlotBytes, "gamma items must fit a B slot");
  static_assert(kBHalfOff % 1024 == 0 && kGTileBytes % 1024 == 0, "operand tiles start on swizzle atoms");
  constexpr uint32_t kIdescH = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  constexpr uint32_t kIdescG = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kWideNC >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int SA = P.SA, SB = P.SB;
  const uint32_t b_base = smem_base + (uint32_t)SA * kATileBytes;
  uint8_t* aux = smem_al + (size_t)SA * kATileBytes + (size_t)SB * kBSlotBytes;
  uint64_t* afull = reinterpret_cast<uint64_t*>(aux);                  // [kMaxStages]  (leader's are the live ones)
  uint64_t* aempty = afull + kMaxStages;
  uint64_t* bfull = aempty + kMaxStages;                               // leader's
  uint64_t* bempty = bfull + kMaxStages;
  uint64_t* acc_full = bempty + kMaxStages;                            // [1] MMA -> epilogues (multicast)
  uint64_t* acc_free = acc_full + 1;                                   // [1] leader only: epilogues of both CTAs -> MMA
  uint64_t* x2_ready = acc_free + 1;                                   // [1] leader only
  uint64_t* norm_full = x2_ready + 1;                                  // [1] multicast, one phase per norm chunk
  uint64_t* norm_free = norm_full + 1;                                 // [1] leader only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(norm_free + 1);
  float* s_bias = reinterpret_cast<float*>(aux + 512);
  float* s_beta = s_bias + NP * P.nbias;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool gdn = (P.act == LDIC_ACT_GDN || P.act == LDIC_ACT_IGDN);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 2 * kEpiThreads);
    mbar_init(x2_ready, 2 * kEpiThreads);
    mbar_init(norm_full, 1);
    mbar_init(norm_free, 2 * kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    if (gdn) prefetch_tmap(&tmG);
  }
  if (warp == kMmaWarp) tmem_alloc_cg2(tmem_ptr, kTmemCols);
  for (int i = threadIdx.x; i < NP * P.nbias; i += kThreads) s_bias[i] = P.bias ? P.bias[i] : 0.f;
  for (int i = threadIdx.x; i < NP; i += kThreads) s_beta[i] = (gdn && P.beta) ? P.beta[i] : 1.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int gk = gdn ? GK : 0;
  const int npairs = (int)gridDim.x >> 1, pi = (int)blockIdx.x >> 1;
  const int total_super = P.super_per_job * P.njobs;
  const int nt = (total_super - pi + npairs - 1) / npairs;
  const uint32_t afull_L = mapa_shared(smem_u32(afull), 0), bfull_L = mapa_shared(smem_u32(bfull), 0);

  if (warp >= kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == kProdWarp) {
      // ===================== A producer: activation tiles (both CTAs) =====================
      uint32_t sa = 0, pa = 0;
      auto adv = [&]() { if (++sa == (uint32_t)SA) { sa = 0; pa ^= 1; } };
      const int cs = P.mode == 2 ? 2 : 1;
      int cur_job = -1, ntaps = 0, tap_begin = 0;
      uint32_t my_w0 = 0, my_w1 = 0;
      for (int it = 0; it < nt; ++it) {
        const TileCoord tc = decode_tile2(P, pi + it * npairs, (int)rank);
        if (tc.job != cur_job) {
          cur_job = tc.job;
          ntaps = P.jobs[cur_job].ntaps; tap_begin = P.jobs[cur_job].tap_begin;
          if (lane < ntaps) {
            const Tap t = P.taps[tap_begin + lane];
            my_w0 = (uint32_t)(uint8_t)t.dx | ((uint32_t)(uint8_t)t.dy << 8) | ((uint32_t)t.nkc << 24);
            my_w1 = (uint32_t)t.a_c0;
          }
        }
        const int x0 = tc.x0 * cs, y0 = tc.y0 * cs;
        for (int tp = 0; tp < ntaps; ++tp) {
          const uint32_t w0 = __shfl_sync(0xffffffffu, my_w0, tp), w1 = __shfl_sync(0xffffffffu, my_w1, tp);
          const int dx = (int)(int8_t)(w0 & 0xff), dy = (int)(int8_t)((w0 >> 8) & 0xff);
          const int nkc = (int)(w0 >> 24), a_c0 = (int)w1;
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&aempty[sa], pa ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&afull[sa], 2 * kATileBytes);
              tma_load_4d_cg2(smem_base + sa * kATileBytes, &tmA, afull_L + 8u * sa, a_c0 + kc * kBlockK, x0 + dx, y0 + dy, tc.n0);
            }
            adv();
          }
        }
        // The tile's gk x^2 slots are written by the epilogue warps, not by TMA.  Their `afull` barriers still have to
        // complete one phase per ring pass (the MMA warp's parity bookkeeping is per pass), so the leader arrives on them
        // once the slot's previous occupant has been consumed.
        for (int kb = 0; kb < gk; ++kb) {
          mbar_wait(&aempty[sa], pa ^ 1);
          if (leader && elect_one()) mbar_arrive(&afull[sa]);
          adv();
        }
      }
    } else if (warp == kProdBWarp) {
      // ===================== B producer: weight rows of both column halves, gamma rows (both CTAs) ============
      uint32_t sb = 0, pb = 0;
      auto adv = [&]() { if (++sb == (uint32_t)SB) { sb = 0; pb ^= 1; } };
      const int row_half = (int)rank * (NH / 2), row_g = (int)rank * (kWideNC / 2);
      int cur_job = -1, ntaps = 0, tap_begin = 0;
      uint32_t my_w0 = 0, my_w1 = 0;
      for (int it = 0; it < nt; ++it) {
        const int job = (pi + it * npairs) / P.super_per_job;
        if (job != cur_job) {
          cur_job = job;
          ntaps = P.jobs[cur_job].ntaps; tap_begin = P.jobs[cur_job].tap_begin;
          if (lane < ntaps) {
            const Tap t = P.taps[tap_begin + lane];
            my_w0 = (uint32_t)t.nkc;
            my_w1 = (uint32_t)t.b_c0;
          }
        }
        for (int tp = 0; tp < ntaps; ++tp) {
          const int nkc = (int)__shfl_sync(0xffffffffu, my_w0, tp), b_c0 = (int)__shfl_sync(0xffffffffu, my_w1, tp);
          const int brow = (tap_begin + tp) * NP + row_half;
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&bempty[sb], pb ^ 1);
            if (elect_one()) {
              const uint32_t dst = b_base + sb * kBSlotBytes;
              if (leader) mbar_expect_tx(&bfull[sb], 2 * kBSlotBytes);
              tma_load_2d_cg2(dst, &tmW, bfull_L + 8u * sb, b_c0 + kc * kBlockK, brow);
              tma_load_2d_cg2(dst + kBHalfOff, &tmW, bfull_L + 8u * sb, b_c0 + kc * kBlockK, brow + NH);
            }
            adv();
          }
        }
        if (gk) {
          for (int c = 0; c < NCH; ++c) {
            for (int j = 0; j < GITEMS; ++j) {
              mbar_wait(&bempty[sb], pb ^ 1);
              if (elect_one()) {
                const uint32_t dst = b_base + sb * kBSlotBytes;
                if (leader) mbar_expect_tx(&bfull[sb], 2 * kWideGPI * kGTileBytes);
#pragma unroll
                for (int kbi = 0; kbi < kWideGPI; ++kbi)
                  tma_load_2d_cg2(dst + kbi * kGTileBytes, &tmG, bfull_L + 8u * sb, (j * kWideGPI + kbi) * kBlockK,
                                  c * kWideNC + row_g);
              }
              adv();
            }
          }
        }
      }
    } else if (warp == kMmaWarp && leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      auto adv_a = [&]() { if (++sa == (uint32_t)SA) { sa = 0; pa ^= 1; } };
      auto adv_b = [&]() { if (++sb == (uint32_t)SB) { sb = 0; pb ^= 1; } };
      const uint32_t hi = desc_hi(1024);
      const uint32_t a_lo0 = desc_lo(smem_base), b_lo0 = desc_lo(b_base);
      int cur_job = -1, nkb = 0;
      for (int it = 0; it < nt; ++it) {
        const int job = (pi + it * npairs) / P.super_per_job;
        if (job != cur_job) { cur_job = job; nkb = P.jobs[cur_job].nkb; }
        mbar_wait_cl(acc_free, ((uint32_t)it & 1u) ^ 1u);          // the previous tile's accumulator has been read
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_cl(&afull[sa], pa);
          mbar_wait_cl(&bfull[sb], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t alo = a_lo0 + sa * (uint32_t)(kATileBytes >> 4), blo = b_lo0 + sb * (uint32_t)(kBSlotBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              umma_bf16_lh_cg2(tmem_base, alo + 2 * k, hi, blo + 2 * k, hi, kIdescH, (kb | k) != 0);
              umma_bf16_lh_cg2(tmem_base + NH, alo + 2 * k, hi, blo + (uint32_t)(kBHalfOff >> 4) + 2 * k, hi, kIdescH, (kb | k) != 0);
            }
            tc_commit_mc(&aempty[sa]);
            tc_commit_mc(&bempty[sb]);
          }
          adv_a(); adv_b();
        }
        if (elect_one()) tc_commit_mc(acc_full);
        if (gk) {
          mbar_wait_cl(x2_ready, (uint32_t)it & 1u);                // x^2 tiles of both CTAs are in the next gk A slots
          tc_fence_after();
          for (int c = 0; c < NCH; ++c) {
            const uint32_t g = (uint32_t)(it * NCH + c);
            mbar_wait_cl(norm_free, (g & 1u) ^ 1u);                 // the previous norm chunk has been read
            tc_fence_after();
            for (int j = 0; j < GITEMS; ++j) {
              mbar_wait_cl(&bfull[sb], pb);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t blo = b_lo0 + sb * (uint32_t)(kBSlotBytes >> 4);
#pragma unroll
                for (int kbi = 0; kbi < kWideGPI; ++kbi) {
                  const int kb = j * kWideGPI + kbi;
                  uint32_t xs = sa + (uint32_t)kb; if (xs >= (uint32_t)SA) xs -= (uint32_t)SA;
                  const uint32_t alo = a_lo0 + xs * (uint32_t)(kATileBytes >> 4);
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma_bf16_lh_cg2(tmem_base + NP, alo + 2 * k, hi, blo + (uint32_t)(kbi * (kGTileBytes >> 4)) + 2 * k, hi,
                                     kIdescG, (kb | k) != 0);
                }
                tc_commit_mc(&bempty[sb]);
              }
              adv_b();
            }
            if (elect_one()) tc_commit_mc(norm_full);
          }
          for (int kb = 0; kb < gk; ++kb) {                          // hand the x^2 slots back to the A producers
            if (elect_one()) tc_commit_mc(&aempty[sa]);
            adv_a();
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own tile) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int q = warp & 3, h = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int xi = r & (P.TW - 1), yi = (r >> P.tw_shift) & (P.TH - 1), ni = r >> (P.tw_shift + P.th_shift);
    const bool igdn = (P.act == LDIC_ACT_IGDN);
    const uint32_t acc_free_cl = mapa_shared(smem_u32(acc_free), 0), x2_ready_cl = mapa_shared(smem_u32(x2_ready), 0),
                   norm_free_cl = mapa_shared(smem_u32(norm_free), 0);
    const uint32_t tacc = tmem_base + lane_sel, tnorm = tmem_base + lane_sel + NP;
    const uint32_t row_off = (uint32_t)r * 128u, rx = (uint32_t)(r & 7);
    const bool odd = lane & 1;
    uint32_t apos = 0;                           // A-ring position (mod SA) of the tile's first stage
    for (int it = 0; it < nt; ++it) {
      const TileCoord tc = decode_tile2(P, pi + it * npairs, (int)rank);
      const Job jb = P.jobs[tc.job];
      const int gx_ = tc.x0 + xi, gy_ = tc.y0 + yi, gn_ = tc.n0 + ni;
      const bool valid = (gx_ < P.Wg) && (gy_ < P.Hg) && (gn_ < P.B) && !(P.dbg_nostore & 1);
      const long long pix_base = (long long)gn_ * P.out_sN + (long long)(gy_ * P.sy + jb.oy_off) * P.out_sY +
                                 (long long)(gx_ * P.sx + jb.ox_off) * P.out_sX + jb.out_off;
      const long long pb_other = __shfl_xor_sync(0xffffffffu, pix_base, 1);
      const bool valid_other = __shfl_xor_sync(0xffffffffu, (int)valid, 1) != 0;
      const float* sb = s_bias + (P.nbias > 1 ? tc.job * NP : 0);
      // 32 finished columns [col, col+32) of this thread's pixel -> global (fp32: own sectors; bf16: lane-pair
      // transposed stores, 64 contiguous bytes per pixel per instruction)
      auto store32 = [&](const float (&v)[32], int col) {
        if (P.out_f32) {
          if (valid) {
            float* dst = reinterpret_cast<float*>(P.out) + pix_base + col;
#pragma unroll
            for (int j = 0; j < 4; ++j) st_global_v8(dst + 8 * j, &v[8 * j]);
          }
        } else {
          uint32_t pk[16], rv[8], d1[8], d2[8];
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
#pragma unroll
          for (int e = 0; e < 8; ++e) rv[e] = __shfl_xor_sync(0xffffffffu, odd ? pk[e] : pk[8 + e], 1);
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(P.out) + col + (odd ? 16 : 0);
#pragma unroll
          for (int e = 0; e < 8; ++e) { d1[e] = odd ? rv[e] : pk[e]; d2[e] = odd ? pk[8 + e] : rv[e]; }
          if (odd ? valid_other : valid) st_global_v8(out + (odd ? pb_other : pix_base), d1);
          if (odd ? valid : valid_other) st_global_v8(out + (odd ? pix_base : pb_other), d2);
        }
      };
      apos += (uint32_t)jb.nkb; apos %= (uint32_t)SA;               // first x^2 slot of this tile
      mbar_wait(acc_full, (uint32_t)it & 1u);
      tc_fence_after();
      if (gk) {
        // ---- pass 1: x = acc + bias; x^2 -> bf16 -> the tile's gk A slots.  All earlier users of those slots are MMAs
        // that completed before acc_full did (in-order tensor pipe), so no slot barrier is needed here. ----
#pragma unroll 1
        for (int cb = 0; cb < NH / 32; ++cb) {
          const int col = h * NH + cb * 32;
          uint32_t xu[32];
          tmem_ld32(tacc + col, xu);
          tmem_ld_wait();
          uint32_t xs = apos + (uint32_t)(col >> 6); if (xs >= (uint32_t)SA) xs -= (uint32_t)SA;
          const uint32_t a_addr = smem_base + xs * kATileBytes + row_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { x8[e] = __uint_as_float(xu[8 * j + e]) + sb[col + 8 * j + e]; x8[e] *= x8[e]; }
            const uint32_t chunk = (uint32_t)(((col & 63) >> 3) + j);
            st_shared_v4(a_addr + ((chunk ^ rx) << 4), pack_bf16x2(x8[0], x8[1]), pack_bf16x2(x8[2], x8[3]),
                         pack_bf16x2(x8[4], x8[5]), pack_bf16x2(x8[6], x8[7]));
          }
        }
        tc_fence_before();
        fence_async_smem();
        mbar_arrive_cluster(x2_ready_cl);
        // ---- pass 2, one norm chunk at a time: out = x * rsqrt(norm + beta)  (IGDN: x * sqrt) ----
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const uint32_t g = (uint32_t)(it * NCH + c);
          const int col = c * kWideNC + h * (kWideNC / 2);           // this thread's 64 columns of the chunk
          mbar_wait(norm_full, g & 1u);
          tc_fence_after();
          uint32_t n0[32], n1[32], x0[32];
          tmem_ld32(tnorm + h * (kWideNC / 2), n0);
          tmem_ld32(tnorm + h * (kWideNC / 2) + 32, n1);
          tmem_ld32(tacc + col, x0);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive_cluster(norm_free_cl);                         // the norm buffer can take the next chunk
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = __uint_as_float(x0[k]) + sb[col + k];
            const float nrm = __uint_as_float(n0[k]) + s_beta[col + k];
            const float rs = rsqrt_approx(nrm);
            v[k] = x * (igdn ? nrm * rs : rs);
          }
          tmem_ld32(tacc + col + 32, x0);                            // second half of x, in flight under the stores
          store32(v, col);
          tmem_ld_wait();
          if (c == NCH - 1) { tc_fence_before(); mbar_arrive_cluster(acc_free_cl); }   // the accumulator can take the next tile
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = __uint_as_float(x0[k]) + sb[col + 32 + k];
            const float nrm = __uint_as_float(n1[k]) + s_beta[col + 32 + k];
            const float rs = rsqrt_approx(nrm);
            v[k] = x * (igdn ? nrm * rs : rs);
          }
          store32(v, col + 32);
        }
        apos += (uint32_t)gk; apos %= (uint32_t)SA;
      } else {
        // ---- plain epilogue: bias + activation ----
#pragma unroll 1
        for (int cb = 0; cb < NH / 32; ++cb) {
          const int col = h * NH + cb * 32;
          uint32_t xu[32];
          tmem_ld32(tacc + col, xu);
          tmem_ld_wait();
          if (cb == NH / 32 - 1) { tc_fence_before(); mbar_arrive_cluster(acc_free_cl); }
          float v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            float x = __uint_as_float(xu[k]) + sb[col + k];
            if (P.act == LDIC_ACT_RELU) x = fmaxf(x, 0.f);
            else if (P.act == LDIC_ACT_LEAKY02) x = x > 0.f ? x : 0.2f * x;
            else if (P.act == LDIC_ACT_LEAKY001) x = x > 0.f ? x : 0.01f * x;
            v[k] = x;
          }
          store32(v, col);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------
// First analysis layer (Cin = 3): ZeroPad2d((1,2,1,2)) + Conv2d(3, NP, 5, 2) + GDN straight from the
// NCHW fp32 image (model/net.py:97-99).  No patch matrix in HBM: per tile of 64 x 2 output pixels
//   warp 8        TMA: one (136 x 7 x 3) fp32 box of the image into a 2-deep raw ring (out-of-image
//                 reads are zero-filled = the ZeroPad2d); once per CTA: both K blocks of the packed
//                 weights and the whole gamma matrix, which stay RESIDENT in shared memory
//   warps 10,11   patch builders: each thread turns two pixels' 5x5x3 windows into bf16 rows of the
//                 K-major, 128B-swizzled A operand (K = 75 padded to 80: one 64-wide block + one
//                 16-wide MMA step of a second block)
//   warp 9        MMA issuer: 5 K steps of conv, then the GDN contraction of the previous tile
//   warps 0..7    the common epilogue (x^2 operand tiles go into the same 16 KB ring as the A tiles)
// Ring order: see `a_pos` in the kernel (the x^2 slots of tile t-1 ride before, between or after the A slots of tile t).
// ---------------------------------------------------------------------------------
constexpr int kFirstTW = 64, kFirstTH = 2;
constexpr int kRawX0 = 4;                           // the box starts 4 columns left of the tile (TMA needs a 16-byte
                                                    // aligned start in the innermost dimension; only 1 column is padding)
constexpr int kRawW = 2 * kFirstTW + 4 + kRawX0;    // 136 input columns per box (131 used), 544-byte rows
constexpr int kRawH = 2 * kFirstTH + 3;             // 7 input rows
constexpr int kRawBytes = 3 * kRawH * kRawW * 4;    // 11088
constexpr int kRawSlot = (kRawBytes + 127) / 128 * 128;
// uint8 image (eval_net.py:84: x = (u/255)*2-1): the box starts 16 columns left of the tile (16-byte aligned start),
// rows are 160 bytes; out-of-image bytes are zero-filled by TMA but a zero LEVEL is x = -1, not the ZeroPad2d's 0,
// so the patch builders mask the out-of-image taps themselves.
constexpr int kRawX08 = 16;
constexpr int kRawW8 = (2 * kFirstTW + 4 + kRawX08 + 15) / 16 * 16;   // 160
constexpr int kRawBytes8 = 3 * kRawH * kRawW8;                       // 3360
constexpr int kStage8Off = kRawBytes8;                               // bf16 copy of the box behind it in the same raw slot
static_assert(kRawBytes8 % 16 == 0 && kRawBytes8 + 2 * kRawBytes8 <= kRawSlot, "uint8 box + its bf16 copy share a raw ring slot");
constexpr int kFirstK = 75;
constexpr int kBuildThreads = 64;                   // warps 10 and 11

__device__ __forceinline__ uint32_t ld_shared_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (uint32_t)v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

template <int NP, bool U8 = false, int EW = kEpiWarps>
__global__ void __launch_bounds__(EW * 32 + 128, 1)
conv_first_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ ConvParams P) {
  constexpr int kBTileBytes = NP * kBlockK * 2;
  constexpr int GK = NP / 64;
  // roles: warps 0..EW-1 epilogue, EW TMA, EW+1 MMA, EW+2 / EW+3 patch builders
  constexpr int kEpiThreads = EW * 32, kThreads = EW * 32 + 128, kProdWarp = EW, kMmaWarp = EW + 1, kProdBWarp = EW + 2;
  constexpr int kEpiWarps = EW;
  constexpr int kEpiRegs = EW == 8 ? 216 : 144;
  static_assert(EW == 8 || EW == 12, "first layer: 8 or 12 epilogue warps");
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = P.stages;                                              // 16 KB ring slots
  const uint32_t w_base = smem_base;                                   // 2 K blocks of weights
  const uint32_t g_base = w_base + 2 * kBTileBytes;                    // GK K blocks of gamma
  const uint32_t ring_base = g_base + GK * kBTileBytes;
  const uint32_t raw_base = ring_base + (uint32_t)S * kATileBytes;
  uint8_t* aux = smem_al + (size_t)(2 + GK) * kBTileBytes + (size_t)S * kATileBytes + 2 * kRawSlot;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);               // [kMaxStages] builders -> MMA
  uint64_t* empty_bar = full_bar + kMaxStages;                         // [kMaxStages] MMA -> builders / epilogue
  uint64_t* acc_full = empty_bar + kMaxStages;                         // [2]
  uint64_t* buf_free = acc_full + 2;
  uint64_t* x2_ready = buf_free + 2;
  uint64_t* norm_full = x2_ready + 2;
  uint64_t* rfull = norm_full + 2;                                     // [2] TMA -> builders
  uint64_t* rempty = rfull + 2;                                        // [2] builders -> TMA
  uint64_t* wfull = rempty + 2;                                        // [1] resident weights + gamma landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);
  float* s_bias = reinterpret_cast<float*>(aux + 256);
  float* s_beta = s_bias + NP;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool gdn = (P.act == LDIC_ACT_GDN || P.act == LDIC_ACT_IGDN);
  const int gk = gdn ? GK : 0;
  const int per_tile = 2 + gk;                                         // ring slots per tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kBuildThreads); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&buf_free[i], kEpiThreads);
      mbar_init(&x2_ready[i], kEpiThreads);
      mbar_init(&norm_full[i], 1);
      mbar_init(&rfull[i], 1);
      mbar_init(&rempty[i], kBuildThreads);
    }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    if (gdn) prefetch_tmap(&tmG);
    if (P.tma_store) prefetch_tmap(&tmY);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_ptr, kTmemCols);
  for (int i = threadIdx.x; i < NP; i += kThreads) {
    s_bias[i] = P.bias ? P.bias[i] : 0.f;
    s_beta[i] = (gdn && P.beta) ? P.beta[i] : 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int ntiles_cta = (P.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // Ring order (identical in all roles): [A0 A1](tile 0), then for every tile t >= 1: the first `ins` A slots of tile t,
  // the gk x^2 slots of tile t-1, the other A slots of tile t; the x^2 slots of the last tile close the sequence.
  // ins (tuning key first_insert, default 2): with 0 the GDN contraction of tile t-1 is issued before the conv stages of
  // tile t and never waits for the patch builders; bit-identical results, measured no faster.
  const int ins = gk ? P.gdn_insert : 2;
  auto a_pos = [&](int it, int a) -> uint32_t {          // ring position of A slot a (0 / 1) of tile it
    if (it == 0) return (uint32_t)a;
    return (uint32_t)(2 + per_tile * (it - 1) + a + (a < ins ? 0 : gk));
  };

  if (warp >= kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == kProdWarp) {
      // ===================== TMA: resident operands once, then the raw image boxes =====================
      if (elect_one()) {
        mbar_expect_tx(wfull, (uint32_t)(2 + gk) * kBTileBytes);
        tma_load_2d(w_base, &tmW, wfull, 0, 0);
        tma_load_2d(w_base + kBTileBytes, &tmW, wfull, kBlockK, 0);
        for (int kb = 0; kb < gk; ++kb) tma_load_2d(g_base + kb * kBTileBytes, &tmG, wfull, kb * kBlockK, 0);
      }
      for (int it = 0; it < ntiles_cta; ++it) {
        const TileCoord tc = decode_tile(P, blockIdx.x + it * gridDim.x);
        const int rs = it & 1;
        mbar_wait(&rempty[rs], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&rfull[rs], U8 ? kRawBytes8 : kRawBytes);
          // x start 4 (16) columns left of the tile: TMA needs a 16-byte aligned start in the innermost dimension
          tma_load_4d(raw_base + rs * kRawSlot, &tmX, &rfull[rs], 2 * tc.x0 - (U8 ? kRawX08 : kRawX0), 2 * tc.y0 - 1, 0, tc.n0);
        }
      }
    } else if (warp == kMmaWarp) {
      // ===================== MMA issuer =====================
      const uint32_t hi = desc_hi(1024);
      const uint32_t ring_lo = desc_lo(ring_base), w_lo = desc_lo(w_base), g_lo = desc_lo(g_base);
      auto gdn_of = [&](int j) {
        const int bsel = j & 1;
        // x^2 slots of tile j: after the first `ins` A slots of tile j+1, or right after tile j's own A slots when it is
        // the CTA's last tile (the same arithmetic as EpiRing::insert_after in the epilogue)
        const uint32_t pos0 = (uint32_t)(2 + per_tile * j) + (j + 1 < ntiles_cta ? (uint32_t)ins : 0u);
        mbar_wait(&x2_ready[bsel], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        for (int kb = 0; kb < gk; ++kb) {
          const uint32_t slot = (pos0 + (uint32_t)kb) % (uint32_t)S;
          if (elect_one()) {
            const uint32_t alo = ring_lo + slot * (kATileBytes >> 4), blo = g_lo + kb * (kBTileBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_lh(tmem_base + bsel * kBufCols, alo + 2 * k, hi, blo + 2 * k, hi, kIdesc, (kb | k) != 0);
            tc_commit(&empty_bar[slot]);
          }
        }
        if (elect_one()) tc_commit(&norm_full[bsel]);
      };
      mbar_wait(wfull, 0);
      for (int it = 0; it < ntiles_cta; ++it) {
        const int bsel = it & 1;
        const uint32_t d_tmem = tmem_base + bsel * kBufCols;
        if (gk && it > 0 && ins == 0) gdn_of(it - 1);
        mbar_wait(&buf_free[bsel], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        // K block 0: 64 patch values, K block 1: 16 (11 real + zero padding) -> one K step
        {
          const uint32_t pos = a_pos(it, 0), slot = pos % (uint32_t)S;
          mbar_wait(&full_bar[slot], (pos / (uint32_t)S) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t alo = ring_lo + slot * (kATileBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) umma_bf16_lh(d_tmem, alo + 2 * k, hi, w_lo + 2 * k, hi, kIdesc, k != 0);
            tc_commit(&empty_bar[slot]);
          }
        }
        if (gk && it > 0 && ins == 1) gdn_of(it - 1);
        {
          const uint32_t pos = a_pos(it, 1), slot = pos % (uint32_t)S;
          mbar_wait(&full_bar[slot], (pos / (uint32_t)S) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t alo = ring_lo + slot * (kATileBytes >> 4);
            umma_bf16_lh(d_tmem, alo, hi, w_lo + (kBTileBytes >> 4), hi, kIdesc, 1u);
            tc_commit(&empty_bar[slot]);
            tc_commit(&acc_full[bsel]);
          }
        }
        if (gk && it > 0 && ins >= 2) gdn_of(it - 1);
      }
      if (gk) gdn_of(ntiles_cta - 1);
    } else {
      // ===================== patch builders (warps 10, 11) =====================
      const int tb = (int)threadIdx.x - 32 * kProdBWarp;                // 0..63 = pixel x inside the tile (warps kProdBWarp, +1)
      for (int it = 0; it < ntiles_cta; ++it) {
        const int rs = it & 1;
        const uint32_t p0 = a_pos(it, 0), p1 = a_pos(it, 1);
        const uint32_t s0 = p0 % (uint32_t)S, s1 = p1 % (uint32_t)S;
        mbar_wait(&rfull[rs], (uint32_t)(it >> 1) & 1u);
        mbar_wait(&empty_bar[s0], ((p0 / (uint32_t)S) & 1u) ^ 1u);
        mbar_wait(&empty_bar[s1], ((p1 / (uint32_t)S) & 1u) ^ 1u);
        const uint32_t a0 = ring_base + s0 * kATileBytes, a1 = ring_base + s1 * kATileBytes;
        const uint32_t raw = raw_base + rs * kRawSlot;
        if (!(P.dbg_nostore & 2)) {
        if constexpr (U8) {
          // uint8 image: first turn the whole box into the bf16 operand values x = (u/255)*2-1 once (every level is used
          // by ~6 patches), zero outside the image (TMA zero-fills the LEVELS there, but a zero level is x = -1, not the
          // ZeroPad2d's 0).  bf16((u/255)*2-1) == bf16(fma(u, 2/255, -1)) for all 256 levels (tests/test_oracle_golden.py).
          const TileCoord tc = decode_tile(P, blockIdx.x + it * gridDim.x);
          const int bx0 = 2 * tc.x0 - kRawX08, by0 = 2 * tc.y0 - 1;      // image position of the box origin
          for (int g = tb; g < kRawBytes8 / 4; g += kBuildThreads) {
            const uint32_t w4 = ld_shared_u32(raw + 4u * (uint32_t)g);
            const int xx = (g % (kRawW8 / 4)) * 4, rr = (g / (kRawW8 / 4)) % kRawH;
            const bool rowok = (unsigned)(by0 + rr) < (unsigned)P.tail_H;
            float v4[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const bool ok = rowok && (unsigned)(bx0 + xx + b) < (unsigned)P.tail_W;
              v4[b] = ok ? __fmaf_rn((float)((w4 >> (8 * b)) & 0xffu), 2.f / 255.f, -1.f) : 0.f;
            }
            st_shared_v2(raw + kStage8Off + 8u * (uint32_t)g, pack_bf16x2(v4[0], v4[1]), pack_bf16x2(v4[2], v4[3]));
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kBuildThreads) : "memory");   // the two builder warps only
        }
#pragma unroll
        for (int ly = 0; ly < kFirstTH; ++ly) {
          const int r = ly * kFirstTW + tb;                             // tile row = TMEM lane
          const uint32_t src = U8 ? raw + kStage8Off + (uint32_t)((2 * ly) * kRawW8 + 2 * tb + (kRawX08 - 1)) * 2u
                                  : raw + (uint32_t)((2 * ly) * kRawW + 2 * tb + (kRawX0 - 1)) * 4u;
          const uint32_t row_off = (uint32_t)r * 128u, rx = (uint32_t)(r & 7);
#pragma unroll
          for (int j = 0; j < 10; ++j) {                                // 16-byte chunks: k = 8j .. 8j+7
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (U8) {
                uint32_t h2[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int k = 8 * j + 2 * e + u;                      // k = (ky*5 + kx)*3 + c
                  h2[u] = 0u;
                  if (k < kFirstK) h2[u] = ld_shared_u16(src + (uint32_t)(((k % 3) * kRawH + k / 15) * kRawW8 + (k / 3) % 5) * 2u);
                }
                pk[e] = h2[0] | (h2[1] << 16);
              } else {
                float v[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int k = 8 * j + 2 * e + u;                      // k = (ky*5 + kx)*3 + c
                  v[u] = 0.f;
                  if (k < kFirstK) v[u] = ld_shared_f32(src + (uint32_t)(((k % 3) * kRawH + k / 15) * kRawW + (k / 3) % 5) * 4u);
                }
                pk[e] = pack_bf16x2(v[0], v[1]);
              }
            }
            const uint32_t dst = (j < 8 ? a0 : a1) + row_off + ((((uint32_t)(j & 7)) ^ rx) << 4);
            st_shared_v4(dst, pk[0], pk[1], pk[2], pk[3]);
          }
        }
        }
        fence_async_smem();
        mbar_arrive(&full_bar[s0]);
        mbar_arrive(&full_bar[s1]);
        mbar_arrive(&rempty[rs]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    EpiRing R;
    R.tma_out = nullptr; R.epi_threads = 0;
    R.pair_store = 0;
    R.ring_base = ring_base; R.slot_bytes = kATileBytes; R.nslots = (uint32_t)S; R.empty_bar = empty_bar;
    R.acc_full = acc_full; R.buf_free = buf_free; R.x2_ready = x2_ready; R.norm_full = norm_full;
    R.s_bias = s_bias; R.s_beta = s_beta; R.insert_after = P.gdn_insert;
    if (P.tma_store) { R.tma_out = &tmY; R.epi_threads = kEpiThreads; }
    epilogue_role<NP, false, EW>(P, R, tmem_base, gk, ntiles_cta, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------
// Host side: tap tables, tiling, tensor maps
// ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

typedef CUresult (*PFN_replaceAddress)(CUtensorMap*, void*);
PFN_replaceAddress get_replace_address() {
  static PFN_replaceAddress fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapReplaceAddress", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_replaceAddress)p;
    tried = true;
  }
  return fn;
}

enum MapType { MAP_BF16_SW128 = 0, MAP_F32 = 1, MAP_U8 = 2 };   // operand tiles (swizzled) / raw image boxes (linear)
int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides = nullptr, int type = MAP_BF16_SW128) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(LDIC_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  if (elem_strides) for (int i = 0; i < rank; ++i) es[i] = elem_strides[i];
  const bool linear = type != MAP_BF16_SW128;
  CUresult r = enc(m, type == MAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (type == MAP_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), (cuuint32_t)rank,
                   const_cast<void*>(base), dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   linear ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(LDIC_ECUDA, "cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                box[0], box[1], rank > 2 ? box[2] : 0);
  return LDIC_OK;
}

struct Layer {          // everything derived from an LdicConvDesc
  int mode, k, transposed;
  int Np, ngroups, Cg;
  int njobs, ntaps_total, nbias;
  Job jobs[kMaxJobs];
  Tap taps[kMaxTaps];
  signed char tap_ky[kMaxTaps][4], tap_kx[kMaxTaps][4];   // per (tap, group): kernel cell or -1
  short tap_co_base[kMaxTaps];                            // first logical output channel of the tap's job
  int cin_map, map_N, map_M;   // 0: ci = k - cin_offset;  1: context conv1 (y | h2 channel segments)
  int Kw;                      // K extent (columns) of one packed weight block
  long long vC, vW, vH, vN;    // activation view behind the TMA map (mode 1 splits vW into parity x vW/2)
  long long vPitch;            // pixels between consecutive rows of the view in memory (= vW unless the input is a column band)
  int Wg, Hg, Bg;              // pixel grid the M tiles walk over
  int oB, Ho, Wo, Cs;          // output tensor [oB, Ho, Wo, Cs]
  long long out_sN, out_sY, out_sX;
  int sy, sx;
};

// rows/cols of the 5x5 transposed kernels hit by output parity p at input offset d (-1,0,+1); -1 = none
inline int gs_k(int p, int d) {   // o = 2i - 1 + k   (ZeroPad2d((1,0,1,0)) + k5 s2 p3 op1)
  if (p == 0) return d == 0 ? 1 : (d == -1 ? 3 : -1);
  return d == 1 ? 0 : (d == 0 ? 2 : 4);
}
inline int hs_k(int p, int d) {   // o = 2i - 2 + k   (k5 s2 p2 op1)
  if (p == 0) return d == 1 ? 0 : (d == 0 ? 2 : 4);
  return d == 1 ? 1 : (d == 0 ? 3 : -1);
}

int build_layer(const LdicConvDesc* d, Layer* L) {
  memset(L, 0, sizeof(*L));
  if (!d) return fail(LDIC_EINVAL, "conv: null desc");
  if (d->B < 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) return fail(LDIC_EINVAL, "conv: bad shape");
  if (d->Cin_pad % 64 || d->Cin_pad < d->Cin) return fail(LDIC_EINVAL, "conv: Cin_pad must be a multiple of 64 >= Cin");
  if (d->Cout_pad < d->Cout || d->Cout_pad % 8) return fail(LDIC_EINVAL, "conv: Cout_pad must be >= Cout and a multiple of 8");
  for (int t = 0; t < kMaxTaps; ++t) for (int g = 0; g < 4; ++g) L->tap_ky[t][g] = L->tap_kx[t][g] = -1;
  L->ngroups = 1; L->Cg = d->Cout_pad; L->sy = L->sx = 1; L->njobs = 1; L->nbias = 1;
  L->Wg = d->W; L->Hg = d->H; L->Bg = d->B; L->Ho = d->H; L->Wo = d->W; L->oB = d->B;
  L->vC = d->Cin_pad; L->vW = d->W; L->vH = d->H; L->vN = d->B;
  L->Kw = d->Cin_pad;
  const int nkc_full = d->Cin_pad / 64;
  auto add_tap = [&](int dx, int dy, int px, int a_c0 = 0, int b_c0 = 0, int nkc = -1) {
    Tap t;
    t.dx = (short)dx; t.dy = (short)dy; t.px = (short)px; t.nkc = (short)(nkc < 0 ? nkc_full : nkc);
    t.a_c0 = a_c0; t.b_c0 = b_c0;
    L->taps[L->ntaps_total] = t;
    return L->ntaps_total++;
  };
  auto job = [](int ntaps, int tap_begin, int oy, int ox, int out_off) { Job j{ntaps, tap_begin, 0, oy, ox, out_off}; return j; };
  bool custom_out = false;
  switch (d->kind) {
    case LDIC_CONV_S2_5x5_P12:
    case LDIC_CONV_S2_5x5_P2: {
      if ((d->H & 1) || (d->W & 1)) return fail(LDIC_EINVAL, "conv s2: H and W must be even");
      const int pad = d->kind == LDIC_CONV_S2_5x5_P12 ? 1 : 2;
      L->mode = 1; L->k = 5;
      L->Ho = d->H / 2; L->Wo = d->W / 2; L->Wg = L->Wo; L->Hg = L->Ho;
      for (int ky = 0; ky < 5; ++ky) for (int kx = 0; kx < 5; ++kx) {
        int off = kx - pad;
        int t = add_tap(off >> 1, ky - pad, off & 1);
        L->tap_ky[t][0] = ky; L->tap_kx[t][0] = kx;
      }
      L->jobs[0] = job(25, 0, 0, 0, 0);
      break;
    }
    case LDIC_CONV_S1_3x3_P1: {
      L->mode = 0; L->k = 3;
      for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
        int t = add_tap(kx - 1, ky - 1, 0);
        L->tap_ky[t][0] = ky; L->tap_kx[t][0] = kx;
      }
      L->jobs[0] = job(9, 0, 0, 0, 0);
      break;
    }
    case LDIC_CONV_1x1: {
      L->mode = 0; L->k = 1;
      int t = add_tap(0, 0, 0);
      L->tap_ky[t][0] = 0; L->tap_kx[t][0] = 0;
      L->jobs[0] = job(1, 0, 0, 0, 0);
      break;
    }
    case LDIC_CONV_FIRST_5x5S2: {   // x is the NCHW fp32 image; K = (ky*5+kx)*3 + c, 75 padded to 128 in the packed weights
      if (d->Cin != 3 || d->Cin_pad != 128) return fail(LDIC_EINVAL, "first conv: Cin must be 3 and Cin_pad 128");
      if (d->aux0 != 0 && d->aux0 != 1) return fail(LDIC_EINVAL, "first conv: aux0 selects the image type (0 fp32 in [-1,1], 1 uint8 levels)");
      if ((d->H & 1) || (d->W & 3)) return fail(LDIC_EINVAL, "first conv: H must be even and W a multiple of 4");
      L->mode = 0; L->k = 5; L->cin_map = 2;
      L->Ho = d->H / 2; L->Wo = d->W / 2; L->Wg = L->Wo; L->Hg = L->Ho;
      int t = add_tap(0, 0, 0, 0, 0, 2);
      L->tap_ky[t][0] = 0; L->tap_kx[t][0] = 0;
      L->jobs[0] = job(1, 0, 0, 0, 0);
      break;
    }
    case LDIC_DECONV_GS_5x5:
    case LDIC_DECONV_HS_5x5: {
      L->mode = 0; L->k = 5; L->transposed = 1;
      L->Ho = 2 * d->H; L->Wo = 2 * d->W; L->sy = L->sx = 2; L->njobs = 4;
      auto kk = d->kind == LDIC_DECONV_GS_5x5 ? gs_k : hs_k;
      for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px) {
        Job jb = job(0, L->ntaps_total, py, px, 0);
        for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
          int ky = kk(py, dy), kx = kk(px, dx);
          if (ky < 0 || kx < 0) continue;
          int t = add_tap(dx, dy, 0);
          L->tap_ky[t][0] = ky; L->tap_kx[t][0] = kx;
          jb.ntaps++;
        }
        L->jobs[py * 2 + px] = jb;
      }
      break;
    }
    case LDIC_DECONV_S1_3x3: {   // o = i - 1 + k  ->  input offset d = 1 - k
      L->mode = 0; L->k = 3; L->transposed = 1;
      for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
        int t = add_tap(dx, dy, 0);
        L->tap_ky[t][0] = 1 - dy; L->tap_kx[t][0] = 1 - dx;
      }
      L->jobs[0] = job(9, 0, 0, 0, 0);
      break;
    }
    case LDIC_DECONV_GS_5x5_MERGED: {
      L->mode = 0; L->k = 5; L->transposed = 1;
      L->Ho = 2 * d->H; L->Wo = 2 * d->W; L->sy = L->sx = 2;
      L->ngroups = 4;
      for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
        int t = add_tap(dx, dy, 0);
        for (int g = 0; g < 4; ++g) {
          int ky = gs_k(g >> 1, dy), kx = gs_k(g & 1, dx);
          if (ky >= 0 && kx >= 0) { L->tap_ky[t][g] = ky; L->tap_kx[t][g] = kx; }
        }
      }
      L->jobs[0] = job(9, 0, 0, 0, 0);
      break;
    }
    // ---- context model (model/net.py:289-319), one latent position = one 4x4 patch -------------
    case LDIC_CTX_CONV1: {
      // x = [B,h,w, y_round(N) | h2(N)] ; output cell (i,j) of the patch at (y,x) is a 3x3 conv over
      // patch cells (i+di, j+dj) in [0,3]^2 = image pixels (y+i+di-3, x+j+dj-2) (zero outside the
      // image: BlockSample pads with zeros, model/net.py:238); the y sampler masks cells (3,2),(3,3)
      // (:227-230) -> those taps contract over the h2 channels only.
      // aux1 = M | shear << 8 | x_org << 12 | w_in << 16.  shear: the input image is stored sheared by `shear` columns per
      // row (column offset + shear * row offset); w_in > 0: the input image is w_in columns wide and the output grid (d->W
      // columns) starts at its column x_org -- the wavefront decoder computes ONE column of a 10-column band.
      // aux0 = N | pitch << 12: pitch > 0 = the band is a view into a wider image with `pitch` pixels per row.
      const int N = d->aux0 & 0xfff, pitch = d->aux0 >> 12;
      const int M = d->aux1 & 0xff, shear = (d->aux1 >> 8) & 0xf, x_org = (d->aux1 >> 12) & 0xf, w_in = d->aux1 >> 16;
      if (w_in) {
        if (w_in < x_org + d->W) return fail(LDIC_EINVAL, "ctx conv1: the output columns must lie inside the input band");
        L->vW = w_in;
        if (pitch) {
          if (pitch < w_in) return fail(LDIC_EINVAL, "ctx conv1: row pitch smaller than the band");
          L->vPitch = pitch;
        }
      } else if (pitch) {
        return fail(LDIC_EINVAL, "ctx conv1: a row pitch needs a band (w_in)");
      }
      if (N <= 0 || N % 64 || M < 0 || M >= N || d->Cin != 2 * N - M || d->Cin_pad != 2 * N || d->Cout != N || d->Cout_pad != N)
        return fail(LDIC_EINVAL, "ctx conv1: need aux0=N (multiple of 64), aux1=M, Cin=2N-M, Cin_pad=2N, Cout=Cout_pad=N");
      L->mode = 0; L->k = 3; L->cin_map = 1; L->map_N = N; L->map_M = M; L->njobs = 16;
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        Job jb = job(0, L->ntaps_total, 0, 0, (i * 4 + j) * N);
        for (int di = -1; di <= 1; ++di) for (int dj = -1; dj <= 1; ++dj) {
          const int ci = i + di, cj = j + dj;
          if (ci < 0 || ci > 3 || cj < 0 || cj > 3) continue;
          const bool masked = (ci == 3 && cj >= 2);
          const int tdx = cj - 2 + shear * (ci - 3) + x_org;
          int t = masked ? add_tap(tdx, ci - 3, 0, N, N, N / 64) : add_tap(tdx, ci - 3, 0, 0, 0, 2 * N / 64);
          L->tap_ky[t][0] = di + 1; L->tap_kx[t][0] = dj + 1;
          jb.ntaps++;
        }
        L->jobs[i * 4 + j] = jb;
      }
      L->oB = d->B * d->H * d->W; L->Ho = 4; L->Wo = 4;
      L->out_sX = 16LL * N; L->out_sY = (long long)d->W * 16 * N; L->out_sN = (long long)d->H * d->W * 16 * N;
      custom_out = true;
      break;
    }
    case LDIC_CTX_CONV2:     // Conv2d(3, s2, p1) on the 4x4 patch -> 2x2          (model/net.py:297)
    case LDIC_CTX_CONV3:     // Conv2d(3, s1, p1) on the 2x2 patch                 (model/net.py:299)
    case LDIC_CTX_FC: {      // Linear over the (c,h,w)-flattened 2x2 patch         (model/net.py:302,311-312)
      const int N = d->Cin;
      const int cells_in = d->kind == LDIC_CTX_CONV2 ? 16 : 4;
      const int side = d->kind == LDIC_CTX_CONV2 ? 4 : 2;
      if (N % 64 || d->Cin_pad != N || d->H != side || d->W != side)
        return fail(LDIC_EINVAL, "ctx conv2/conv3/fc: Cin=Cin_pad multiple of 64, H=W=%d", side);
      L->mode = 0; L->k = d->kind == LDIC_CTX_FC ? 2 : 3;
      L->vC = (long long)cells_in * N; L->vW = d->B; L->vH = 1; L->vN = 1;
      L->Wg = d->B; L->Hg = 1; L->Bg = 1; L->Kw = N;
      L->oB = d->B;
      if (d->kind == LDIC_CTX_FC) {
        L->njobs = 2; L->nbias = 2; L->Ho = 1; L->Wo = 2;
        for (int g = 0; g < 2; ++g) {
          Job jb = job(4, L->ntaps_total, 0, 0, g * d->Cout_pad);
          for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
            int t = add_tap(0, 0, 0, (a * 2 + b) * N, 0, N / 64);
            L->tap_ky[t][0] = a; L->tap_kx[t][0] = b; L->tap_co_base[t] = (short)(g * d->Cout);
          }
          L->jobs[g] = jb;
        }
        L->out_sX = 2LL * d->Cout_pad;
      } else {
        L->njobs = 4; L->Ho = 2; L->Wo = 2;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
          Job jb = job(0, L->ntaps_total, 0, 0, (a * 2 + b) * d->Cout_pad);
          for (int di = -1; di <= 1; ++di) for (int dj = -1; dj <= 1; ++dj) {
            int ci, cj;
            if (d->kind == LDIC_CTX_CONV2) { ci = 2 * a + di; cj = 2 * b + dj; }      // stride 2, pad 1
            else { ci = a + di; cj = b + dj; }                                           // stride 1, pad 1
            if (ci < 0 || ci >= side || cj < 0 || cj >= side) continue;
            int t = add_tap(0, 0, 0, (ci * side + cj) * N, 0, N / 64);
            L->tap_ky[t][0] = di + 1; L->tap_kx[t][0] = dj + 1;
            jb.ntaps++;
          }
          L->jobs[a * 2 + b] = jb;
        }
        L->out_sX = 4LL * d->Cout_pad;
      }
      L->out_sY = L->out_sN = 0;
      custom_out = true;
      break;
    }
    default:
      return fail(LDIC_EINVAL, "conv: unknown kind %d", d->kind);
  }
  L->Np = L->ngroups * L->Cg;
  L->Cs = L->Cg;
  if (L->Np != 64 && L->Np != 128 && L->Np != 192 && L->Np != 256 && L->Np != 384)
    return fail(LDIC_EINVAL, "conv: accumulator width %d (groups %d x Cout_pad %d) must be 64/128/192/256/384", L->Np,
                L->ngroups, L->Cg);
  if (!custom_out) {
    L->out_sX = L->Cg; L->out_sY = (long long)L->Wo * L->Cg; L->out_sN = (long long)L->Ho * L->Wo * L->Cg;
  }
  for (int j = 0; j < L->njobs; ++j) {
    int n = 0;
    for (int t = 0; t < L->jobs[j].ntaps; ++t) n += L->taps[L->jobs[j].tap_begin + t].nkc;
    L->jobs[j].nkb = n;
  }
  return LDIC_OK;
}

void choose_tile(int Wg, int Hg, int B, int* TW, int* TH, int* TN) {
  double best = 1e30;
  int bw = 128, bh = 1, bn = 1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      int tn = 128 / (tw * th);
      long long tiles = (long long)((Wg + tw - 1) / tw) * ((Hg + th - 1) / th) * ((B + tn - 1) / tn);
      if ((double)tiles < best - 1e-9) { best = (double)tiles; bw = tw; bh = th; bn = tn; }
    }
  }
  *TW = bw; *TH = bh; *TN = bn;
}

// ---------------------------------------------------------------------------------
// Launch plans.  Everything a launch needs that depends only on (layer descriptor, tensor addresses) -- the tap /
// tile tables, the three TMA descriptors, the kernel variant, grid and shared-memory size -- is built once and kept
// in a small cache, so a repeated eager call costs a hash lookup and one cudaLaunchKernelEx (no getenv, no
// cuTensorMapEncodeTiled, no table construction on the hot path).
// ---------------------------------------------------------------------------------
enum PlanKernel { PK_PAIR = 0, PK_W3 = 1, PK_WIDE = 2, PK_FIRST = 3, PK_FIRST_U8 = 4, PK_PAIR_RES = 5, PK_PAIR_TF = 6 };
struct Plan {
  ConvParams P;
  CUtensorMap a, w, g, y;   // y: output map of the first layer's TMA stores (P.tma_store)
  int kernel, np, grid;
  int epi_warps;           // first-layer kernel: 12 epilogue warps at 192 channels (three per TMEM lane quadrant), else 8
  size_t smem;
  int kind;
};

// per-device lazily initialised launch state of one kernel instantiation (guarded by g_init_mu)
struct KernelState { bool attr_set[kMaxDevices]; int max_clusters[kMaxDevices]; };

template <typename K>
int prepare_kernel(K kern, KernelState& st, bool cluster, int* max_clusters) {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return fail(LDIC_ECUDA, "conv: bad current device %d", dev);
  std::lock_guard<std::mutex> init_lock(g_init_mu);
  if (!st.attr_set[dev]) {
    LDIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (cluster) {
      const int sms = num_sms();
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.blockDim = dim3(kThreads); cfg.gridDim = dim3(sms); cfg.dynamicSmemBytes = 227 * 1024 - 1024;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sms / 2 - 4; }
      st.max_clusters[dev] = n < sms / 2 ? n : sms / 2;
    }
    st.attr_set[dev] = true;
  }
  if (max_clusters) *max_clusters = st.max_clusters[dev];
  return LDIC_OK;
}

template <typename K>
int launch_cluster2(K kern, const char* name, const Plan& pl, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads); cfg.gridDim = dim3(pl.grid); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, pl.a, pl.w, pl.g, pl.P);
  if (e != cudaSuccess) return fail(LDIC_ECUDA, "%s launch: %s", name, cudaGetErrorString(e));
  return check_launch(name);
}

template <int NP, bool W3, bool RES = false, bool TF = false>
int plan_pair(Plan* pl) {
  static KernelState ks;
  int mc = 0;
  int rc = prepare_kernel(conv_tc2_kernel<NP, W3, RES, TF>, ks, true, &mc);
  if (rc) return rc;
  const int total_super = pl->P.super_per_job * pl->P.njobs;
  pl->grid = 2 * (total_super < mc ? total_super : mc);
  return LDIC_OK;
}
template <int NP>
int plan_wide(Plan* pl) {
  static KernelState ks;
  int mc = 0;
  int rc = prepare_kernel(conv_wide_kernel<NP>, ks, true, &mc);
  if (rc) return rc;
  const int total_super = pl->P.super_per_job * pl->P.njobs;
  pl->grid = 2 * (total_super < mc ? total_super : mc);
  return LDIC_OK;
}
template <int NP, bool U8, int EW>
int plan_first(Plan* pl) {
  static KernelState ks;
  int rc = prepare_kernel(conv_first_kernel<NP, U8, EW>, ks, false, nullptr);
  if (rc) return rc;
  const int sms = num_sms();
  pl->grid = pl->P.total_tiles < sms ? pl->P.total_tiles : sms;
  return LDIC_OK;
}

int finish_plan(Plan* pl) {
  switch (pl->kernel) {
    case PK_PAIR:
      switch (pl->np) {
        case 64: return plan_pair<64, false>(pl);
        case 128: return plan_pair<128, false>(pl);
        case 192: return plan_pair<192, false>(pl);
        case 256: return plan_pair<256, false>(pl);
      }
      break;
    case PK_PAIR_RES:
      switch (pl->np) {
        case 64: return plan_pair<64, false, true>(pl);
        case 128: return plan_pair<128, false, true>(pl);
        case 192: return plan_pair<192, false, true>(pl);
      }
      break;
    case PK_PAIR_TF:
      switch (pl->np) {
        case 128: return plan_pair<128, false, false, true>(pl);
        case 192: return plan_pair<192, false, false, true>(pl);
      }
      break;
    case PK_W3: return plan_pair<192, true>(pl);
    case PK_WIDE: return plan_wide<384>(pl);
    case PK_FIRST:
      switch (pl->np) {
        case 64: return plan_first<64, false, 8>(pl);
        case 128: return plan_first<128, false, 8>(pl);
        case 192: return pl->epi_warps == 12 ? plan_first<192, false, 12>(pl) : plan_first<192, false, 8>(pl);
      }
      break;
    case PK_FIRST_U8:
      switch (pl->np) {
        case 64: return plan_first<64, true, 8>(pl);
        case 128: return plan_first<128, true, 8>(pl);
        case 192: return pl->epi_warps == 12 ? plan_first<192, true, 12>(pl) : plan_first<192, true, 8>(pl);
      }
      break;
  }
  return fail(LDIC_EINVAL, "conv: no kernel for variant %d with %d accumulator columns", pl->kernel, pl->np);
}

int launch_plan(const Plan& pl, cudaStream_t st) {
  switch (pl.kernel) {
    case PK_PAIR:
      switch (pl.np) {
        case 64: return launch_cluster2(conv_tc2_kernel<64, false>, "conv_tc2_kernel", pl, st);
        case 128: return launch_cluster2(conv_tc2_kernel<128, false>, "conv_tc2_kernel", pl, st);
        case 192: return launch_cluster2(conv_tc2_kernel<192, false>, "conv_tc2_kernel", pl, st);
        case 256: return launch_cluster2(conv_tc2_kernel<256, false>, "conv_tc2_kernel", pl, st);
      }
      break;
    case PK_PAIR_RES:
      switch (pl.np) {
        case 64: return launch_cluster2(conv_tc2_kernel<64, false, true>, "conv_tc2_kernel(+residual)", pl, st);
        case 128: return launch_cluster2(conv_tc2_kernel<128, false, true>, "conv_tc2_kernel(+residual)", pl, st);
        case 192: return launch_cluster2(conv_tc2_kernel<192, false, true>, "conv_tc2_kernel(+residual)", pl, st);
      }
      break;
    case PK_PAIR_TF:
      switch (pl.np) {
        case 128: return launch_cluster2(conv_tc2_kernel<128, false, false, true>, "conv_tc2_kernel(tf32)", pl, st);
        case 192: return launch_cluster2(conv_tc2_kernel<192, false, false, true>, "conv_tc2_kernel(tf32)", pl, st);
      }
      break;
    case PK_W3: return launch_cluster2(conv_tc2_kernel<192, true>, "conv_tc2_kernel(wide tail)", pl, st);
    case PK_WIDE: return launch_cluster2(conv_wide_kernel<384>, "conv_wide_kernel", pl, st);
    case PK_FIRST:
    case PK_FIRST_U8: {
      const bool u8 = pl.kernel == PK_FIRST_U8;
      switch (pl.np) {
#define LDIC_LAUNCH_FIRST(NPV, U8V, EWV) \
  conv_first_kernel<NPV, U8V, EWV><<<pl.grid, EWV * 32 + 128, pl.smem, st>>>(pl.a, pl.w, pl.g, pl.y, pl.P)
        case 64: if (u8) LDIC_LAUNCH_FIRST(64, true, 8); else LDIC_LAUNCH_FIRST(64, false, 8); break;
        case 128: if (u8) LDIC_LAUNCH_FIRST(128, true, 8); else LDIC_LAUNCH_FIRST(128, false, 8); break;
        case 192:
          if (pl.epi_warps == 12) { if (u8) LDIC_LAUNCH_FIRST(192, true, 12); else LDIC_LAUNCH_FIRST(192, false, 12); }
          else { if (u8) LDIC_LAUNCH_FIRST(192, true, 8); else LDIC_LAUNCH_FIRST(192, false, 8); }
          break;
#undef LDIC_LAUNCH_FIRST
        default: return fail(LDIC_EINVAL, "first conv: unsupported Np %d", pl.np);
      }
      return check_launch("conv_first_kernel");
    }
  }
  return fail(LDIC_EINVAL, "conv: bad plan");
}

// generic weight packer: Wp[t][n][k]
struct PackTable {
  int ntaps, Np, Cg, ngroups, Cin, Cout, Kw, cin_offset, k, transposed, cin_map, map_N, map_M, nbias;
  signed char ky[kMaxTaps][4], kx[kMaxTaps][4];
  short co_base[kMaxTaps];
};
template <typename TW>
__global__ void k_pack_weights(const float* __restrict__ w, const float* __restrict__ bias, PackTable T,
                               TW* __restrict__ wp, float* __restrict__ bp) {
  const long long total = (long long)T.ntaps * T.Np * T.Kw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int kc = (int)(i % T.Kw);
    long long r = i / T.Kw;
    int n = (int)(r % T.Np);
    int t = (int)(r / T.Np);
    int g = n / T.Cg, co = n - g * T.Cg;
    int ci;
    if (T.cin_map == 1) ci = kc < T.map_N ? (kc >= T.map_M ? kc - T.map_M : -1) : (T.map_N - T.map_M) + (kc - T.map_N);
    else ci = kc - T.cin_offset;
    float v = 0.f;
    int ky = T.ky[t][g], kx = T.kx[t][g];
    if (T.cin_map == 2) {                    // first layer: k = (ky*5 + kx)*Cin + ci
      const int tap = kc / T.Cin;
      ci = kc < 25 * T.Cin ? kc - tap * T.Cin : -1;
      ky = tap / 5; kx = tap - ky * 5;
    }
    if (ky >= 0 && co < T.Cout && ci >= 0 && ci < T.Cin) {
      const int cot = co + T.co_base[t];
      const int Ctot = T.Cout * T.nbias;     // total logical output channels of the weight tensor
      long long idx = T.transposed ? ((((long long)ci * Ctot + cot) * T.k + ky) * T.k + kx)
                                   : ((((long long)cot * T.Cin + ci) * T.k + ky) * T.k + kx);
      v = w[idx];
    }
    if constexpr (std::is_same<TW, float>::value) wp[i] = __uint_as_float(to_tf32(v));   // TF32 parity mode
    else wp[i] = __float2bfloat16_rn(v);
  }
  if (bp) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < T.Np * T.nbias; n += gridDim.x * blockDim.x) {
      int jb = n / T.Np, co = (n % T.Np) % T.Cg;
      bp[n] = (bias && co < T.Cout) ? bias[jb * T.Cout + co] : 0.f;
    }
  }
}

// SM partition (LdicConvDesc::sm_limit): the kernels are persistent with one CTA (or one CTA of a pair) per SM and walk
// their tiles with a grid stride, so capping the grid caps the SMs the launch occupies; the rest stay free for a
// kernel of another stream.
void apply_sm_limit(const LdicConvDesc* d, Plan* pl) {
  if (d->sm_limit == 0) return;
  const int sms = num_sms();
  int avail = d->sm_limit > 0 ? d->sm_limit : sms + d->sm_limit;
  if (avail > sms) avail = sms;
  if (avail < 2) avail = 2;
  if (pl->kernel != PK_FIRST && pl->kernel != PK_FIRST_U8) avail &= ~1;       // whole CTA pairs
  if (pl->grid > avail) pl->grid = avail;
}

int build_plan_first(const LdicConvDesc* d, const Layer& L, const void* x, const void* w_packed, const float* bias_packed,
                     const void* gamma_bf16, const float* beta_tiled, void* y, Plan* pl) {
  const bool gdn = d->act == LDIC_ACT_GDN || d->act == LDIC_ACT_IGDN;
  const bool u8 = d->aux0 == 1;
  if (L.Np > 192) return fail(LDIC_EINVAL, "first conv: at most 192 output channels (weights and gamma stay resident in shared memory)");
  if (u8 && (d->W % 16)) return fail(LDIC_EINVAL, "first conv (uint8 image): W must be a multiple of 16");
  ConvParams& P = pl->P;
  memset(&P, 0, sizeof(P));
  P.mode = 0; P.TW = kFirstTW; P.TH = kFirstTH; P.TN = 1; P.tw_shift = 6; P.th_shift = 1;
  P.tiles_x = (L.Wg + P.TW - 1) / P.TW; P.tiles_y = (L.Hg + P.TH - 1) / P.TH; P.tiles_n = L.Bg;
  P.tiles_per_job = P.tiles_x * P.tiles_y * P.tiles_n; P.njobs = 1; P.total_tiles = P.tiles_per_job;
  P.Wg = L.Wg; P.Hg = L.Hg; P.B = L.Bg;
  P.gdn_kblocks = gdn ? L.Np / 64 : 0;
  P.act = d->act; P.out_f32 = d->out_f32;
  P.ngroups = 1; P.Cg = L.Cg; P.sy = P.sx = 1; P.nbias = 1;
  P.out_sX = L.out_sX; P.out_sY = L.out_sY; P.out_sN = L.out_sN;
  P.jobs[0] = L.jobs[0];
  P.bias = bias_packed; P.beta = beta_tiled; P.out = y;
  P.dbg_nostore = tuning().debug_nostore;
  {
    const int fi = tuning().first_insert;
    P.gdn_insert = gdn ? (fi < 0 ? 0 : (fi > 2 ? 2 : fi)) : 2;    // see conv_first_kernel: ring order of the x^2 slots
  }
  P.tail_H = d->H; P.tail_W = d->W;             // image size: the uint8 patch builder masks out-of-image taps itself
  const int fixed = (2 + L.Np / 64) * L.Np * kBlockK * 2 + 2 * kRawSlot + 256 + 2 * L.Np * 4 + 1024;
  int S = (227 * 1024 - fixed) / kATileBytes;
  if (S > kMaxStages) S = kMaxStages;
  if (S < P.gdn_kblocks + 2) return fail(LDIC_EINVAL, "first conv: not enough shared memory for the operand ring");
  P.stages = S;
  pl->smem = (size_t)(2 + L.Np / 64) * L.Np * kBlockK * 2 + (size_t)S * kATileBytes + 2 * kRawSlot + 256 +
             2 * L.Np * sizeof(float) + 1024 /*align*/;
  if (pl->smem > 227 * 1024) return fail(LDIC_EINVAL, "first conv: shared memory budget exceeded (%zu)", pl->smem);
  int rc;
  {
    const cuuint64_t es = u8 ? 1 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)d->W, (cuuint64_t)d->H, 3, (cuuint64_t)d->B};
    cuuint64_t str[3] = {(cuuint64_t)d->W * es, (cuuint64_t)d->H * d->W * es, (cuuint64_t)3 * d->H * d->W * es};
    cuuint32_t box[4] = {(cuuint32_t)(u8 ? kRawW8 : kRawW), (cuuint32_t)kRawH, 3, 1};
    if ((rc = encode_map(&pl->a, x, 4, dims, str, box, nullptr, u8 ? MAP_U8 : MAP_F32))) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)L.Kw, (cuuint64_t)L.Np};
    cuuint64_t str[1] = {(cuuint64_t)L.Kw * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)L.Np};
    if ((rc = encode_map(&pl->w, w_packed, 2, dims, str, box))) return rc;
  }
  if (gdn) {
    cuuint64_t dims[2] = {(cuuint64_t)L.Np, (cuuint64_t)L.Np};
    cuuint64_t str[1] = {(cuuint64_t)L.Np * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)L.Np};
    if ((rc = encode_map(&pl->g, gamma_bf16, 2, dims, str, box))) return rc;
  } else {
    pl->g = pl->w;
  }
  pl->y = pl->w;
  // TMA stores of the bf16 output tile through the tile's own x^2 slots (needs the GDN epilogue's slots: one per 64 channels)
  P.tma_store = gdn && !d->out_f32 && y && tuning().first_tma_store && L.out_sX == L.Np && (((uintptr_t)y) & 127) == 0;
  if (P.tma_store) {
    cuuint64_t dims[4] = {(cuuint64_t)L.Np, (cuuint64_t)L.Wg, (cuuint64_t)L.Hg, (cuuint64_t)L.Bg};
    cuuint64_t str[3] = {(cuuint64_t)L.out_sX * 2, (cuuint64_t)L.out_sY * 2, (cuuint64_t)L.out_sN * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kFirstTW, (cuuint32_t)kFirstTH, 1};
    if ((rc = encode_map(&pl->y, y, 4, dims, str, box))) return rc;
  }
  pl->kernel = u8 ? PK_FIRST_U8 : PK_FIRST;
  pl->np = L.Np;
  pl->epi_warps = (L.Np == 192 && tuning().first_epi != 8) ? 12 : 8;
  if ((rc = finish_plan(pl))) return rc;
  apply_sm_limit(d, pl);
  return LDIC_OK;
}

int build_plan(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
               const void* gamma_bf16, const float* beta_tiled, void* y, const LdicConvTail* tail, Plan* pl,
               const void* residual = nullptr) {
  Layer L;
  int rc = build_layer(d, &L);
  if (rc) return rc;
  pl->kind = d->kind;
  const bool gdn = d->act == LDIC_ACT_GDN || d->act == LDIC_ACT_IGDN;
  if (gdn && (!gamma_bf16 || !beta_tiled)) return fail(LDIC_EINVAL, "conv: GDN epilogue needs gamma_bf16 and beta_tiled");
  if (gdn && L.njobs > 4) return fail(LDIC_EINVAL, "conv: GDN epilogue is not available for the context layers");
  if (((uintptr_t)w_packed) & 15) return fail(LDIC_EINVAL, "conv: packed weights must be 16-byte aligned");
  if (L.Cs % 16) return fail(LDIC_EINVAL, "conv: output channel count must be a multiple of 16");
  if (d->kind == LDIC_CONV_FIRST_5x5S2) {
    if (d->precision) return fail(LDIC_EINVAL, "first conv: TF32 mode goes through the patch matrix (ldic_im2col_5x5s2_f32 + LDIC_CONV_1x1)");
    return build_plan_first(d, L, x, w_packed, bias_packed, gamma_bf16, beta_tiled, y, pl);
  }
  // TF32 parity mode (LdicConvDesc.precision = 1): fp32 activations / weights / gamma.  TMA and the UMMA descriptors
  // only see bytes, so the fp32 tensors are described as bf16 tensors with twice the channels: every K block below is
  // 64 "bf16 columns" = 32 fp32 channels, the tap tables count in those units, the kernel issues kind::tf32 MMAs.
  const bool tf32 = d->precision == 1 || d->precision == 2;      // 2: fp32 outputs kept unrounded (the layer feeding the quantiser)
  if (tf32) {
    if (d->kind != LDIC_CONV_S2_5x5_P12 && d->kind != LDIC_CONV_S2_5x5_P2 && d->kind != LDIC_CONV_S1_3x3_P1 && d->kind != LDIC_CONV_1x1)
      return fail(LDIC_EINVAL, "conv: TF32 mode covers the analysis-side kinds (5x5 s2, 3x3 s1, 1x1)");
    if (!d->out_f32 || tail || residual || (L.Np != 128 && L.Np != 192)) return fail(LDIC_EINVAL, "conv: TF32 mode needs fp32 output and 128 / 192 output channels");
    L.vC *= 2; L.Kw *= 2;
    for (int t = 0; t < L.ntaps_total; ++t) { L.taps[t].nkc = (short)(2 * L.taps[t].nkc); L.taps[t].a_c0 *= 2; L.taps[t].b_c0 *= 2; }
    for (int j = 0; j < L.njobs; ++j) L.jobs[j].nkb *= 2;
  }

  const Tuning& tn = tuning();
  ConvParams& P = pl->P;
  memset(&P, 0, sizeof(P));
  // stride-2 gathers: one TMA box with element strides (2,2) over (W,H) per stage
  const bool strided = L.mode == 1;
  P.mode = strided ? 2 : 0;
  choose_tile(L.Wg, L.Hg, L.Bg, &P.TW, &P.TH, &P.TN);
  auto ilog2 = [](int v) { int s = 0; while ((1 << s) < v) ++s; return s; };
  P.tw_shift = ilog2(P.TW); P.th_shift = ilog2(P.TH); P.cg_shift = ilog2(L.Cg);
  if (L.ngroups > 1 && ((1 << P.cg_shift) != L.Cg || L.Cg < 8)) return fail(LDIC_EINVAL, "conv: merged deconv needs a power-of-two Cout_pad >= 8");
  P.tiles_x = (L.Wg + P.TW - 1) / P.TW;
  P.tiles_y = (L.Hg + P.TH - 1) / P.TH;
  P.tiles_n = (L.Bg + P.TN - 1) / P.TN;
  P.tiles_per_job = P.tiles_x * P.tiles_y * P.tiles_n;
  P.njobs = L.njobs;
  P.total_tiles = P.tiles_per_job * P.njobs;
  P.Wg = L.Wg; P.Hg = L.Hg; P.B = L.Bg;
  P.gdn_kblocks = gdn ? L.Np / (tf32 ? 32 : 64) : 0;
  P.act = d->act; P.out_f32 = d->out_f32; P.tf_round_out = d->precision == 1;
  P.ngroups = L.ngroups; P.Cg = L.Cg; P.sy = L.sy; P.sx = L.sx; P.nbias = L.nbias;
  P.out_sX = L.out_sX; P.out_sY = L.out_sY; P.out_sN = L.out_sN;
  for (int j = 0; j < kMaxJobs; ++j) P.jobs[j] = L.jobs[j];
  for (int t = 0; t < kMaxTaps; ++t) P.taps[t] = L.taps[t];
  if (strided) for (int t = 0; t < L.ntaps_total; ++t) { P.taps[t].dx = (short)(2 * L.taps[t].dx + L.taps[t].px); P.taps[t].px = 0; }
  P.bias = bias_packed; P.beta = beta_tiled; P.out = y;
  P.dbg_nostore = tn.debug_nostore;
  P.gdn_insert = tn.gdn_insert >= 1 && tn.gdn_insert <= 16 ? tn.gdn_insert : kGdnInsertDefault;
  if (tail) {
    P.tail_x = tail->x_nchw; P.tail_w = tail->w; P.tail_xo = tail->x_tilde_nchw; P.tail_sq = tail->sq_err;
    P.tail_H = tail->H; P.tail_W = tail->W; P.tail_u8 = tail->x_is_u8; P.tail_tanh = tail->tanh_out;
  }

  // ---- wide-N form of the merged last deconv (epilogue_w3): the three dx taps ride side by side in N = 3 * Np, the dy
  // taps are three stride-1 gathers of a 32 x 4 pixel tile (tiles overlap by one halo pixel per side in x) ----
  const bool wide3 = d->kind == LDIC_DECONV_GS_5x5_MERGED && gdn && L.Np == 64 && L.ngroups == 4 && L.Cg == 16 && L.njobs == 1 &&
                     L.ntaps_total == 9 && L.nbias == 1 && tn.tail_wide;
  if (wide3) {
    P.TW = 32; P.TH = 4; P.TN = 1; P.tw_shift = 5; P.th_shift = 2; P.x_ovl = 2;
    P.tiles_x = (L.Wg + P.TW - P.x_ovl - 1) / (P.TW - P.x_ovl);
    P.tiles_y = (L.Hg + P.TH - 1) / P.TH;
    P.tiles_n = L.Bg;
    P.tiles_per_job = P.tiles_x * P.tiles_y * P.tiles_n;
    P.total_tiles = P.tiles_per_job;
    const short nkc = L.taps[0].nkc;
    for (int t = 0; t < 3; ++t) {            // packed weights are [tap = (dy+1)*3 + (dx+1)][Np][K]: rows t*3*Np .. +3*Np = one dy
      Tap tp; memset(&tp, 0, sizeof(tp));
      tp.dx = 0; tp.dy = (short)(t - 1); tp.nkc = nkc;
      P.taps[t] = tp;
    }
    P.jobs[0].ntaps = 3; P.jobs[0].tap_begin = 0; P.jobs[0].nkb = 3 * nkc;
  }
  if (tail && L.Np > 128) return fail(LDIC_EINVAL, "conv tail: the merged last deconv must have at most 128 accumulator columns");
  P.super_per_job = (P.tiles_per_job + 1) / 2;

  const bool wide = L.Np > 256;              // single-buffered 384-column accumulator, chunked gamma contraction
  const int NpK = wide3 ? 3 * L.Np : L.Np;   // accumulator columns of the kernel
  if (wide) {
    if (L.Np != 384) return fail(LDIC_EINVAL, "conv: accumulator width %d not supported (64/128/192/256/384)", L.Np);
    if (L.ngroups != 1) return fail(LDIC_EINVAL, "conv: merged sub-pixel phases need at most 256 accumulator columns");
    const int bslot = (L.Np / 2) * kBlockK * 2;
    const int budget = 227 * 1024 - 1024 - 512 - (L.nbias + 1) * L.Np * 4 - 64;
    int SA = gdn ? L.Np / 64 + 2 : 5;        // the GDN phase holds Np/64 x^2 tiles in the A ring
    int SB = (budget - SA * kATileBytes) / bslot;
    while (SB < 3 && SA > (gdn ? L.Np / 64 + 1 : 3)) { --SA; SB = (budget - SA * kATileBytes) / bslot; }
    if (SB > kMaxStages) SB = kMaxStages;
    if (SA > kMaxStages) SA = kMaxStages;
    if (SB < 2) return fail(LDIC_EINVAL, "conv (wide): shared memory budget exceeded");
    P.SA = SA; P.SB = SB;
    pl->smem = (size_t)SA * kATileBytes + (size_t)SB * bslot + 1024 + 512 + (size_t)(L.nbias + 1) * L.Np * sizeof(float) + 64;
  } else {
    const int stage2 = kATileBytes + (NpK / 2) * kBlockK * 2;
    int st2 = (227 * 1024 - 2048 - (L.nbias + 1) * NpK * 4) / stage2;
    if (st2 > kMaxStages) st2 = kMaxStages;
    if (tn.stages_cap >= 2 && tn.stages_cap < st2) st2 = tn.stages_cap;
    P.stages = st2;
    if (gdn && st2 < P.gdn_kblocks + 1) return fail(LDIC_EINVAL, "conv: not enough pipeline stages for the GDN epilogue");
    pl->smem = (size_t)st2 * stage2 + 1024 + 256 + (size_t)(L.nbias + 1) * NpK * sizeof(float);
  }
  if (pl->smem > 227 * 1024) return fail(LDIC_EINVAL, "conv: shared memory budget exceeded (%zu)", pl->smem);

  const cuuint64_t C = (cuuint64_t)L.vC;
  {
    const cuuint64_t pitch = (cuuint64_t)(L.vPitch ? L.vPitch : L.vW);
    cuuint64_t dims[4] = {C, (cuuint64_t)L.vW, (cuuint64_t)L.vH, (cuuint64_t)L.vN};
    cuuint64_t str[3] = {C * 2, pitch * C * 2, (cuuint64_t)L.vH * pitch * C * 2};
    if (strided) {
      cuuint32_t box[4] = {64, (cuuint32_t)(2 * P.TW), (cuuint32_t)(2 * P.TH), (cuuint32_t)P.TN};
      cuuint32_t es[4] = {1, 2, 2, 1};
      if ((rc = encode_map(&pl->a, x, 4, dims, str, box, es))) return rc;
    } else {
      cuuint32_t box[4] = {64, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
      if ((rc = encode_map(&pl->a, x, 4, dims, str, box))) return rc;
    }
  }
  {
    // each CTA of a pair loads half of the weight rows (wide kernel: half of each column half = a quarter per box)
    cuuint64_t dims[2] = {(cuuint64_t)L.Kw, (cuuint64_t)L.ntaps_total * L.Np};
    cuuint64_t str[1] = {(cuuint64_t)L.Kw * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(wide ? L.Np / 4 : NpK / 2)};
    if ((rc = encode_map(&pl->w, w_packed, 2, dims, str, box))) return rc;
  }
  if (gdn) {
    const cuuint64_t gcols = (cuuint64_t)L.Np * (tf32 ? 2 : 1);
    cuuint64_t dims[2] = {gcols, (cuuint64_t)L.Np};
    cuuint64_t str[1] = {gcols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(wide ? kWideNC / 2 : L.Np / 2)};
    if ((rc = encode_map(&pl->g, gamma_bf16, 2, dims, str, box))) return rc;
  } else {
    pl->g = pl->w;
  }
  pl->kernel = wide ? PK_WIDE : (wide3 ? PK_W3 : (tf32 ? PK_PAIR_TF : PK_PAIR));
  if (residual) {
    if (gdn || wide || wide3 || L.ngroups != 1 || L.njobs != 1 || L.Np > 192 || tail || (((uintptr_t)residual) & 31))
      return fail(LDIC_EINVAL, "conv: the fused residual needs a plain (no GDN, one job) layer of at most 192 output channels and a 32-byte aligned residual");
    pl->kernel = PK_PAIR_RES;
    P.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  }
  pl->np = L.Np;
  if ((rc = finish_plan(pl))) return rc;
  apply_sm_limit(d, pl);
  return LDIC_OK;
}

// ---- plan cache -----------------------------------------------------------------------------------
// The key holds what a plan depends on structurally: the layer, its (long-lived) parameter tensors and the device.
// The activation addresses x / y and the tail's pointers change from call to call (torch's allocator hands out
// different blocks) and are patched into a copy of the cached plan: cuTensorMapReplaceAddress for the one TMA
// descriptor that embeds x, plain parameter fields for the rest.
struct PlanKey {
  LdicConvDesc d;
  const void *w, *bias, *gamma, *beta;
  int has_tail, tail_H, tail_W, tail_u8, tail_tanh, has_y, has_res, device;
  unsigned epoch;
  bool operator==(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) == 0; }
};
struct PlanKeyHash {
  size_t operator()(const PlanKey& k) const {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(PlanKey); ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::mutex g_plan_mu;
std::unordered_map<PlanKey, std::shared_ptr<Plan>, PlanKeyHash> g_plans;
constexpr size_t kMaxPlans = 1024;

unsigned long long* g_timeout_host = nullptr;
int ensure_timeout_report() {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return LDIC_OK;
  std::lock_guard<std::mutex> init_lock(g_init_mu);
  static bool done[kMaxDevices] = {};
  if (done[dev]) return LDIC_OK;
  done[dev] = true;                     // diagnostics only: on any failure run without the report buffer
  unsigned long long* h = g_timeout_host;
  if (!h) {
    if (cudaHostAlloc((void**)&h, 8 * sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return LDIC_OK;
    }
    memset(h, 0, 8 * sizeof(unsigned long long));
    g_timeout_host = h;
  }
  unsigned long long* dptr = nullptr;
  if (cudaHostGetDevicePointer((void**)&dptr, h, 0) != cudaSuccess) { cudaGetLastError(); return LDIC_OK; }
  if (cudaMemcpyToSymbol(g_timeout_report, &dptr, sizeof(dptr)) != cudaSuccess) cudaGetLastError();
  return LDIC_OK;
}

}  // namespace

extern "C" int ldic_debug_last_timeout(unsigned long long* out5) {
  if (!g_timeout_host || !out5) return 0;
  for (int i = 0; i < 5; ++i) out5[i] = g_timeout_host[i];
  return g_timeout_host[0] != 0;
}

extern "C" int ldic_conv_n_cols(const LdicConvDesc* d) {
  Layer L;
  if (build_layer(d, &L) != LDIC_OK) return -1;
  return L.Np;
}
extern "C" int ldic_conv_bias_elems(const LdicConvDesc* d) {
  Layer L;
  if (build_layer(d, &L) != LDIC_OK) return -1;
  return L.Np * L.nbias;
}
extern "C" long long ldic_conv_weight_elems(const LdicConvDesc* d) {
  Layer L;
  if (build_layer(d, &L) != LDIC_OK) return -1;
  return (long long)L.ntaps_total * L.Np * L.Kw;
}
extern "C" void ldic_conv_out_shape(const LdicConvDesc* d, int* Ho, int* Wo) {
  Layer L;
  if (build_layer(d, &L) != LDIC_OK) { *Ho = *Wo = -1; return; }
  *Ho = L.Ho; *Wo = L.Wo;
}
extern "C" int ldic_conv_out_dims(const LdicConvDesc* d, int* dims4) {
  Layer L;
  int rc = build_layer(d, &L);
  if (rc) return rc;
  dims4[0] = L.oB; dims4[1] = L.Ho; dims4[2] = L.Wo; dims4[3] = L.Cs;
  return LDIC_OK;
}

extern "C" int ldic_conv_pack_weights(const LdicConvDesc* d, const float* w, const float* bias, int cin_offset,
                                      void* w_packed, float* bias_packed, void* stream) {
  Layer L;
  int rc = build_layer(d, &L);
  if (rc) return rc;
  if (cin_offset < 0 || (L.cin_map == 0 && L.Kw == d->Cin_pad && cin_offset + d->Cin > d->Cin_pad))
    return fail(LDIC_EINVAL, "pack: cin_offset out of range");
  PackTable T;
  memset(&T, 0, sizeof(T));
  T.ntaps = L.ntaps_total; T.Np = L.Np; T.Cg = L.Cg; T.ngroups = L.ngroups; T.Cin = d->Cin; T.Cout = d->Cout;
  T.Kw = L.Kw; T.cin_offset = cin_offset; T.k = L.k; T.transposed = L.transposed;
  T.cin_map = L.cin_map; T.map_N = L.map_N; T.map_M = L.map_M; T.nbias = L.nbias;
  for (int t = 0; t < kMaxTaps; ++t) {
    T.co_base[t] = L.tap_co_base[t];
    for (int g = 0; g < 4; ++g) { T.ky[t][g] = L.tap_ky[t][g]; T.kx[t][g] = L.tap_kx[t][g]; }
  }
  long long total = (long long)T.ntaps * T.Np * T.Kw;
  int grid = (int)((total + 255) / 256 > num_sms() * 8 ? num_sms() * 8 : (total + 255) / 256);
  if (d->precision) k_pack_weights<float><<<grid, 256, 0, (cudaStream_t)stream>>>(w, bias, T, (float*)w_packed, bias_packed);
  else k_pack_weights<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(w, bias, T, (__nv_bfloat16*)w_packed, bias_packed);
  return check_launch("k_pack_weights");
}

namespace {
int conv_forward_impl(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                      const void* gamma_bf16, const float* beta_tiled, void* y, const LdicConvTail* tail, void* stream,
                      const void* residual = nullptr);
}
extern "C" int ldic_conv_forward(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                                 const void* gamma_bf16, const float* beta_tiled, void* y, void* stream) {
  if (!y) return fail(LDIC_EINVAL, "conv: null output tensor");
  return conv_forward_impl(d, x, w_packed, bias_packed, gamma_bf16, beta_tiled, y, nullptr, stream);
}
extern "C" int ldic_conv_forward_residual(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                                          const void* residual_bf16, void* y, void* stream) {
  if (!y || !residual_bf16) return fail(LDIC_EINVAL, "conv: null output / residual tensor");
  if (d && (d->act == LDIC_ACT_GDN || d->act == LDIC_ACT_IGDN)) return fail(LDIC_EINVAL, "conv: the fused residual excludes the GDN epilogue");
  return conv_forward_impl(d, x, w_packed, bias_packed, nullptr, nullptr, y, nullptr, stream, residual_bf16);
}
extern "C" int ldic_conv_forward_fused_tail(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                                            const void* gamma_bf16, const float* beta_tiled, void* y_or_null,
                                            const LdicConvTail* tail, void* stream) {
  if (!tail || !tail->x_nchw || !tail->w || !tail->sq_err) return fail(LDIC_EINVAL, "conv tail: x_nchw, w and sq_err are required");
  if (!d || d->kind != LDIC_DECONV_GS_5x5_MERGED) return fail(LDIC_EINVAL, "conv tail: only the merged last deconv carries the fused tail");
  if (tail->H != 2 * d->H || tail->W != 2 * d->W) return fail(LDIC_EINVAL, "conv tail: image size must be twice the layer input");
  return conv_forward_impl(d, x, w_packed, bias_packed, gamma_bf16, beta_tiled, y_or_null, tail, stream);
}
namespace {
void print_debug_timing(const Plan& pl, cudaStream_t st) {     // debugging aid only (LDIC_DEBUG_TIMING=1): synchronises
  unsigned long long h[32];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, pl.P.dbg, sizeof(h), cudaMemcpyDeviceToHost);
  fprintf(stderr, "[ldic timing] kind %d Np %d tiles/cta %llu stages %llu | mma total %llu cyc: wait_full %llu wait_buf %llu wait_x2 %llu "
          "(per stage: total %.0f wait_full %.0f) | producer total %llu wait_empty %llu\n", pl.kind, pl.np, h[5], h[4], h[0], h[1], h[2], h[3],
          h[4] ? (double)h[0] / h[4] : 0.0, h[4] ? (double)h[1] / h[4] : 0.0, h[8], h[9]);
  fprintf(stderr, "[ldic timing]   epilogue total %llu cyc over %llu tiles: wait_acc %llu wait_norm %llu wait_slots %llu | busy pass 1 %llu pass 2 + stores %llu\n",
          h[16], h[20], h[17], h[18], h[19], h[21], h[22]);
}

int conv_forward_impl(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                      const void* gamma_bf16, const float* beta_tiled, void* y, const LdicConvTail* tail, void* stream,
                      const void* residual) {
  if (!d) return fail(LDIC_EINVAL, "conv: null desc");
  if (d->B == 0) { Layer L; return build_layer(d, &L); }
  if (!x || !w_packed || (!y && !tail)) return fail(LDIC_EINVAL, "conv: null tensor");
  int rc;
  if ((rc = ensure_timeout_report())) return rc;
  const Tuning& tn = tuning();
  if ((((uintptr_t)x) & 15) || (((uintptr_t)y) & 31)) return fail(LDIC_EINVAL, "conv: x must be 16-byte and y 32-byte aligned");
  PlanKey key;
  memset(&key, 0, sizeof(key));
  key.d = *d; key.w = w_packed; key.bias = bias_packed; key.gamma = gamma_bf16; key.beta = beta_tiled;
  if (tail) { key.has_tail = 1; key.tail_H = tail->H; key.tail_W = tail->W; key.tail_u8 = tail->x_is_u8; key.tail_tanh = tail->tanh_out; }
  key.has_y = y != nullptr;
  key.has_res = residual != nullptr;
  key.device = current_device(); key.epoch = tn.epoch;
  std::shared_ptr<Plan> pl;
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) pl = it->second;
  }
  cudaStream_t st = (cudaStream_t)stream;
  PFN_replaceAddress repl = get_replace_address();
  if (pl && repl) {
    Plan p = *pl;                                   // per-call copy: this call's activation addresses
    if (repl(&p.a, const_cast<void*>(x)) != CUDA_SUCCESS) return fail(LDIC_ECUDA, "conv: cuTensorMapReplaceAddress failed");
    if (p.P.tma_store) {
      if ((((uintptr_t)y) & 127) || repl(&p.y, y) != CUDA_SUCCESS) return fail(LDIC_ECUDA, "conv: cuTensorMapReplaceAddress (output) failed");
    }
    p.P.out = y;
    p.P.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    if (tail) { p.P.tail_x = tail->x_nchw; p.P.tail_w = tail->w; p.P.tail_xo = tail->x_tilde_nchw; p.P.tail_sq = tail->sq_err; }
    if (p.P.dbg) cudaMemsetAsync(p.P.dbg, 0, 32 * sizeof(unsigned long long), st);
    rc = launch_plan(p, st);
    if (rc == LDIC_OK && p.P.dbg) print_debug_timing(p, st);
    return rc;
  }
  pl = std::make_shared<Plan>();
  if ((rc = build_plan(d, x, w_packed, bias_packed, gamma_bf16, beta_tiled, y, tail, pl.get(), residual))) return rc;
  if (tn.debug_timing) {
    unsigned long long* buf = nullptr;
    if (cudaMalloc(&buf, 32 * sizeof(unsigned long long)) == cudaSuccess) pl->P.dbg = buf;   // lives as long as the plan cache
  }
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (g_plans.size() >= kMaxPlans) g_plans.clear();
    g_plans[key] = pl;
  }
  if (pl->P.dbg) cudaMemsetAsync(pl->P.dbg, 0, 32 * sizeof(unsigned long long), st);
  rc = launch_plan(*pl, st);
  if (rc == LDIC_OK && pl->P.dbg) print_debug_timing(*pl, st);
  return rc;
}
}  // namespace

extern "C" int ldic_conv_plan_cache_size(void) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  return (int)g_plans.size();
}
extern "C" void ldic_conv_plan_cache_clear(void) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plans.clear();
}

// ---------------------------------------------------------------------------------
// CUDA-core fp32 direct convolution of the same layer kinds (validation aid).
// x NHWC fp32 [B,H,W,Cin], w = state-dict layout, y NHWC fp32 [B,Ho,Wo,Cout].
// ---------------------------------------------------------------------------------
namespace {
struct RefTable {
  int njobs, Cin, Cout, k, transposed, mode, H, W, Ho, Wo, sy, sx, Hg, Wg, B;
  Job jobs[kMaxJobs];
  Tap taps[kMaxTaps];
  signed char ky[kMaxTaps], kx[kMaxTaps];
};
__global__ void k_conv_ref(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                           float* __restrict__ y, RefTable T) {
  const long long total = (long long)T.njobs * T.B * T.Hg * T.Wg * T.Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % T.Cout);
    long long r = i / T.Cout;
    int gx = (int)(r % T.Wg); r /= T.Wg;
    int gy = (int)(r % T.Hg); r /= T.Hg;
    int n = (int)(r % T.B);
    int j = (int)(r / T.B);
    const Job jb = T.jobs[j];
    float acc = bias ? bias[co] : 0.f;
    for (int tp = 0; tp < jb.ntaps; ++tp) {
      const Tap tap = T.taps[jb.tap_begin + tp];
      int ix, iy;
      if (T.mode == 1) { ix = 2 * (gx + tap.dx) + tap.px; iy = 2 * gy + tap.dy; }
      else { ix = gx + tap.dx; iy = gy + tap.dy; }
      if (ix < 0 || ix >= T.W || iy < 0 || iy >= T.H) continue;
      const float* xp = x + (((long long)n * T.H + iy) * T.W + ix) * T.Cin;
      const int ky = T.ky[jb.tap_begin + tp], kx = T.kx[jb.tap_begin + tp];
      for (int ci = 0; ci < T.Cin; ++ci) {
        long long widx = T.transposed ? ((((long long)ci * T.Cout + co) * T.k + ky) * T.k + kx)
                                      : ((((long long)co * T.Cin + ci) * T.k + ky) * T.k + kx);
        acc = fmaf(xp[ci], __ldg(w + widx), acc);
      }
    }
    int oy = gy * T.sy + jb.oy_off, ox = gx * T.sx + jb.ox_off;
    y[(((long long)n * T.Ho + oy) * T.Wo + ox) * T.Cout + co] = acc;
  }
}
}  // namespace

extern "C" int ldic_conv_forward_f32_reference_kernel(const LdicConvDesc* d, const float* x_nhwc, const float* w,
                                                      const float* bias, float* y_nhwc, void* stream) {
  LdicConvDesc dd = *d;
  if (dd.kind == LDIC_DECONV_GS_5x5_MERGED) dd.kind = LDIC_DECONV_GS_5x5;
  dd.Cout = 8; dd.Cout_pad = 64; dd.Cin_pad = ((dd.Cin + 63) / 64) * 64;   // only the tap tables are used here
  Layer L;
  int rc = build_layer(&dd, &L);
  if (rc) return rc;
  if (d->B == 0) return LDIC_OK;
  RefTable T;
  memset(&T, 0, sizeof(T));
  T.njobs = L.njobs; T.Cin = d->Cin; T.Cout = d->Cout; T.k = L.k; T.transposed = L.transposed; T.mode = L.mode;
  T.H = d->H; T.W = d->W; T.Ho = L.Ho; T.Wo = L.Wo; T.sy = L.sy; T.sx = L.sx; T.Hg = L.Hg; T.Wg = L.Wg; T.B = d->B;
  for (int j = 0; j < kMaxJobs; ++j) T.jobs[j] = L.jobs[j];
  for (int t = 0; t < kMaxTaps; ++t) { T.taps[t] = L.taps[t]; T.ky[t] = (signed char)L.tap_ky[t][0]; T.kx[t] = (signed char)L.tap_kx[t][0]; }
  long long total = (long long)T.njobs * T.B * T.Hg * T.Wg * T.Cout;
  int grid = (int)((total + 255) / 256 > num_sms() * 16 ? num_sms() * 16 : (total + 255) / 256);
  k_conv_ref<<<grid, 256, 0, (cudaStream_t)stream>>>(x_nhwc, w, bias, y_nhwc, T);
  return check_launch("k_conv_ref");
}

// Probe: does a K-major SWIZZLE_128B UMMA operand descriptor work when its start address is shifted by
// whole 128-byte rows (not 1024-byte aligned), and what must base_offset be?  (Needed to reuse one
// halo'd activation region in shared memory for all filter taps of a convolution.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe tools/umma_shift_probe.cu && ./umma_shift_probe
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int kRows = 200;      // region rows (128 B each)
constexpr int kN = 64;

// out[m*kN+n] for one configuration
__global__ void __launch_bounds__(128, 1) probe(float* out, int shift_rows, int sbo_bytes, int base_off, int group_pitch_rows) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* al = raw + (base - smem_u32(raw));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(al);                  // kRows x 64 bf16, TMA-style swizzle
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(al + kRows * 128);   // kN x 64, 1024-aligned (kRows*128 = 25600)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_addr;
  // B = identity over k, so D[m][n] = A[row(m)][n]: columns n < 32 report which region ROW was read,
  // columns n >= 32 report which K element (all values < 256, exact in bf16)
  for (int i = threadIdx.x; i < kRows * 64; i += 128) {
    int r = i / 64, k = i % 64;
    int chunk = k / 8, e = k % 8;
    int phys = r * 64 + ((chunk ^ (r & 7)) * 8) + e;
    // bf16 has 8 significant bits: encode r (0..199) in the integer part only for k < 32, and k in the rest
    A[phys] = __float2bfloat16(k < 32 ? (float)r : (float)(200 + (k - 32)));
  }
  for (int i = threadIdx.x; i < kN * 64; i += 128) {
    int n = i / 64, k = i % 64;
    int chunk = k / 8, e = k % 8;
    int phys = n * 64 + ((chunk ^ (n & 7)) * 8) + e;
    B[phys] = __float2bfloat16(n == k ? 1.0f : 0.0f);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_addr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_addr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_start = base + shift_rows * 128;
    (void)group_pitch_rows;
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_desc(a_start + k * 32, sbo_bytes, base_off);
      uint64_t bd = make_desc(base + kRows * 128 + k * 32, 1024, 0);
      uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                   "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("timeout\n"); break; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c = 0; c < kN; c += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tm + ((uint32_t)(warp * 32) << 16) + c)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int k = 0; k < 32; ++k) out[row * kN + c + k] = __uint_as_float(r[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm));
}

int main() {
  float* d; cudaMalloc(&d, 128 * kN * sizeof(float));
  float* h = (float*)malloc(128 * kN * sizeof(float));
  size_t smem = kRows * 128 + kN * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // (shift, sbo, base_off, pitch): contiguous rows with shifts; then 8-row groups at a 10-row pitch (SBO 1280)
  int cfg[][4] = {{0, 1024, 0, 8}, {1, 1024, 0, 8}, {1, 1024, 1, 8}, {3, 1024, 0, 8}, {3, 1024, 3, 8}, {8, 1024, 0, 8},
                  {0, 1280, 0, 10}, {1, 1280, 0, 10}, {1, 1280, 1, 10}, {11, 1280, 0, 10}, {11, 1280, 3, 10}, {2, 2560, 0, 20}, {2, 2560, 2, 20}};
  for (auto& c : cfg) {
    cudaMemset(d, 0, 128 * kN * sizeof(float));
    probe<<<1, 128, smem>>>(d, c[0], c[1], c[2], c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg shift %d sbo %d bo %d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 128 * kN * sizeof(float), cudaMemcpyDeviceToHost);
    // expected: output row m reads region row  shift + (m/8)*pitch + m%8 ; D[m][n] = n<32 ? that row : 1000+n
    int bad_row = 0, bad_k = 0;
    for (int m = 0; m < 128; ++m) {
      int er = c[0] + (m / 8) * c[3] + (m % 8);
      for (int n = 0; n < kN; ++n) {
        float exp = n < 32 ? (float)er : (float)(200 + (n - 32));
        if (h[m * kN + n] != exp) { if (n < 32) bad_row++; else bad_k++; }
      }
    }
    printf("shift %2d sbo %4d base_off %d: row mismatches %4d, k mismatches %4d | m=0: %.0f %.0f  m=1: %.0f  m=7: %.0f m=8: %.0f m=9: %.0f k: %.0f %.0f %.0f\n", c[0], c[1], c[2], bad_row,
           bad_k, h[0], h[1], h[kN], h[7 * kN], h[8 * kN], h[9 * kN], h[32], h[40], h[63]);
  }
  return 0;
}

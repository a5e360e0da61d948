// Probe: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16, operands in shared memory, SWIZZLE_128B K-major)
// against N and against the alignment of the A descriptor start.  One CTA per SM, one thread issues `reps` groups of
// 4 MMAs (one 64-wide K block) on resident operands, then commits; cycles are measured from first issue to completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe tools/umma_rate_probe.cu && ./umma_rate_probe
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out, int N, int shift_rows, int reps, int nbufs) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_addr;
  for (int i = threadIdx.x; i < 32768; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_addr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_addr;
  if (threadIdx.x < 32) {                 // whole warp converged, one elected lane issues (as the conv kernels do)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_lo = (((base + 98304u) & 0x3FFFFu) >> 4) | (1u << 16);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t a_lo = (((base + (uint32_t)(g % nbufs) * 24576u + shift_rows * 128) & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                "setp.ne.b32 p, %6, 0;\n\t"
                "mov.b64 da, {%1, %2};\n\t"
                "mov.b64 db, {%3, %4};\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tm),
                "r"(a_lo + 2 * k), "r"(hi), "r"(b_lo + 2 * k), "r"(hi), "r"(idesc), "r"((uint32_t)((r | g | k) != 0))
                : "memory");
          }
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  long long h[2];
  size_t smem = 98304 + 32768 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 2000;
  for (int grid : {1, 148}) {
    for (int N : {32, 64, 96, 128, 192, 256}) {
      for (int shift : {0, 1, 8}) {
        for (int nb : {1, 4}) {
          probe<<<grid, 128, smem>>>(d, N, shift, reps, nb);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid %3d N %3d shift %d abufs %d: %.1f cycles per MMA (issue loop %.1f) floor %d\n", grid, N, shift, nb,
                 (double)h[1] / (4.0 * reps), (double)h[0] / (4.0 * reps), N / 2);
        }
      }
    }
  }
  return 0;
}

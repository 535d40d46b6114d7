// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16 in / fp32 out) as a function of N, M, the
// operand source of A (shared memory descriptor vs tensor memory) and the shared-memory swizzle span.
// One thread per CTA issues `reps` back-to-back MMAs on resident operands and waits for one commit; no TMA, no epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run: ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg { int m, n, ts, sw, reps, ctas_per_sm, mn, nacc; };

template <int NACC>
__global__ void __launch_bounds__(128) mma_rate_kernel(Cfg c, uint32_t idesc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint64_t layout = (c.sw == 128) ? 2 : (c.sw == 64) ? 4 : 6;
    auto desc = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3ffff) >> 4);
      d |= (uint64_t)(c.mn ? (c.sw >> 4) : 1) << 16;                      // MN-major: channel groups one pixel apart (wgrad_tc3)
      d |= (uint64_t)((c.mn ? 10 * c.sw : 8 * c.sw) >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= layout << 61;
      return d;
    };
    // warp-uniform values: read through lane 0 so that the compiler can keep them in uniform registers
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint64_t ad = desc(__shfl_sync(0xffffffffu, smem_u32(smem), 0));
    const uint64_t bd = desc(__shfl_sync(0xffffffffu, smem_u32(smem + 16 * 1024), 0));
    const uint32_t a_tmem = tmu + 256;             // A operand columns when sourced from tensor memory
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    long long t0 = clock64();
    if (elected) {
      for (int r = 0; r < c.reps; r += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t koff = c.mn ? (uint64_t)((k * 2 * 10 * c.sw) >> 4) : (uint64_t)((((c.sw == 32) ? 0 : k) * 32) >> 4);
          if (c.ts) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmu), "r"(a_tmem), "l"(bd + koff), "r"(idesc), "r"((r | k) ? 1u : 0u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmu + (uint32_t)((k % NACC) * 128)), "l"(ad + koff), "l"(bd + koff), "r"(idesc), "r"((r | k) ? 1u : 0u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    long long t_issue = clock64();
    uint32_t ok = 0;
    for (long long spin = 0; !ok && spin < (1ll << 26); ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = ok ? (t1 - t0) : -1; out[512 + blockIdx.x] = t_issue - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

static uint32_t idesc_bf16(int m, int n, int mn = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)mn << 15) | ((uint32_t)mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

int main() {
  long long* out;
  cudaMalloc(&out, 1024 * sizeof(long long));
  const int smem_bytes = (16 + 32 + 1) * 1024;
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(mma_rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  printf("mode sw M N ctas/SM cycles/MMA floor(128*N/256) bytes/MMA smemB/clk\n");
  const int ns[] = {16, 32, 64, 128, 256};
  for (int per_sm = 1; per_sm <= 1; ++per_sm)
    for (int ts = 0; ts <= 1; ++ts)
      for (int sw : {128, 32})
        for (int m : {128, 64})
          for (int n : ns) {
            if (sw == 32 && n > 64) continue;
            Cfg c{m, n, ts, sw, 4096, per_sm, 0, 1};
            mma_rate_kernel<1><<<148 * per_sm, 128, smem_bytes>>>(c, idesc_bf16(m, n), out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s (ts=%d sw=%d m=%d n=%d)\n", cudaGetErrorString(e), ts, sw, m, n); return 1; }
            long long h[1024];
            cudaMemcpy(h, out, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0, iss = 0;
            for (int i = 0; i < 148 * per_sm; ++i) { avg += (double)h[i]; iss += (double)h[512 + i]; }
            avg /= (148.0 * per_sm * c.reps);
            iss /= (148.0 * per_sm * c.reps);
            const double bytes = (ts ? 0 : m * 32) + n * 32;
            printf("%s %3d %3d %3d %d %8.1f %6.1f %6.0f %6.1f issue %6.1f\n", ts ? "TS" : "SS", sw, m, n, per_sm, avg, 128.0 * n / 256, bytes, bytes / avg, iss);
          }
  for (int nacc : {2, 4})
    for (int n : {16, 32, 64, 128}) {
      Cfg c{128, n, 0, 128, 4096, 1, 0, nacc};
      if (nacc == 2) mma_rate_kernel<2><<<148, 128, smem_bytes>>>(c, idesc_bf16(128, n, 0), out);
      else mma_rate_kernel<4><<<148, 128, smem_bytes>>>(c, idesc_bf16(128, n, 0), out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s (nacc)\n", cudaGetErrorString(e)); return 1; }
      long long h[1024];
      cudaMemcpy(h, out, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      printf("SS accumulators %d M 128 N %3d cycles/MMA %8.1f\n", nacc, n, avg / (148.0 * c.reps));
    }
  for (int sw : {32, 64, 128})
    for (int n : {16, 32, 64, 128}) {
      Cfg c{128, n, 0, sw, 4096, 1, 1, 1};
      mma_rate_kernel<1><<<148, 128, smem_bytes>>>(c, idesc_bf16(128, n, 1), out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s (mn sw=%d n=%d)\n", cudaGetErrorString(e), sw, n); return 1; }
      long long h[1024];
      cudaMemcpy(h, out, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < 148; ++i) avg += (double)h[i];
      printf("SS-MNmajor sw %3d M 128 N %3d cycles/MMA %8.1f\n", sw, n, avg / (148.0 * c.reps));
    }
  return 0;
}

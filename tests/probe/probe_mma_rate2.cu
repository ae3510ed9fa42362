// Measures tcgen05.mma (kind::f16, bf16 operands from shared memory, M=128) issue
// cost as a function of N: one CTA per SM issues a long dependent-free stream of
// MMAs on a fixed A/B image and a ring of accumulators; reports cycles per MMA and
// the implied TFLOP/s.  Decides whether the conv kernel should merge the three dz
// slots into one wide-N MMA (is the SS-mode A read from smem the bound at small N?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mma_rate probe_mma_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n",cudaGetErrorString(e),__LINE__);exit(2);} } while(0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) rate(int N, int iters, int distinct_a, int nchain, long long* cycles, int* status, int commit_every = 0, int alt_b = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: 4 distinct images of [2 chunks][186 rows][16B]; B: [2 chunks][256 rows][16B]
  for (int i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1 << 20)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 32768;
    // descriptors and accumulator columns precomputed: the loop body is just the MMA
    uint64_t ad[4], bd = make_desc(b_base, 256 * 16, 128);
    for (int j = 0; j < 4; ++j)
      ad[j] = make_desc(a_base + (distinct_a ? j * 6144 + j * 16 : 0), 186 * 16, 10 * 16);
    // nchain independent accumulators (dependent-accumulate chains) used round-robin
    uint32_t cc[8];
    for (int j = 0; j < 8; ++j) cc[j] = tmem_base + (uint32_t)((j % nchain) * (nchain > 2 ? 128 : 256));
    uint64_t bd2 = make_desc(b_base + (alt_b ? 256 * 16 * 2 : 0), 256 * 16, 128);
    t0 = clock64();
    int since = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(cc[j]), "l"(ad[j & 3]), "l"((j & 1) ? bd2 : bd), "r"(idesc), "r"(1u) : "memory");
      }
      since += 8;
      if (commit_every && since >= commit_every) {
        since = 0;
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  bool ok = false;
  for (long long it = 0; it < 400000000LL && !ok; ++it) {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(p) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    ok = p != 0;
  }
  if (tid == 0) { t1 = clock64(); cycles[blockIdx.x] = t1 - t0; if (!ok) *status = 1; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  const int sms = p.multiProcessorCount;
  long long* dc; int* ds; CK(cudaMalloc(&dc, sms * 8)); CK(cudaMalloc(&ds, 4));
  CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int iters = 20000;
  for (int variant = 0; variant < 6; ++variant)
  for (int nchain : {1})
    for (int N : {96}) {
      const int commit_every = variant == 0 ? 0 : variant == 1 ? 56 : variant == 2 ? 24 : variant == 3 ? 8 : 0;
      const int alt_b = variant >= 4 ? 1 : 0;
      const int ce2 = variant == 5 ? 56 : commit_every;
      const int distinct = 1;
      CK(cudaMemset(ds, 0, 4));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      rate<<<sms, 128, 64 * 1024>>>(N, 2000, distinct, nchain, dc, ds, ce2, alt_b);  // warm-up
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      rate<<<sms, 128, 64 * 1024>>>(N, iters, distinct, nchain, dc, ds, ce2, alt_b);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long hc[256]; int st; CK(cudaMemcpy(hc, dc, sms * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < sms; ++i) avg += hc[i]; avg /= sms;
      const double flops = 2.0 * 128 * N * 16 * (double)iters * sms;
      printf("commit_every=%d alt_b=%d chains=%d distinctA=%d N=%3d  cycles/MMA=%7.2f  time=%.3f ms  %.1f TFLOP/s (bf16)  timeout=%d\n", ce2, alt_b, nchain, distinct, N, avg / iters, ms, flops / (ms * 1e-3) / 1e12, st);
    }
  return 0;
}

// Bring-up probe for tcgen05.mma.cta_group::2 with the conv kernel's operand conventions (not part
// of the product; run on the GPU box via `gpurun`, prints PASS/FAIL and rate lines).
//
// A CTA pair (cluster of 2) executes ONE MMA of M = 256: CTA r provides accumulator rows
// [128 r, 128 r + 128) from ITS shared memory at the descriptor's address, and ITS half of the
// B image (N/2 rows at the descriptor's address); the accumulator tile [128 x N] of each CTA
// lives in its own TMEM at the same column address.  Pinned here against a CPU reference:
//   * which CTA's B half lands in which accumulator columns,
//   * K-major no-swizzle descriptors with a non-dense SBO (halo-brick pitch) in pair mode,
//   * tcgen05.commit ... multicast::cluster arriving on the same barrier offset in both CTAs,
//   * the issue cost per MMA as a function of N (is the smem operand read really halved?).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_cta2 probe_cta2.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) {                                                                \
      printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__);                  \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
               ::: "memory");
}

constexpr int A_ROWS = 186, A_PITCH = 10;   // the conv kernel's stride-1 brick: 18 x 10 + slack
constexpr int K = 32;                       // two K steps of 16
constexpr int A_BYTES = (K / 8) * A_ROWS * 16;
constexpr int B_MAXROWS = 128;
constexpr int B_BYTES = (K / 8) * B_MAXROWS * 16;

// mode 0: numerics (one pass, D written to global); mode 1: rate (iters MMAs, cycles out)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    pair_kernel(int N, int a_off_rows, int mode, int iters, const uint16_t* __restrict__ a_img,
                const uint16_t* __restrict__ b_img, float* __restrict__ d_out,
                long long* __restrict__ cycles, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 32768;
  const int nh = N / 2;
  // A image of this CTA: [K/8 chunks][A_ROWS][8 bf16]; B half: [K/8][nh rows][8 bf16]
  for (int i = tid; i < A_BYTES / 2; i += 128)
    reinterpret_cast<uint16_t*>(a_s)[i] = a_img[(size_t)rank * (A_BYTES / 2) + i];
  for (int i = tid; i < (K / 8) * nh * 8; i += 128)
    reinterpret_cast<uint16_t*>(b_s)[i] = b_img[(size_t)rank * ((K / 8) * nh * 8) + i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // both CTAs' operands, barriers and TMEM are ready
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  long long t0 = 0;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(256 >> 4) << 24);
    const uint32_t a0 = smem_u32(a_s) + a_off_rows * 16, b0 = smem_u32(b_s);
    t0 = clock64();
    const int reps = mode == 0 ? 1 : iters;
#pragma unroll 1
    for (int it = 0; it < reps; ++it) {
#pragma unroll
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = make_desc(a0 + 2 * ks * A_ROWS * 16, A_ROWS * 16, A_PITCH * 16);
        const uint64_t bd = make_desc(b0 + 2 * ks * nh * 16, nh * 16, 128);
        const uint32_t acc = (mode == 0 && ks == 0) ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
            "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(&bar)),
        "h"((uint16_t)3)
        : "memory");
  }
  bool ok = false;
  for (long long it = 0; it < 200000000LL && !ok; ++it) {
    uint32_t p;
    asm volatile(
        "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(p)
        : "r"(smem_u32(&bar)), "r"(0)
        : "memory");
    ok = p != 0;
  }
  if (rank == 0 && tid == 0 && mode == 1) cycles[pair] = clock64() - t0;
  if (!ok && tid == 0) *status = 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode == 0 && ok) {
    // this CTA's accumulator: rows 128*rank + (32*warp + lane), N columns
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
          "%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
            "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = (int)rank * 128 + warp * 32 + lane;
      for (int j = 0; j < 16; ++j)
        d_out[((size_t)pair * 256 + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(512));
}

static uint16_t bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);  // exact for the small integers used here
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int smem_bytes = 64 * 1024;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int* d_status;
  long long* d_cycles;
  CK(cudaMalloc(&d_status, 4));
  CK(cudaMalloc(&d_cycles, sms * 8));

  // ---------------- numerics ----------------
  int fails = 0;
  for (int N : {32, 96, 192, 256}) {
    for (int a_off : {0, 1, 11, 22}) {  // tap shifts: +1 (dx), +pitch+1, +2*pitch+2
      const int nh = N / 2;
      // logical A[256][K], B[N][K]: small integers
      std::vector<float> A(256 * K), B((size_t)N * K);
      for (int r = 0; r < 256; ++r)
        for (int k = 0; k < K; ++k) A[r * K + k] = (float)((r * 3 + k * 7 + N) % 13 - 6);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) B[(size_t)n * K + k] = (float)((n * 5 + k * 11) % 9 - 4);
      // images: A of CTA r: row m (0..127) sits at brick row a_off + (m/8)*A_PITCH + m%8
      std::vector<uint16_t> a_img(2 * (A_BYTES / 2), bf16_bits(100.f)), b_img(2 * (K / 8) * nh * 8);
      for (int r = 0; r < 2; ++r)
        for (int m = 0; m < 128; ++m)
          for (int k = 0; k < K; ++k) {
            const int row = a_off + (m / 8) * A_PITCH + m % 8;
            a_img[(size_t)r * (A_BYTES / 2) + ((k / 8) * A_ROWS + row) * 8 + k % 8] =
                bf16_bits(A[(r * 128 + m) * K + k]);
          }
      // hypothesis: CTA r holds B rows [r*nh, (r+1)*nh)
      for (int r = 0; r < 2; ++r)
        for (int n = 0; n < nh; ++n)
          for (int k = 0; k < K; ++k)
            b_img[(size_t)r * ((K / 8) * nh * 8) + ((k / 8) * nh + n) * 8 + k % 8] =
                bf16_bits(B[(size_t)(r * nh + n) * K + k]);
      uint16_t *da, *db;
      float* dd;
      CK(cudaMalloc(&da, a_img.size() * 2));
      CK(cudaMalloc(&db, b_img.size() * 2));
      CK(cudaMalloc(&dd, (size_t)256 * N * 4));
      CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
      CK(cudaMemset(dd, 0xff, (size_t)256 * N * 4));
      CK(cudaMemset(d_status, 0, 4));
      pair_kernel<<<2, 128, smem_bytes>>>(N, a_off, 0, 1, da, db, dd, d_cycles, d_status);
      CK(cudaDeviceSynchronize());
      std::vector<float> D((size_t)256 * N);
      int st;
      CK(cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
      long bad = 0, bad_swapped = 0;
      for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0.f, ref_sw = 0.f;
          const int nsw = (n + nh) % N;
          for (int k = 0; k < K; ++k) {
            ref += A[r * K + k] * B[(size_t)n * K + k];
            ref_sw += A[r * K + k] * B[(size_t)nsw * K + k];
          }
          if (D[(size_t)r * N + n] != ref) ++bad;
          if (D[(size_t)r * N + n] != ref_sw) ++bad_swapped;
        }
      const bool pass = st == 0 && bad == 0;
      if (!pass) ++fails;
      printf("%s numerics N=%3d a_off=%2d timeout=%d mismatches=%ld (halves swapped: %ld)\n",
             pass ? "PASS" : "FAIL", N, a_off, st, bad, bad_swapped);
      cudaFree(da);
      cudaFree(db);
      cudaFree(dd);
    }
  }

  // ---------------- rate ----------------
  {
    const int N0 = 256, nh0 = N0 / 2;
    std::vector<uint16_t> a_img(2 * (A_BYTES / 2), bf16_bits(1.f)),
        b_img(2 * (K / 8) * nh0 * 8, bf16_bits(1.f));
    uint16_t *da, *db;
    float* dd;
    CK(cudaMalloc(&da, a_img.size() * 2));
    CK(cudaMalloc(&db, b_img.size() * 2));
    CK(cudaMalloc(&dd, 1024));
    CK(cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice));
    const int npairs = sms / 2;
    const int iters = 10000;  // x K/16 = 2 MMAs each
    for (int N : {32, 48, 64, 96, 128, 192, 256}) {
      if (N % 16) continue;
      CK(cudaMemset(d_status, 0, 4));
      pair_kernel<<<2 * npairs, 128, smem_bytes>>>(N, 0, 1, 500, da, db, dd, d_cycles, d_status);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      pair_kernel<<<2 * npairs, 128, smem_bytes>>>(N, 0, 1, iters, da, db, dd, d_cycles, d_status);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> hc(npairs);
      int st;
      CK(cudaMemcpy(hc.data(), d_cycles, npairs * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
      double avg = 0;
      for (long long c : hc) avg += c;
      avg /= npairs;
      const double nmma = 2.0 * iters;
      const double flops = 2.0 * 256 * N * 16 * nmma * npairs;
      printf("rate pair M=256 N=%3d  cycles/MMA=%7.2f  %.1f TFLOP/s (bf16, %d pairs)  timeout=%d\n", N,
             avg / nmma, flops / (ms * 1e-3) / 1e12, npairs, st);
    }
  }
  printf(fails ? "probe_cta2: %d FAILED\n" : "probe_cta2: all numerics cases passed\n", fails);
  return fails ? 1 : 0;
}

// Stand-alone bring-up probe for the tcgen05 conventions the conv kernel relies on
// (not part of the product; run on the GPU box via `gpurun`, prints PASS/FAIL lines).
//
// It pins, against a CPU reference, the exact semantics of
//   * the shared-memory matrix descriptor for K-major, no-swizzle bf16 operands:
//     8-row x 16-byte core matrices, LBO = stride between the two K core matrices of
//     one MMA, SBO = stride between 8-row groups (cute/atom/mma_traits_sm100.hpp,
//     "LayoutType::INTERLEAVE : ((8,n),2):((1,SBO),LBO)");
//   * arbitrary (non-dense) SBO and 16-byte-granular start offsets, which is what
//     lets one halo brick in shared memory serve all 9 in-plane taps of a 3x3x3 conv;
//   * N sub-slices of a wider B image and TMEM column offsets (dz-merged accumulators);
//   * the accumulate predicate and tcgen05.ld 32x32b lane/column mapping.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma probe_umma.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__,     \
             __LINE__);                                                        \
      exit(2);                                                                 \
    }                                                                          \
  } while (0)

struct Case {
  int K;        // multiple of 16
  int N;        // MMA N (multiple of 16)
  int NT;       // rows in the B image (>= n0 + N)
  int n0;       // first B row used
  int G;        // A 8-row-group stride, in rows
  int off;      // A start row offset
  int R;        // rows per chunk array in the A image
  int col0;     // TMEM column offset of the accumulator
  int swap;     // 1: swap LBO/SBO fields (alternative convention)
  int twice;    // 1: issue the whole K loop twice (second time accumulating)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo,
                                              uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16* __restrict__ a_img, const __nv_bfloat16* __restrict__ b_img,
      float* __restrict__ d_out, int* __restrict__ status, Case c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = c.K / 8;
  const uint32_t a_bytes = (uint32_t)nchunk * c.R * 16;
  const uint32_t b_bytes = (uint32_t)nchunk * c.NT * 16;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((a_bytes + 1023) & ~1023u);

  for (uint32_t i = tid; i < a_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (uint32_t i = tid; i < b_bytes / 16; i += 128)
    reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_base_s)),
                 "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    // instruction descriptor: D=f32, A=B=bf16, K-major both, M=128, N=c.N
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_lbo = (uint32_t)c.R * 16, a_sbo = (uint32_t)c.G * 16;
    const uint32_t b_lbo = (uint32_t)c.NT * 16, b_sbo = 128;
    for (int rep = 0; rep <= c.twice; ++rep) {
      for (int ks = 0; ks < c.K / 16; ++ks) {
        const uint32_t a_addr = smem_u32(a_s) + 2 * ks * a_lbo + c.off * 16;
        const uint32_t b_addr = smem_u32(b_s) + 2 * ks * b_lbo + c.n0 * 16;
        const uint64_t ad = c.swap ? make_desc(a_addr, a_sbo, a_lbo)
                                   : make_desc(a_addr, a_lbo, a_sbo);
        const uint64_t bd = c.swap ? make_desc(b_addr, b_sbo, b_lbo)
                                   : make_desc(b_addr, b_lbo, b_sbo);
        const uint32_t acc = (ks > 0 || rep > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(
                tmem_base + c.col0),
            "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(&bar))
        : "memory");
  }
  // bounded wait on phase 0
  bool ok = false;
  for (int it = 0; it < 4000000 && !ok; ++it) {
    uint32_t p;
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(p)
        : "r"(smem_u32(&bar)), "r"(0)
        : "memory");
    ok = p != 0;
  }
  if (!ok && tid == 0) *status = 1;
  ok = __syncthreads_and(ok);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    for (int cb = 0; cb < c.N; cb += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c.col0 + cb;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
          "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),
            "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
            "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
            "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int m = warp * 32 + lane;
      for (int j = 0; j < 32; ++j)
        if (cb + j < c.N) d_out[m * c.N + cb + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(256));
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run_case(const char* name, Case c) {
  const int nchunk = c.K / 8;
  std::vector<__nv_bfloat16> a((size_t)nchunk * c.R * 8), b((size_t)nchunk * c.NT * 8);
  std::vector<float> af(a.size()), bfv(b.size());
  uint32_t s = 12345u + c.K * 7 + c.N * 13 + c.G * 31 + c.off;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
  };
  for (size_t i = 0; i < a.size(); ++i) {
    af[i] = bf(rnd());
    a[i] = __float2bfloat16(af[i]);
  }
  for (size_t i = 0; i < b.size(); ++i) {
    bfv[i] = bf(rnd());
    b[i] = __float2bfloat16(bfv[i]);
  }
  // CPU reference
  std::vector<float> ref((size_t)128 * c.N);
  for (int m = 0; m < 128; ++m) {
    const int row = (m / 8) * c.G + (m % 8) + c.off;
    for (int n = 0; n < c.N; ++n) {
      double acc = 0;
      for (int k = 0; k < c.K; ++k)
        acc += (double)af[((size_t)(k / 8) * c.R + row) * 8 + k % 8] *
               (double)bfv[((size_t)(k / 8) * c.NT + c.n0 + n) * 8 + k % 8];
      ref[(size_t)m * c.N + n] = (float)(acc * (c.twice ? 2.0 : 1.0));
    }
  }
  __nv_bfloat16 *da, *db;
  float* dd;
  int* ds;
  CK(cudaMalloc(&da, a.size() * 2));
  CK(cudaMalloc(&db, b.size() * 2));
  CK(cudaMalloc(&dd, ref.size() * 4));
  CK(cudaMalloc(&ds, 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xFF, ref.size() * 4));
  CK(cudaMemset(ds, 0, 4));
  const size_t smem = (((size_t)nchunk * c.R * 16 + 1023) & ~1023u) +
                      (size_t)nchunk * c.NT * 16 + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(da, db, dd, ds, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s swap=%d  CUDA ERROR %s\n", name, c.swap, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> out(ref.size());
  int st = 0;
  CK(cudaMemcpy(out.data(), dd, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    double d = std::fabs((double)out[i] - (double)ref[i]);
    if (!(d == d)) d = 1e30;
    if (d > maxerr) maxerr = d;
    if (std::fabs(ref[i]) > maxref) maxref = std::fabs(ref[i]);
  }
  const bool pass = st == 0 && maxerr <= 1e-4 * maxref + 1e-5;
  printf("%-28s swap=%d  %s  maxerr=%.3e maxref=%.3e timeout=%d\n", name, c.swap,
         pass ? "PASS" : "FAIL", maxerr, maxref, st);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dd);
  cudaFree(ds);
  return pass;
}

int main() {
  int ok_std = 0, ok_swap = 0, total = 0;
  struct Named {
    const char* n;
    Case c;
  } cases[] = {
      //                     K   N   NT  n0  G  off  R  col0 swap twice
      {"dense_K16_N32", {16, 32, 32, 0, 8, 0, 128, 0, 0, 0}},
      {"dense_K32_N32", {32, 32, 32, 0, 8, 0, 128, 0, 0, 0}},
      {"dense_K64_N64", {64, 64, 64, 0, 8, 0, 128, 0, 0, 0}},
      {"G10_off0_K32_N32", {32, 32, 32, 0, 10, 0, 186, 0, 0, 0}},
      {"G10_off11_K32_N32", {32, 32, 32, 0, 10, 11, 186, 0, 0, 0}},
      {"G10_off22_K32_N96", {32, 96, 96, 0, 10, 22, 186, 0, 0, 0}},
      {"G20_off1_K32_N96_col32", {32, 96, 96, 0, 20, 1, 330, 32, 0, 0}},
      {"slice_n0_32_N64_col64", {32, 64, 96, 32, 10, 3, 186, 64, 0, 0}},
      {"K64_N192", {64, 192, 192, 0, 10, 5, 186, 0, 0, 0}},
      {"twice_acc", {32, 32, 32, 0, 10, 7, 186, 96, 0, 1}},
      {"N16", {32, 16, 16, 0, 10, 2, 186, 16, 0, 0}},
      {"N48_slice16", {32, 48, 64, 16, 10, 2, 186, 48, 0, 0}},
  };
  for (auto& nc : cases) {
    ++total;
    Case c = nc.c;
    c.swap = 0;
    ok_std += run_case(nc.n, c);
  }
  printf("SUMMARY std=%d/%d swap=%d/%d\n", ok_std, total, ok_swap, total);
  return 0;
}

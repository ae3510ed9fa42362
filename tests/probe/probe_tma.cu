// Stand-alone bring-up probe for the TMA (cp.async.bulk.tensor) conventions the conv kernel's
// TMA loader relies on (not part of the product; run on the GPU box, prints PASS/FAIL lines).
//
// Tensor: the "pre-split" activation layout [S = planes*2*chunks][H][W][8 x bf16] seen as a
// rank-4 uint16 tensor {8, W, H, S}.  Pins, against a CPU gather:
//   * a plain box {8, BX, BY, NS} at a start coordinate with NEGATIVE x / y (halo): out-of-range
//     elements arrive as zeros, the full box byte count is signalled on the mbarrier, the box
//     lands dense in shared memory as [s][y][x][16 B];
//   * elementStrides {1, 2, 2, 1}: the box is given in un-strided tensor coordinates
//     (boxDim = 2*n - 1 covers n elements), ceil(boxDim / stride) elements are written, and the
//     byte count signalled is that of the elements written;
//   * destination addresses that are 128-byte aligned but not more.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tma probe_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e = (x);                                                                  \
    if (e != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);      \
      exit(2);                                                                            \
    }                                                                                     \
  } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

struct Req {
  int x, y, s;       // start coordinates (dims 1, 2, 3)
  int dst_off;       // byte offset in shared memory
  int bytes;         // expected transaction bytes
};

__global__ void probe_kernel(const __grid_constant__ CUtensorMap map, Req r, uint8_t* out,
                             int out_bytes, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) smem[i] = 0xEE;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)),
                 "r"(r.bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem + r.dst_off)),
        "l"(&map), "r"(0), "r"(r.x), "r"(r.y), "r"(r.s), "r"(smem_u32(&bar))
        : "memory");
    int ok = 0;
    for (int it = 0; it < (1 << 22) && !ok; ++it) {
      asm volatile(
          "{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}\n"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0)
          : "memory");
    }
    *status = ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) out[i] = smem[i];
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) {
    printf("FAIL: no cuTensorMapEncodeTiled entry point\n");
    return 1;
  }
  EncodeFn encode = (EncodeFn)fn;
  const int W = 37, H = 21, S = 6;
  std::vector<uint16_t> h((size_t)S * H * W * 8);
  for (int s = 0; s < S; ++s)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int e = 0; e < 8; ++e)
          h[(((size_t)s * H + y) * W + x) * 8 + e] = (uint16_t)(1 + ((s * 31 + y) * 41 + x) * 8 + e);
  uint16_t* d = nullptr;
  CK(cudaMalloc(&d, h.size() * 2));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  uint8_t* d_out = nullptr;
  int* d_status = nullptr;
  const int OUT = 64 * 1024;
  CK(cudaMalloc(&d_out, OUT));
  CK(cudaMalloc(&d_status, 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OUT));
  int fails = 0;
  struct Case { const char* name; int bx, by, ns, stride, x, y, s, dst; };
  const Case cases[] = {
      {"plain box, interior", 10, 18, 1, 1, 3, 2, 1, 0},
      {"plain box, negative start (halo)", 10, 18, 1, 1, -1, -1, 0, 128},
      {"plain box, past the far edges", 10, 18, 1, 1, 30, 10, 5, 384},
      {"plain box, 4 slabs", 10, 18, 4, 1, -1, 7, 2, 0},
      {"stride 2, interior", 9, 17, 1, 2, 5, 1, 0, 0},
      {"stride 2, negative odd start, 2 slabs", 9, 17, 2, 2, -1, -1, 3, 2560},
      {"stride 2, even start past the edge", 9, 17, 2, 2, 26, 0, 4, 128},
  };
  for (const Case& c : cases) {
    CUtensorMap map;
    const cuuint64_t dims[4] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S};
    const cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    // elements loaded along a strided dim: n; box in un-strided coordinates: (n-1)*stride + 1
    const cuuint32_t box[4] = {8, (cuuint32_t)((c.bx - 1) * c.stride + 1),
                               (cuuint32_t)((c.by - 1) * c.stride + 1), (cuuint32_t)c.ns};
    const cuuint32_t estr[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("FAIL: %s: encode returned %d\n", c.name, (int)r);
      ++fails;
      continue;
    }
    Req rq{c.x, c.y, c.s, c.dst, c.bx * c.by * c.ns * 16};
    CK(cudaMemset(d_status, 0, 4));
    probe_kernel<<<1, 128, OUT>>>(map, rq, d_out, OUT, d_status);
    CK(cudaDeviceSynchronize());
    int status = 0;
    std::vector<uint8_t> o(OUT);
    CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o.data(), d_out, OUT, cudaMemcpyDeviceToHost));
    long bad = 0;
    if (!status) {
      printf("FAIL: %s: mbarrier never completed with %d expected bytes\n", c.name, rq.bytes);
      ++fails;
      continue;
    }
    const uint16_t* o16 = reinterpret_cast<const uint16_t*>(o.data() + c.dst);
    for (int s = 0; s < c.ns; ++s)
      for (int j = 0; j < c.by; ++j)
        for (int i = 0; i < c.bx; ++i)
          for (int e = 0; e < 8; ++e) {
            const int gs = c.s + s, gy = c.y + j * c.stride, gx = c.x + i * c.stride;
            uint16_t want = 0;
            if (gs >= 0 && gs < S && gy >= 0 && gy < H && gx >= 0 && gx < W)
              want = h[(((size_t)gs * H + gy) * W + gx) * 8 + e];
            const uint16_t got = o16[(((size_t)s * c.by + j) * c.bx + i) * 8 + e];
            if (got != want) ++bad;
          }
    // nothing outside the destination range may have been touched
    for (int i = 0; i < OUT; ++i)
      if ((i < c.dst || i >= c.dst + rq.bytes) && o[i] != 0xEE) ++bad;
    printf("%s: %s (%ld mismatches)\n", bad ? "FAIL" : "PASS", c.name, bad);
    if (bad) ++fails;
  }
  printf("%s\n", fails ? "PROBE FAILED" : "PROBE OK");
  return fails ? 1 : 0;
}

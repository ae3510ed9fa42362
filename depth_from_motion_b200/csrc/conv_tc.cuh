// tcgen05 tensor-core implicit-GEMM 3x3x3 convolution (stride 1, pad 1) for sm_100a.
//
// Formulation (see DESIGN.md "conv_tc"):
//   * activations are channels-last fp32 in HBM; a persistent CTA owns an 8x16 (x,y)
//     output tile and marches along z.  For every input plane the loader warps
//     gather a 10x18 halo brick, apply the fused input transform (GroupNorm /
//     BatchNorm affine, ReLU, residual add -- or the plane-sweep warp for the first
//     layer), split every value into bf16 hi + bf16 lo and store both in shared
//     memory in the UMMA K-major *no-swizzle* core-matrix layout
//     [chunk of 8 channels][brick row][16 B].  In that layout a 3x3 in-plane tap is
//     just a 16-byte-granular shift of the descriptor start address and the 8-row
//     group stride (SBO) is the brick row pitch, so ONE brick serves all 9 taps
//     (validated on hardware by tests/probe/probe_umma.cu).
//   * the three dz taps are folded into the N dimension: one MMA of input plane z
//     against the weight image rows [kz=2 | kz=1 | kz=0] accumulates into the three
//     TMEM accumulator slots of output planes z-1, z, z+1 at once (N = 3*NCTA).
//     Each input plane is therefore read from HBM/L2 once per tile, and the MMA runs
//     at N = 96 where the SS-mode smem operand read stops being the bound
//     (tests/probe/probe_mma_rate.cu: cycles/MMA = max(N/2, 32+N/4, 47)).
//   * fp32 parity: x*w is evaluated as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo on the bf16
//     tensor pipe with fp32 accumulation in TMEM (~2^-17 relative), because a single
//     TF32/bf16 pass does not hold the 1e-3 end-to-end tolerance through 22 layers.
//   * weights (hi+lo images of all 27 taps, <= 110.6 KB) stay resident in shared
//     memory for the life of the CTA; layers with Cin*Cout > 1024 are split over
//     output-channel groups of NCTA = 1024/Cin handled by different CTAs.
//   * warp roles: warps 0-3 epilogue (TMEM -> registers -> global), warps 4-7 loaders,
//     warp 8 lane 0 issues every tcgen05.mma / tcgen05.commit; smem stages and TMEM
//     slots are handed over with mbarriers.
#pragma once
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "simt_kernels.cuh"

namespace dfm {

// ----------------------------------------------------------------------------------
// host side: weight images
// ----------------------------------------------------------------------------------
inline uint16_t bf16_rn_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf16_bits_to_float(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

constexpr int TC_BX = 8, TC_BY = 16;          // output tile (x, y)
constexpr int TC_PX = TC_BX + 2, TC_PY = TC_BY + 2;
constexpr int TC_ROWS = 186;                   // 180 brick rows padded (== 2 mod 8)
constexpr int TC_CG = 32;                      // channels per pipeline stage
constexpr int TC_STAGE_BYTES = 2 * (TC_CG / 8) * TC_ROWS * 16;  // hi + lo
constexpr int TC_NSTAGE = 4;
constexpr int TC_NSLOT = 8;
constexpr int TC_THREADS = 288;

inline bool tc_supported(int Cin, int Cout, int transposed) {
  if (transposed) return false;
  if (Cin != 32 && Cin != 64) return false;
  const int ncta = 1024 / Cin;
  return Cout % ncta == 0 && Cout <= 64;
}
inline bool tc_geom_supported(const ConvGeom& g) {
  return !g.transposed && g.sd == 1 && g.sh == 1 && g.sw == 1 && g.pd == 1 && g.ph == 1 &&
         g.pw == 1 && g.Do == g.Di && g.Ho == g.Hi && g.Wo == g.Wi;
}

struct TcWeights {
  uint8_t* dev = nullptr;  // [nsplit][image bytes]
  int Cin = 0, Cout = 0, ncta = 0, nsplit = 0;
  size_t image_bytes = 0;

  bool ready() const { return dev != nullptr; }
  void release() {
    if (dev) cudaFree(dev);
    dev = nullptr;
  }
  // packed: [27][Cin][Cout] fp32 (tap = kz*9 + ky*3 + kx)
  bool build(const float* packed, int cin, int cout, std::string* err) {
    release();
    Cin = cin;
    Cout = cout;
    ncta = 1024 / cin;
    nsplit = cout / ncta;
    const int kch = cin / 8, nrow = 3 * ncta;
    const size_t tap_bytes = (size_t)kch * nrow * 16;  // one (dy,dx) tap, hi or lo
    image_bytes = 2 * 9 * tap_bytes;
    std::vector<uint16_t> img((size_t)nsplit * image_bytes / 2);
    for (int s = 0; s < nsplit; ++s)
      for (int hl = 0; hl < 2; ++hl)
        for (int t = 0; t < 9; ++t)
          for (int kc = 0; kc < kch; ++kc)
            for (int r = 0; r < nrow; ++r)
              for (int e = 0; e < 8; ++e) {
                const int kz = 2 - r / ncta;  // rows [kz=2 | kz=1 | kz=0] <-> planes z-1,z,z+1
                const int co = s * ncta + r % ncta, ci = kc * 8 + e;
                const float w = packed[((size_t)(kz * 9 + t) * cin + ci) * cout + co];
                const uint16_t hi = bf16_rn_bits(w);
                const uint16_t lo = bf16_rn_bits(w - bf16_bits_to_float(hi));
                const size_t off = (size_t)s * image_bytes / 2 +
                                   ((((size_t)hl * 9 + t) * kch + kc) * nrow + r) * 8 + e;
                img[off] = hl ? lo : hi;
              }
    if (cudaMalloc(&dev, img.size() * 2) != cudaSuccess ||
        cudaMemcpy(dev, img.data(), img.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
      if (err) *err = "TcWeights: device upload failed";
      release();
      return false;
    }
    return true;
  }
};

// ----------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must not hang the GPU; on timeout flag the error
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  for (int it = 0; it < (1 << 26); ++it) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  atomicExch(err, 1);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem, uint64_t ad, uint64_t bd,
                                          uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
      "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, no-swizzle smem matrix descriptor (LBO: K core-matrix stride, SBO: 8-row group stride)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void split_store(const float v[8], uint8_t* hi_dst, uint8_t* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ----------------------------------------------------------------------------------
// loaders: 8 consecutive channels [c0, c0+8) of input voxel (z, y, x), in bounds
// ----------------------------------------------------------------------------------
struct SrcLoader8 {
  Src s;
  int C, H, W;
  __device__ __forceinline__ void load8(int z, int y, int x, int c0, float v[8]) const {
    const long long base = (((long long)z * H + y) * W + x) * C + c0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t < s.n) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(s.t[t].x + base));
        const float4 b = __ldg(reinterpret_cast<const float4*>(s.t[t].x + base) + 1);
        float u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (s.t[t].scale) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(s.t[t].scale + c0));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(s.t[t].scale + c0) + 1);
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(s.t[t].shift + c0));
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(s.t[t].shift + c0) + 1);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = fmaf(u[i], sc[i], sh[i]);
        }
        if (s.t[t].relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = fmaxf(u[i], 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += u[i];
      }
    }
    if (s.outer_relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
  }
};

struct WarpLoader8 {
  WarpLoader w;
  __device__ __forceinline__ void load8(int z, int y, int x, int c0, float v[8]) const {
    c0 += w.first;
    if (c0 < w.C) {
      const float* p = w.cur + ((long long)(y * w.g.step) * w.g.Wf + x * w.g.step) * w.C + c0;
      const float4 a = __ldg(reinterpret_cast<const float4*>(p));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      return;
    }
    c0 -= w.C;
    float fx, fy;
    warp_coord(w.g, x, y, __ldg(w.depths + z), fx, fy);
    const Taps t = bilinear_taps(fx, fy, w.g.Hf, w.g.Wf);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t.w[k] != 0.f) {
        const float* p = w.prev + (long long)t.off[k] * w.C + c0;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = fmaf(t.w[k], a.x, v[0]); v[1] = fmaf(t.w[k], a.y, v[1]);
        v[2] = fmaf(t.w[k], a.z, v[2]); v[3] = fmaf(t.w[k], a.w, v[3]);
        v[4] = fmaf(t.w[k], b.x, v[4]); v[5] = fmaf(t.w[k], b.y, v[5]);
        v[6] = fmaf(t.w[k], b.z, v[6]); v[7] = fmaf(t.w[k], b.w, v[7]);
      }
    }
  }
};

struct TcParams {
  const uint8_t* wimg;
  float* out;
  int D, H, W, Cout;
  int tiles_x, tiles_y, nsplit, nseg, seg_len;
  int n_items;
  int* err;
};

struct TcItem {
  int x0, y0, split, z_lo, z_hi;
};
__device__ __forceinline__ TcItem tc_decode(const TcParams& p, int item) {
  TcItem it;
  it.split = item % p.nsplit;
  int r = item / p.nsplit;
  const int tx = r % p.tiles_x;
  r /= p.tiles_x;
  const int ty = r % p.tiles_y;
  const int seg = r / p.tiles_y;
  it.x0 = tx * TC_BX;
  it.y0 = ty * TC_BY;
  it.z_lo = seg * p.seg_len;
  it.z_hi = min(p.D, it.z_lo + p.seg_len);
  return it;
}

// ----------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------
template <int CIN, int NCTA, class Loader>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_s1_kernel(TcParams p, Loader ld) {
  constexpr int KCH = CIN / 8;                      // 16-byte channel chunks of the weights
  constexpr int NCG = CIN / TC_CG;                  // pipeline stages per input plane
  constexpr int NROW = 3 * NCTA;                    // weight image rows (3 dz slots)
  constexpr uint32_t TAP_BYTES = KCH * NROW * 16;   // one (dy,dx) tap, hi or lo
  constexpr uint32_t W_BYTES = 2 * 9 * TAP_BYTES;
  constexpr uint32_t A_LBO = TC_ROWS * 16, A_SBO = TC_PX * 16;
  constexpr uint32_t A_HL = (TC_CG / 8) * TC_ROWS * 16;  // hi -> lo array offset in a stage
  constexpr uint32_t B_LBO = NROW * 16, B_SBO = 128;
  constexpr uint32_t TMEM_COLS = TC_NSLOT * NCTA;   // 256 or 128 (power of two >= 32)

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + TC_NSTAGE * TC_STAGE_BYTES);
  // bars: [0,NSTAGE) full_a, [NSTAGE,2NSTAGE) empty_a, then NSLOT full_acc, NSLOT empty_acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NSTAGE + 2 * TC_NSLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_a = [&](int s) { return bar0 + 8u * s; };
  auto empty_a = [&](int s) { return bar0 + 8u * (TC_NSTAGE + s); };
  auto full_acc = [&](int s) { return bar0 + 8u * (2 * TC_NSTAGE + s); };
  auto empty_acc = [&](int s) { return bar0 + 8u * (2 * TC_NSTAGE + TC_NSLOT + s); };

  // resident weights of this CTA's output-channel group (blockIdx.x % nsplit is constant
  // over the items of a CTA because gridDim.x is a multiple of nsplit)
  {
    const int split = blockIdx.x % p.nsplit;
    const uint4* src = reinterpret_cast<const uint4*>(p.wimg + (size_t)split * W_BYTES);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (uint32_t i = tid; i < W_BYTES / 16; i += TC_THREADS) dst[i] = __ldg(src + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < TC_NSTAGE; ++s) {
      mbar_init(full_a(s), 128);
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < TC_NSLOT; ++s) {
      mbar_init(full_acc(s), 1);
      mbar_init(empty_acc(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ============================ loaders ============================
    const int lt = tid - 128;
    constexpr int NITEM = (TC_PX * TC_PY * (TC_CG / 8) + 127) / 128;  // 720 / 128 -> 6
    uint32_t stage_ctr = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const TcItem it = tc_decode(p, item);
      int soff[NITEM], gx[NITEM], gy[NITEM];
      bool inb[NITEM], live[NITEM];
#pragma unroll
      for (int k = 0; k < NITEM; ++k) {
        const int i = lt + k * 128;
        live[k] = i < TC_PX * TC_PY * (TC_CG / 8);
        const int chunk = i % (TC_CG / 8), pos = i / (TC_CG / 8);
        const int bx = pos % TC_PX, by = pos / TC_PX;
        gx[k] = it.x0 - 1 + bx;
        gy[k] = it.y0 - 1 + by;
        inb[k] = live[k] && gx[k] >= 0 && gx[k] < p.W && gy[k] >= 0 && gy[k] < p.H;
        soff[k] = (chunk * TC_ROWS + pos) * 16;
      }
      const int chunk = lt % (TC_CG / 8);
      const int zi0 = max(it.z_lo - 1, 0), zi1 = min(it.z_hi, p.D - 1);
      for (int zi = zi0; zi <= zi1; ++zi) {
#pragma unroll 1
        for (int cg = 0; cg < NCG; ++cg, ++stage_ctr) {
          const int s = stage_ctr % TC_NSTAGE;
          mbar_wait(empty_a(s), ((stage_ctr / TC_NSTAGE) & 1) ^ 1, p.err);
          uint8_t* st = a_s + s * TC_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < NITEM; ++k) {
            if (!live[k]) continue;
            float v[8];
            if (inb[k]) {
              ld.load8(zi, gy[k], gx[k], cg * TC_CG + chunk * 8, v);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = 0.f;
            }
            split_store(v, st + soff[k], st + A_HL + soff[k]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(full_a(s));
        }
      }
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t w_base = smem_u32(w_s), a_base = smem_u32(a_s);
      uint32_t stage_ctr = 0, plane_ctr = 0;  // plane_ctr: running index of output planes
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const TcItem it = tc_decode(p, item);
        const int zi0 = max(it.z_lo - 1, 0), zi1 = min(it.z_hi, p.D - 1);
        // output plane zo of this item lives in slot (plane_base + zo - z_lo) % NSLOT
        const uint32_t plane_base = plane_ctr;
        for (int zi = zi0; zi <= zi1; ++zi) {
          const int zo_a = max(zi - 1, it.z_lo), zo_b = min(zi + 1, it.z_hi - 1);
          // a slot is fresh when this is the first input plane that touches it
          for (int zo = zo_a; zo <= zo_b; ++zo) {
            if (zi == max(zo - 1, zi0)) {
              const uint32_t j = plane_base + (zo - it.z_lo);
              mbar_wait(empty_acc(j % TC_NSLOT), ((j / TC_NSLOT) & 1) ^ 1, p.err);
            }
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
          for (int cg = 0; cg < NCG; ++cg, ++stage_ctr) {
            const int s = stage_ctr % TC_NSTAGE;
            mbar_wait(full_a(s), (stage_ctr / TC_NSTAGE) & 1, p.err);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_st = a_base + s * TC_STAGE_BYTES;
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t a_tap = a_st + ((tap / 3) * TC_PX + tap % 3) * 16;
              const uint32_t b_tap = w_base + tap * TAP_BYTES + (cg * (TC_CG / 8)) * B_LBO;
              // runs of output planes with the same fresh state, not crossing the ring end
              int zo = zo_a;
              while (zo <= zo_b) {
                const bool fresh = (cg == 0 && tap == 0) && (zi == max(zo - 1, zi0));
                const uint32_t j0 = plane_base + (zo - it.z_lo);
                int zend = zo;
                while (zend + 1 <= zo_b &&
                       ((cg == 0 && tap == 0) && (zi == max(zend, zi0))) == fresh &&
                       (j0 + (zend + 1 - zo)) % TC_NSLOT != 0)
                  ++zend;
                const int nrun = zend - zo + 1;
                const uint32_t n0 = (uint32_t)(zo - (zi - 1)) * NCTA;  // first weight-image row
                const uint32_t d_tmem = tmem_base + (j0 % TC_NSLOT) * NCTA;
                const uint32_t idesc = idesc_bf16(nrun * NCTA);
#pragma unroll
                for (int ks = 0; ks < TC_CG / 16; ++ks) {
                  const uint32_t a_hi = a_tap + 2 * ks * A_LBO, a_lo = a_hi + A_HL;
                  const uint32_t b_hi = b_tap + 2 * ks * B_LBO + n0 * 16;
                  const uint32_t b_lo = b_hi + 9 * TAP_BYTES;
                  const uint64_t dah = umma_desc(a_hi, A_LBO, A_SBO);
                  const uint64_t dal = umma_desc(a_lo, A_LBO, A_SBO);
                  const uint64_t dbh = umma_desc(b_hi, B_LBO, B_SBO);
                  const uint64_t dbl = umma_desc(b_lo, B_LBO, B_SBO);
                  umma_bf16(d_tmem, dah, dbh, idesc, (fresh && ks == 0) ? 0u : 1u);
                  umma_bf16(d_tmem, dal, dbh, idesc, 1u);
                  umma_bf16(d_tmem, dah, dbl, idesc, 1u);
                }
                zo = zend + 1;
              }
            }
            umma_commit(empty_a(s));  // stage may be refilled once these MMAs retire
          }
          // completed output planes
          if (zi - 1 >= it.z_lo) {
            const uint32_t j = plane_base + (zi - 1 - it.z_lo);
            umma_commit(full_acc(j % TC_NSLOT));
          }
          if (zi == zi1 && zi <= it.z_hi - 1) {
            const uint32_t j = plane_base + (zi - it.z_lo);
            umma_commit(full_acc(j % TC_NSLOT));
          }
        }
        plane_ctr += it.z_hi - it.z_lo;
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue (warps 0-3) ============================
    uint32_t plane_ctr = 0;
    const int m = warp * 32 + lane;  // accumulator row == TMEM lane
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const TcItem it = tc_decode(p, item);
      const int x = it.x0 + (m & 7), y = it.y0 + (m >> 3);
      const bool ok = x < p.W && y < p.H;
      for (int zo = it.z_lo; zo < it.z_hi; ++zo, ++plane_ctr) {
        const int slot = plane_ctr % TC_NSLOT;
        mbar_wait(full_acc(slot), (plane_ctr / TC_NSLOT) & 1, p.err);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[NCTA];
        tmem_ld<NCTA>(tmem_base + ((uint32_t)(warp * 32) << 16) + slot * NCTA, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_acc(slot));
        if (ok) {
          float4* dst = reinterpret_cast<float4*>(
              p.out + (((long long)zo * p.H + y) * p.W + x) * p.Cout + it.split * NCTA);
#pragma unroll
          for (int q = 0; q < NCTA / 4; ++q)
            dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                 __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS));
}

// ----------------------------------------------------------------------------------
// host launch
// ----------------------------------------------------------------------------------
struct TcErrFlag {
  int* dev = nullptr;
  int* get() {
    if (!dev) {
      cudaMalloc(&dev, sizeof(int));
      cudaMemset(dev, 0, sizeof(int));
    }
    return dev;
  }
};
inline TcErrFlag& tc_err_flag() {
  static TcErrFlag f;
  return f;
}
// returns non-zero if any tensor-core kernel timed out on a barrier since the last call
inline int tc_consume_error() {
  int h = 0;
  if (tc_err_flag().dev) {
    cudaMemcpy(&h, tc_err_flag().dev, sizeof(int), cudaMemcpyDeviceToHost);
    if (h) cudaMemset(tc_err_flag().dev, 0, sizeof(int));
  }
  return h;
}

template <int CIN, int NCTA, class Loader>
bool tc_launch(const Loader& ld, const TcWeights& w, float* out, const ConvGeom& g,
               cudaStream_t st, std::string* err) {
  constexpr int KCH = CIN / 8, NROW = 3 * NCTA;
  constexpr size_t W_BYTES = (size_t)2 * 9 * KCH * NROW * 16;
  const size_t smem = W_BYTES + (size_t)TC_NSTAGE * TC_STAGE_BYTES +
                      (2 * TC_NSTAGE + 2 * TC_NSLOT) * 8 + 16;
  static bool attr_set = false;
  auto kern = conv_tc_s1_kernel<CIN, NCTA, Loader>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      if (err) *err = "conv_tc: cannot reserve shared memory";
      return false;
    }
    attr_set = true;
  }
  int sms = 148;
  {
    static int cached = 0;
    if (!cached) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    }
    sms = cached;
  }
  TcParams p{};
  p.wimg = w.dev;
  p.out = out;
  p.D = g.Do;
  p.H = g.Ho;
  p.W = g.Wo;
  p.Cout = g.Cout;
  p.tiles_x = (g.Wo + TC_BX - 1) / TC_BX;
  p.tiles_y = (g.Ho + TC_BY - 1) / TC_BY;
  p.nsplit = w.nsplit;
  const int base = p.tiles_x * p.tiles_y * p.nsplit;
  const int grid_cap = std::max(p.nsplit, sms / p.nsplit * p.nsplit);
  // z segments: balance the persistent grid against the two halo planes a cut costs
  int best_seg = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= std::min(g.Do, 16); ++ns) {
    const int len = (g.Do + ns - 1) / ns;
    const int nseg = (g.Do + len - 1) / len;
    const long long items = (long long)base * nseg;
    const long long grid = std::min<long long>(items, grid_cap);
    const long long rounds = (items + grid - 1) / grid;
    const double cost = (double)rounds * (len + 2);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best_seg = nseg;
      p.seg_len = len;
    }
  }
  p.nseg = best_seg;
  p.n_items = base * p.nseg;
  p.err = tc_err_flag().get();
  int grid = std::min(p.n_items, grid_cap);
  grid = std::max(p.nsplit, grid / p.nsplit * p.nsplit);
  kern<<<grid, TC_THREADS, smem, st>>>(p, ld);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("conv_tc launch: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

template <class Loader>
bool tc_dispatch(const Loader& ld, const TcWeights& w, float* out, const ConvGeom& g,
                 cudaStream_t st, std::string* err) {
  if (g.Cin == 32) return tc_launch<32, 32, Loader>(ld, w, out, g, st, err);
  if (g.Cin == 64) return tc_launch<64, 16, Loader>(ld, w, out, g, st, err);
  if (err) *err = "conv_tc: unsupported Cin";
  return false;
}

inline bool tc_conv_src(const Src& s, const TcWeights& w, float* out, const ConvGeom& g,
                        cudaStream_t st, std::string* err) {
  if (s.n > 2) {
    if (err) *err = "conv_tc: at most two input terms";
    return false;
  }
  SrcLoader8 ld{s, g.Cin, g.Hi, g.Wi};
  return tc_dispatch(ld, w, out, g, st, err);
}
inline bool tc_conv_warp(const WarpLoader& wl, const TcWeights& w, float* out,
                         const ConvGeom& g, cudaStream_t st, std::string* err) {
  WarpLoader8 ld{wl};
  return tc_dispatch(ld, w, out, g, st, err);
}

}  // namespace dfm

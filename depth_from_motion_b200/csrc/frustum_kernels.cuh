// FrustumToVoxel (necks/feature_transformation.py:68-173):
// the pseudo-lidar voxel grid is projected into the plane-sweep frustum, the 32-channel
// stereo feature volume, the depth distribution and the 2-D semantic feature are sampled
// there, and the 64-channel result feeds Conv3d + GroupNorm + ReLU + AvgPool3d((4,1,1)).
#pragma once
#include "common.cuh"

namespace dfm {

struct FrustumParams {
  float P[12];            // cam2img[:3], row-major 3x4
  float pad_w, pad_h;     // img_metas[0]['pad_shape']
  float dmin, dspan;      // depth_cfg depth_min, depth_max - depth_min
  int D, Ho, Wo;          // stereo feature volume (low-res plane sweep)
  int Hs, Ws;             // semantic feature map
  int f;                  // softmax volume is [f*D][f*Ho][f*Wo]
  int nx, ny, nz;
  int cat_img, sem_atten, stereo_atten;
};

// grid_sample(align_corners=True) unnormalisation of a [-1,1] coordinate
__device__ __forceinline__ float fr_unnorm(float n, int size) {
  return (n + 1.f) * 0.5f * (float)(size - 1);
}

// One softmax tap: value of the x`f` upsampled (align_corners=True, depth_head.py:196-199),
// depth-softmaxed logits at integer full-res (kz, ky, kx), rebuilt from the low-res logits and
// the per-pixel (max, 1/sum) written by depth_head_kernel.
__device__ __forceinline__ float fr_softmax_tap(const float* __restrict__ cost,
                                                const float2* __restrict__ norm, int D, int Ho,
                                                int Wo, int f, int kz, int ky, int kx) {
  const int OW = Wo * f, OH = Ho * f, OD = D * f;
  const float sx = OW > 1 ? (float)(Wo - 1) / (OW - 1) : 0.f;
  const float sy = OH > 1 ? (float)(Ho - 1) / (OH - 1) : 0.f;
  const float sz = OD > 1 ? (float)(D - 1) / (OD - 1) : 0.f;
  const float fx = sx * kx, fy = sy * ky, fz = sz * kz;
  const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
  const int x1 = x0 + (x0 < Wo - 1 ? 1 : 0), y1 = y0 + (y0 < Ho - 1 ? 1 : 0);
  const int z1 = z0 + (z0 < D - 1 ? 1 : 0);
  const float lx1 = fx - x0, ly1 = fy - y0, lz1 = fz - z0;
  const int o[4] = {y0 * Wo + x0, y0 * Wo + x1, y1 * Wo + x0, y1 * Wo + x1};
  const float w[4] = {1.f - lx1, lx1, 1.f - ly1, ly1};
  const long long plane = (long long)Ho * Wo;
  const float b0 = dh_plane(cost, z0 * plane, o, w);
  const float b1 = z1 == z0 ? b0 : dh_plane(cost, z1 * plane, o, w);
  const float v = (1.f - lz1) * b0 + lz1 * b1;
  const float2 mn = __ldg(norm + (long long)ky * OW + kx);
  return __expf(v - mn.x) * mn.y;
}

// One axis of a (bi/tri)linear tap pair under grid_sample's zero padding: element offsets of
// the two taps (clamped into range so the load is always legal) and their weights (0 when the
// tap is outside or the voxel is not sampled at all).
__device__ __forceinline__ void fr_axis(float g, int size, int stride, bool on, int& o0, int& o1,
                                        float& w0, float& w1) {
  const float g0 = floorf(g);
  const float a = g - g0;
  const int i0 = (int)g0, i1 = i0 + 1;
  w0 = (on && i0 >= 0 && i0 < size) ? 1.f - a : 0.f;
  w1 = (on && i1 >= 0 && i1 < size) ? a : 0.f;
  o0 = min(max(i0, 0), size - 1) * stride;
  o1 = min(max(i1, 0), size - 1) * stride;
}

// One warp per 32 consecutive voxels.  Phase A, lane = voxel: projection, validity, the depth
// distribution (8 scattered taps per voxel).  Phase B, 8 lanes x float4 per voxel: the warp walks
// its 32 voxels four at a time, broadcasting each one's sample position inside its lane group,
// and moves 128-byte channel rows (8 stereo taps + 4 semantic taps in, 2 rows out).  (The first version recomputed the projection in all
// 32 lanes of a per-voxel warp and was instruction-issue bound at 515 instructions/voxel.)
// stereo: channels-last [D][Ho][Wo][32]; sem: channels-last [Hs][Ws][32]; softmax: materialised
// [fD][fH][fW] volume or nullptr, in which case (cost, norm) rebuild the taps; out:
// channels-last [nz][ny][nx][32 or 64].
__global__ void __launch_bounds__(256)
frustum_gather_kernel(FrustumParams p, const float* __restrict__ xs, const float* __restrict__ ys,
                      const float* __restrict__ zs, const float* __restrict__ stereo,
                      const float* __restrict__ sem, const float* __restrict__ softmax,
                      const float* __restrict__ cost, const float2* __restrict__ norm,
                      float* __restrict__ out) {
  const unsigned FULL = 0xffffffffu;
  const long long v0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  const int lane = threadIdx.x & 31;
  const long long nvox = (long long)p.nx * p.ny * p.nz;
  if (v0 >= nvox) return;
  const int nlive = (int)min(32LL, nvox - v0);

  // ---- phase A: this lane's voxel ---------------------------------------------------------
  const long long vox = min(v0 + lane, nvox - 1);
  const int ix = (int)(vox % p.nx), iy = (int)((vox / p.nx) % p.ny);
  const int iz = (int)(vox / ((long long)p.nx * p.ny));
  // pseudo-lidar (x, y, z) -> rect camera (-y, -z, x)   (feature_transformation.py:175-177)
  const float X = -__ldg(ys + iy), Y = -__ldg(zs + iz), Z = __ldg(xs + ix);
  const float pu = p.P[0] * X + p.P[1] * Y + p.P[2] * Z + p.P[3];
  const float pv = p.P[4] * X + p.P[5] * Y + p.P[6] * Z + p.P[7];
  const float pw = p.P[8] * X + p.P[9] * Y + p.P[10] * Z + p.P[11];
  const float u = pu / pw, v = pv / pw;
  const float nxn = u / (p.pad_w - 1.f) * 2.f - 1.f;
  const float nyn = v / (p.pad_h - 1.f) * 2.f - 1.f;
  const float nzn = (Z - p.dmin) / p.dspan * 2.f - 1.f;
  const bool finite = fabsf(nxn) < 1e8f && fabsf(nyn) < 1e8f && fabsf(nzn) < 1e8f;
  const bool valid2d = finite && u >= 0.f && u <= p.pad_w && v >= 0.f && v <= p.pad_h;
  const bool valid = valid2d && nzn >= -1.f && nzn <= 1.f;

  float disp = 0.f;
  const bool need_disp = p.stereo_atten || (p.sem_atten && p.cat_img);
  if (need_disp && valid) {
    const int OW = p.Wo * p.f, OH = p.Ho * p.f, OD = p.D * p.f;
    const float gx = fr_unnorm(nxn, OW), gy = fr_unnorm(nyn, OH), gz = fr_unnorm(nzn, OD);
    const float x0f = floorf(gx), y0f = floorf(gy), z0f = floorf(gz);
    const float ax = gx - x0f, ay = gy - y0f, az = gz - z0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dx = t & 1, dy = (t >> 1) & 1, dz = t >> 2;
      const int kx = (int)x0f + dx, ky = (int)y0f + dy, kz = (int)z0f + dz;
      if (kx >= 0 && kx < OW && ky >= 0 && ky < OH && kz >= 0 && kz < OD) {
        const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay) * (dz ? az : 1.f - az);
        const float sv = softmax ? __ldg(softmax + ((long long)kz * OH + ky) * OW + kx)
                                 : fr_softmax_tap(cost, norm, p.D, p.Ho, p.Wo, p.f, kz, ky, kx);
        disp = fmaf(wgt, sv, disp);
      }
    }
  }
  // per-axis tap tables of this voxel: clamped element offsets (loads always in bounds) and
  // weights with the zero padding folded in (an out-of-range tap gets weight 0)
  int so[6], mo[4];
  float sw[6], mw[4];
  fr_axis(fr_unnorm(nzn, p.D), p.D, p.Ho * p.Wo * 32, valid, so[0], so[1], sw[0], sw[1]);
  fr_axis(fr_unnorm(nyn, p.Ho), p.Ho, p.Wo * 32, valid, so[2], so[3], sw[2], sw[3]);
  fr_axis(fr_unnorm(nxn, p.Wo), p.Wo, 32, valid, so[4], so[5], sw[4], sw[5]);
  const bool sem_on = valid2d && (!p.sem_atten || disp != 0.f);
  fr_axis(fr_unnorm(nyn, p.Hs), p.Hs, p.Ws * 32, sem_on, mo[0], mo[1], mw[0], mw[1]);
  fr_axis(fr_unnorm(nxn, p.Ws), p.Ws, 32, sem_on, mo[2], mo[3], mw[2], mw[3]);
  if (p.stereo_atten) {
    sw[0] *= disp;
    sw[1] *= disp;
  }
  if (p.sem_atten) {
    mw[0] *= disp;
    mw[1] *= disp;
  }
  const int my_flags = (valid ? 1 : 0) | (sem_on ? 2 : 0);

  // ---- phase B: 8 lanes x float4 per voxel, four voxels per step ---------------------------
  // (one voxel per step with lane = channel was latency-bound: 21 shuffles + 12 loads in a
  // dependent chain per voxel; four independent chains per step hide it)
  const int cout = p.cat_img ? 64 : 32;
  const int grp = lane >> 3, gl = lane & 7;
  const float4* sl = reinterpret_cast<const float4*>(stereo) + gl;
  const float4* ml = reinterpret_cast<const float4*>(sem) + gl;
  for (int jj = 0; jj < 8; ++jj) {
    const int j = jj * 4 + grp;
    if (jj * 4 >= nlive) break;  // warp-uniform
    const int flags = __shfl_sync(FULL, my_flags, j);
    int jo[6];
    float jw[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      jo[i] = __shfl_sync(FULL, so[i], j);
      jw[i] = __shfl_sync(FULL, sw[i], j);
    }
    float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (flags & 1) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int dx = t & 1, dy = (t >> 1) & 1, dz = t >> 2;
        const float wgt = jw[dz] * jw[2 + dy] * jw[4 + dx];
        const float4 f = __ldg(sl + ((jo[dz] + jo[2 + dy] + jo[4 + dx]) >> 2));
        sv.x = fmaf(wgt, f.x, sv.x);
        sv.y = fmaf(wgt, f.y, sv.y);
        sv.z = fmaf(wgt, f.z, sv.z);
        sv.w = fmaf(wgt, f.w, sv.w);
      }
    }
    float* o = out + (v0 + j) * cout + 4 * gl;
    if (j < nlive) *reinterpret_cast<float4*>(o) = sv;
    if (!p.cat_img) continue;
    // semantic feature (bilinear; the reference samples a depth-1 volume at z = 0)
    int ko[4];
    float kw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ko[i] = __shfl_sync(FULL, mo[i], j);
      kw[i] = __shfl_sync(FULL, mw[i], j);
    }
    float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (flags & 2) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int dx = t & 1, dy = t >> 1;
        const float wgt = kw[dy] * kw[2 + dx];
        const float4 f = __ldg(ml + ((ko[dy] + ko[2 + dx]) >> 2));
        mv.x = fmaf(wgt, f.x, mv.x);
        mv.y = fmaf(wgt, f.y, mv.y);
        mv.z = fmaf(wgt, f.z, mv.z);
        mv.w = fmaf(wgt, f.w, mv.w);
      }
    }
    if (j < nlive) *reinterpret_cast<float4*>(o + 32) = mv;
  }
}

// GroupNorm + ReLU + AvgPool3d((4,1,1)) of the last voxel conv, channels-last
// [nz][ny*nx][32] raw -> NCDHW [32][nz/4][ny][nx]   (feature_transformation.py:160-171)
__global__ void __launch_bounds__(256)
frustum_pool_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                    const float* __restrict__ shift, int nzo, long long HW,
                    float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  const long long pos0 = (long long)blockIdx.x * 32;
  const int zo = blockIdx.y;
  const float sc = __ldg(scale + tx), sh = __ldg(shift + tx);
  for (int pp = ty; pp < 32; pp += 8) {
    const long long pos = pos0 + pp;
    float a = 0.f;
    if (pos < HW) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        a += fmaxf(fmaf(__ldg(raw + ((long long)(4 * zo + k) * HW + pos) * 32 + tx), sc, sh), 0.f);
    }
    tile[pp][tx] = a * 0.25f;
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const long long pos = pos0 + tx;
    if (pos < HW) out[((long long)c * nzo + zo) * HW + pos] = tile[tx][c];
  }
}

}  // namespace dfm

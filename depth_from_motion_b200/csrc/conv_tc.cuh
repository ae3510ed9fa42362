// placeholder replaced below
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
namespace dfm {
struct TcWeights {
  bool build(const float*, int, int, std::string*) { return true; }
  bool ready() const { return false; }
  void release() {}
};
inline bool tc_supported(int, int, int) { return false; }
inline bool tc_geom_supported(const ConvGeom&) { return false; }
inline bool tc_conv_src(const Src&, const TcWeights&, float*, const ConvGeom&, cudaStream_t, std::string*) { return false; }
inline bool tc_conv_warp(const WarpLoader&, const TcWeights&, float*, const ConvGeom&, cudaStream_t, std::string*) { return false; }
}

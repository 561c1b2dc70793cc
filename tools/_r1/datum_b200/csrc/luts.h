// datum_b200 — launch interface of the 2D LUT kernels (internal to libdatum_ibl_cuda).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace ibl
{
  struct WaterColorParams
  {
    float deep[3], shallow[3], fresnel[3];
    float depthscale, fresnelbias, fresnelpower;
  };

  // tools/ibl.cpp:189-237, 292-308: split-sum environment BRDF LUT; words and/or fp32 triples
  cudaError_t launch_envbrdf(int width, int height, int samples, uint32_t *words, float *f32, cudaStream_t stream);

  // tools/ibl.cpp:312-329
  cudaError_t launch_watercolor(WaterColorParams const &params, int width, int height, uint32_t *words, cudaStream_t stream);

  // register-resident FFMA chains: `iters` x 16 FMAs per thread, for the FP32 roofline denominator
  cudaError_t launch_fma_peak(float *sink, int blocks, int threads, int iters, cudaStream_t stream);

  // same flop count per thread issued as 8 two-wide fma.rn.f32x2 (FFMA2) chains
  cudaError_t launch_fma2_peak(float *sink, int blocks, int threads, int iters, cudaStream_t stream);
}

// datum_b200 — launch interface of the equirect -> cube stage (internal to libdatum_ibl_cuda).
#pragma once

#include "ibl_math.cuh"

#include <cuda_runtime.h>
#include <cstdint>

namespace ibl
{
  struct ResampleParams
  {
    float4 const *image;   // HDRImage::bits, imgw*imgh RGBA fp32 (tools/hdr.h:24)
    int imgw, imgh;
    int width, height;     // cube face size
    float area_x, area_y;  // tools/hdr.cpp:345
    Quatf quats[6];        // tools/hdr.cpp:335-343 (same table as tools/ibl.cpp:253-261)
    uint32_t *dst;         // 6*width*height rgbe words
  };

  // tools/hdr.cpp:347-356: box-filtered equirect lookup per cube texel, packed with rgbe()
  cudaError_t launch_equirect_resample(ResampleParams const &p, cudaStream_t stream);

  // tools/assetbuilder.cpp:443-462: six ARGB32 face images -> level 0 of the payload.  Per pixel
  // rgbe(srgba(pixel)) (color.h:125-128: c/255 then pow 2.2, table `lut` of the 256 possible values),
  // rows mirrored vertically (QImage::mirrored), faces in argument order.  `argb` = 6*width*height pixels.
  cudaError_t launch_ingest_argb32(uint32_t const *argb, float const *lut, int width, int height, uint32_t *dst, int sm_count, cudaStream_t stream);

  // tools/hdr.cpp:173-318 on one level: the twelve 0.3/0.4/0.3 edge blends, in the reference's order
  cudaError_t launch_blend_edges(uint32_t *level, int width, int height, cudaStream_t stream);
}
